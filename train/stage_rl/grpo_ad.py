"""SC-GRPO entry point with the reference's CLI (ref: train/stage_rl/grpo_ad.py:31-213), so that
`scripts/train/SC_GRPO/*.sh` (`torchrun ... train/stage_rl/grpo_ad.py --deepspeed ... --num_generations 4 ...`) run
unchanged against the B200-native trainer. The reward callbacks are the user's own `reward.py` (imported from
PYTHONPATH exactly as the reference does, `from reward import *`); they stay CPU Python and are called with the
reference's convention. The dataset mapping below restates `make_conversation` (:135-181).
"""
import logging
import os
import sys
from dataclasses import dataclass, field
from functools import partial
from typing import Optional

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from configs import GRPOConfig  # noqa: E402
from trainer import SCGRPOTrainer  # noqa: E402
from iad_r1_b200.grpo_config import ModelConfig, ScriptArguments, TrlParser, get_peft_config  # noqa: E402

logger = logging.getLogger(__name__)

_FORMAT = (
    "If you find anomalies in the test image, structure your response with the following format:"
    "<think>[Your process of observation and reasoning is here]</think>"
    "<location>[The location of the anomaly in the image]</location>"
    "<type>[The type of anomaly in the image]</type><answer>[Your final answer is here(yes or no)]</answer>"
    "If no anomalies are detected in the test image, structure your response with the following format:"
    "<think>[Your process of observation and reasoning is here]</think>"
    "<answer>[Your final answer is here(yes or no)]</answer>")
SYSTEM_PROMPTS = {
    1: ("You are an expert in detecting anomalies in image. Your task is to detect if there are any anomalies in the test "
        "image." + _FORMAT + "{Question}"),
    0: ("You are an expert in detecting anomalies in images. I will provide you with two images: a reference image (first) "
        "showing a normal object without defects, and a test image (second) that needs inspection."
        "Your task is to compare these images and determine if there are any anomalies in the test image. Use the reference "
        "image as a baseline for what is considered normal." + _FORMAT +
        "Remember that the first image is always the reference (normal) image, and the second image is the test image that "
        "needs inspection.{Question}"),
}
QUESTION_PROMPTS = {
    1: ("You are an expert in detecting defects in image. Your task is to detect if there are any defects in the test image."
        "{Question}"),
    0: ("You are an expert in detecting defects in image. I will provide you with two images: a reference image (first) "
        "showing a normal object without defects, and a test image (second) that needs inspection."
        "Your task is to compare these images and determine if there are any anomalies in the test image. Use the reference "
        "image as a baseline for what is considered normal.{Question}"),
}


@dataclass
class GRPOScriptArguments(ScriptArguments):
    reward_funcs: list[str] = field(default_factory=lambda: ["accuracy", "format"])
    use_vllm_for_gen: str = field(default="true")
    use_system_prompt: str = field(default="false")
    image_path: Optional[str] = field(default="/data")
    max_pixels: Optional[int] = field(default=12845056)
    min_pixels: Optional[int] = field(default=3136)
    single_img: int = field(default=1)


def make_conversation(example, image_path=None, use_system_prompt=False, single_img=1):
    if not ("image" in example and example["image"]):
        return example
    raw = example["image"]
    items = raw if isinstance(raw, list) else [raw]
    images = []
    for item in items:
        if isinstance(item, str):
            images.append(os.path.join(image_path, item))
        elif isinstance(item, dict):
            images.append(os.path.join(image_path, item["path"]))
        else:
            raise TypeError("Unsupported Format.")
    user = [*[{"type": "image"} for _ in images]]
    if use_system_prompt:
        user.append({"type": "text", "text": example["problem"]})
        prompt = [{"role": "system", "content": SYSTEM_PROMPTS[single_img]}, {"role": "user", "content": user}]
    else:
        user.append({"type": "text", "text": QUESTION_PROMPTS[single_img].format(Question=example["problem"])})
        prompt = [{"role": "user", "content": user}]
    return {"prompt": prompt, "image": images}


def main(script_args, training_args, model_args):
    use_system_prompt = script_args.use_system_prompt != "false"
    use_vllm_for_gen = script_args.use_vllm_for_gen != "false"
    if script_args.single_img not in (0, 1):
        raise ValueError("The single_img parameter can only be 0 or 1")
    try:
        import reward as user_rewards  # the user's callbacks (ref: train/stage_rl/reward.py), found via PYTHONPATH
    except ImportError as e:
        raise ImportError("reward.py (accuracy_reward / consistency_reward) must be importable from PYTHONPATH, exactly as "
                          "in the reference tree (train/stage_rl/reward.py + reward_process/)") from e
    registry = {"accuracy": user_rewards.accuracy_reward, "format": user_rewards.consistency_reward}
    reward_funcs = [registry[name] for name in script_args.reward_funcs]

    from datasets import load_dataset
    if not script_args.dataset_name.endswith(".json"):
        raise ValueError("dataset_name must be a .json file of {id, image, problem, solution} rows (README.md:104-119)")
    dataset = load_dataset("json", data_files=script_args.dataset_name)
    dataset = dataset.map(partial(make_conversation, image_path=script_args.image_path,
                                  use_system_prompt=use_system_prompt, single_img=script_args.single_img))
    for split in dataset:
        if "messages" in dataset[split].column_names:
            dataset[split] = dataset[split].remove_columns("messages")

    trainer = SCGRPOTrainer(
        model=model_args.model_name_or_path, reward_funcs=reward_funcs, args=training_args,
        train_dataset=dataset[script_args.dataset_train_split],
        eval_dataset=dataset[script_args.dataset_test_split] if training_args.eval_strategy != "no" else None,
        peft_config=get_peft_config(model_args), attn_implementation=model_args.attn_implementation,
        max_pixels=script_args.max_pixels, min_pixels=script_args.min_pixels, use_vllm_for_gen=use_vllm_for_gen)
    trainer.train()
    trainer.save_model(training_args.output_dir)
    if training_args.push_to_hub:
        trainer.push_to_hub(dataset_name=script_args.dataset_name)


if __name__ == "__main__":
    parser = TrlParser((GRPOScriptArguments, GRPOConfig, ModelConfig))
    script_args, training_args, model_args = parser.parse_args_and_config()
    main(script_args, training_args, model_args)
