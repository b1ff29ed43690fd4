"""Context measurement, not a bench line: the reference's OWN rollout engine (vLLM, `self.llm.generate`,
ref: train/stage_rl/trainer/sc_grpo_trainer.py:343-358, 667) on this box, at the rollout shape of bench.py's headline
configuration - 16 prompts x G = 8 copies = 128 rows, C = 512 tokens each, temperature 0.9 / top_k 50 / top_p 0.9, prefix
caching on (the reference submits every prompt G times and lets vLLM share the prefix).

What it is NOT: there is no tokenizer / processor / checkpoint offline, so the model is the TEXT decoder of
Qwen2.5-VL-3B (Qwen2ForCausalLM at the same widths, dummy weights) fed token ids - the vision tower and the image
tokens' prefill are left out, which favours vLLM slightly. vLLM is library code (like cuBLAS); nothing in the product
imports it.

    python tools/vllm_rollout_probe.py [--rows 128] [--groups 16] [--prompt-len 297] [--completion-len 512]
prints one JSON line {"engine": "vllm", "tok_per_s": ..., "ms_per_rollout": ...}.
"""
import argparse
import json
import os
import tempfile
import time

os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")
os.environ.setdefault("VLLM_ENABLE_V1_MULTIPROCESSING", "0")
os.environ.setdefault("VLLM_LOGGING_LEVEL", "WARNING")
os.environ.setdefault("VLLM_NO_USAGE_STATS", "1")
os.environ.setdefault("DO_NOT_TRACK", "1")

SHAPES = {
    "qwen2.5-vl-3b": dict(hidden_size=2048, intermediate_size=11008, num_hidden_layers=36, num_attention_heads=16,
                          num_key_value_heads=2, vocab_size=151936, tie_word_embeddings=True),
    "qwen2.5-vl-7b": dict(hidden_size=3584, intermediate_size=18944, num_hidden_layers=28, num_attention_heads=28,
                          num_key_value_heads=4, vocab_size=152064, tie_word_embeddings=False),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="qwen2.5-vl-3b", choices=sorted(SHAPES))
    ap.add_argument("--groups", type=int, default=16)
    ap.add_argument("--num-generations", type=int, default=8)
    ap.add_argument("--prompt-len", type=int, default=297)
    ap.add_argument("--completion-len", type=int, default=512)
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--eager", action="store_true")
    a = ap.parse_args()

    import random
    from vllm import LLM, SamplingParams
    d = tempfile.mkdtemp(prefix="vllm_probe_")
    cfg = dict(architectures=["Qwen2ForCausalLM"], model_type="qwen2", max_position_embeddings=32768, rms_norm_eps=1e-6,
               rope_theta=1000000.0, torch_dtype="bfloat16", hidden_act="silu", bos_token_id=151643, eos_token_id=151645,
               use_sliding_window=False, **SHAPES[a.model])
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(cfg, f)
    rows = a.groups * a.num_generations
    t0 = time.perf_counter()
    llm = LLM(model=d, load_format="dummy", skip_tokenizer_init=True, dtype="bfloat16", seed=0,
              max_model_len=a.prompt_len + a.completion_len + 16, gpu_memory_utilization=0.5, enable_prefix_caching=True,
              max_num_seqs=rows, generation_config="vllm", enforce_eager=a.eager)
    t_init = time.perf_counter() - t0
    sp = SamplingParams(n=1, temperature=0.9, top_k=50, top_p=0.9, max_tokens=a.completion_len, ignore_eos=True,
                        detokenize=False)
    times = []
    for rep in range(a.repeats + 1):                         # first pass is the warm-up
        rng = random.Random(rep)
        prompts = []
        for _ in range(a.groups):                            # fresh prompts every pass: no cross-pass prefix hits
            ids = [rng.randrange(1000, 150000) for _ in range(a.prompt_len)]
            prompts += [{"prompt_token_ids": ids}] * a.num_generations
        t = time.perf_counter()
        outs = llm.generate(prompts, sampling_params=sp, use_tqdm=False)
        dt = time.perf_counter() - t
        n_tok = sum(len(c.token_ids) for o in outs for c in o.outputs)
        assert n_tok == rows * a.completion_len, n_tok
        if rep:
            times.append(dt)
    t_s = sum(times) / len(times)
    import vllm
    print(json.dumps({"engine": f"vllm {vllm.__version__}", "model": a.model + " text decoder, dummy weights", "rows": rows,
                      "prompt_len": a.prompt_len, "completion_len": a.completion_len, "cuda_graphs": not a.eager,
                      "tok_per_s": rows * a.completion_len / t_s, "ms_per_rollout": 1000 * t_s,
                      "ms_per_decode_step_upper_bound": 1000 * t_s / a.completion_len, "init_s": t_init,
                      "passes_timed": len(times)}), flush=True)


if __name__ == "__main__":
    main()
