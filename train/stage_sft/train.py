"""PA-SFT entry point with the reference's CLI (ref: train/stage_sft/train.py:15-28 -> llamafactory `run_exp`), so that
`scripts/train/PA_SFT/*.sh` (`torchrun ... train/stage_sft/train.py --stage sft --do_train --dataset ... --template ...`)
run against the B200-native forward/backward. Only `--stage sft --finetuning_type full` is in scope."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from iad_r1_b200.sft_trainer import run_sft  # noqa: E402


def main():
    run_sft()


if __name__ == "__main__":
    main()
