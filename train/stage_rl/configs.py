"""Drop-in for ref: train/stage_rl/configs.py (`from configs import GRPOConfig`): same field surface, B200 backend."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from iad_r1_b200.grpo_config import GRPOConfig  # noqa: E402,F401
