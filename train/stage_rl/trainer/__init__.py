"""Drop-in for ref: train/stage_rl/trainer/__init__.py (`from trainer import SCGRPOTrainer`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from iad_r1_b200.trainer import SCGRPOTrainer  # noqa: E402,F401

__all__ = ["SCGRPOTrainer"]
