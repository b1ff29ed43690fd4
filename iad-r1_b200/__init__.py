"""iad-r1_b200: B200-native SC-GRPO / PA-SFT hot path for IAD-R1.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic on the hot path runs in
hand-written sm_100a CUDA behind the C ABI declared in include/iadr1_b200.h (libiadr1_b200.so).
"""
__version__ = "0.1.0"
