"""uncltmo_b200 - B200-native implementation of the UnCLTMO tone-mapping hot path.

Python/PyTorch is the host (device memory, streams, torch.distributed); every operator on the path is a
hand-written sm_100a CUDA kernel in libuncltmo_b200.so, reached through the C ABI in include/uncltmo_b200.h.
There is no CPU fallback and no Triton / torch.compile path.
"""
__version__ = "0.1.0"
