"""B200-native pileup calling hot path of Clair3-RNA (tensor generation + pileup network).

The compute path is the CUDA library built from csrc/ (libc3r_b200.so, C ABI in
include/c3r_b200.h).  There is no CPU fallback: importing `engine` without the
built library raises.
"""
__version__ = "0.1.0"
