"""mvedit_b200 -- B200-native (sm_100a) hot path for MVEdit's multi-view denoise -> 3D-adapter loop.

Host side mirrors the reference's operator interface for the path; the arithmetic
lives in libmvedit_b200.so (hand-written CUDA behind the C ABI of include/mvedit_b200.h).
There is no CPU fallback: calling an op without the built library or without a CUDA device raises.
"""
__version__ = '0.1.0'
