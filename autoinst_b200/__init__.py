"""autoinst_b200 — B200-native (sm_100a) implementation of AutoInst's chunk-level normalized-cuts path.

Python host code (this package + the drop-in `ncuts` package) calls hand-written CUDA kernels in
`lib/libautoinst_ncuts.so` through the C ABI declared in `include/autoinst_ncuts.h`.
There is no CPU fallback.
"""
__version__ = "0.1.0"
