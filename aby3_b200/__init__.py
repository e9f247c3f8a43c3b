"""aby3_b200 -- B200-native implementation of ABY3's replicated-share
multiplication hot path.  The product is native: CUDA kernels
behind a C ABI (include/aby3cu.h, libaby3cu.so) and a C++ sh3 facade
(aby3_b200/sh3, libsh3.so).  The Python modules here only bind those libraries
for tests and bench.py; nothing in this package imports oracle/."""

__all__ = ["abi"]
