"""aby3_b200 -- B200-native implementation of ABY3's replicated-share
multiplication hot path.  The product is native: CUDA kernels
behind a C ABI (include/aby3cu.h, libaby3cu.so) and a C++ sh3 facade
(aby3_b200/sh3, libsh3.so).  The Python modules here only bind those libraries
for tests and bench.py; nothing in this package imports oracle/."""

import os

# A three-party session keeps a dozen streams busy on one GPU (per party: its own stream, the second compute stream, an
# upload and a download stream).  With the default of 8 hardware work queues several of them share a queue, and a kernel
# that is ready waits behind an unrelated one that is not (the truncation pairs issued ahead never ran ahead).  Must be
# set before the CUDA context exists; libaby3cu does the same in aby3cu_ctx_create as a fallback.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

__all__ = ["abi"]
