"""Launches every HBM-/AES-bound kernel of the C ABI twice (warm + measured) on inputs larger than L2, for
`ncu --set full --clock-control none` captures (profiles/): dram bytes, achieved bandwidth, pipe utilisation."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aby3_b200 import abi  # noqa: E402

lib = abi.lib
KA, KB = bytes(range(16)), bytes(range(50, 66))


def main():
    n = 1 << 25                       # 256 MiB per int64 array
    ctx = abi.Ctx(0)
    bufs = [ctx.alloc(8 * n) for _ in range(7)]
    p = [b.p for b in bufs]
    width = 1 << 24
    rb = lib.aby3cu_bin_row_bytes(width)
    mem0, mem1 = ctx.alloc(192 * rb), ctx.alloc(192 * rb)
    gates = np.array([[i, 64 + i, 128 + i, 8] for i in range(64)], dtype=np.uint32)
    dg = ctx.upload(gates)
    xg = np.array([[i, 64 + i, 128 + i, 6] for i in range(64)], dtype=np.uint32)
    dx = ctx.upload(xg)
    for rep in range(2):
        abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, KA, 0, p[0], 8 * n))
        for b in bufs[1:4]:
            abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, KB, 0, b.p, 8 * n))
        abi.check(lib.aby3cu_zero_share(ctx.h, KA, KB, 0, p[1], p[4], n, 0))
        abi.check(lib.aby3cu_mul_hadamard(ctx.h, p[0], p[1], p[2], p[3], KA, KB, 0, p[4], n))
        abi.check(lib.aby3cu_mul_hadamard_trunc(ctx.h, p[0], p[1], p[2], p[3], KA, 4, KB, 4, 16, p[4], p[5], p[6], n))
        abi.check(lib.aby3cu_trunc_tuple(ctx.h, KA, 4, KB, 4, 16, None, p[4], p[5], p[6], n))
        abi.check(lib.aby3cu_trunc_finish(ctx.h, p[0], p[1], p[2], p[3], n, 16))
        abi.check(lib.aby3cu_share_op(ctx.h, 0, p[0], p[1], p[2], n))
        abi.check(lib.aby3cu_combine3(ctx.h, 0, p[0], p[1], p[2], p[3], n))
        abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, KA, 0, mem0.p, 192 * rb))
        abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, KB, 0, mem1.p, 192 * rb))
        abi.check(lib.aby3cu_bit_transpose(ctx.h, p[0], width, 64, 8, mem0.p, rb, None))
        abi.check(lib.aby3cu_bit_transpose(ctx.h, mem0.p, 64, width, rb, p[1], 8, None))
        abi.check(lib.aby3cu_bin_level(ctx.h, dg.p, 64, mem0.p, mem1.p, rb, KA, KB, 0))
        abi.check(lib.aby3cu_bin_level(ctx.h, dx.p, 64, mem0.p, mem1.p, rb, KA, KB, 0))
        abi.check(lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_TCGEN05, p[0], p[1], p[2], p[3], 4096, 4096, 4096, p[4], 1))
        ctx.sync()
    ctx.close()
    print("ok")


if __name__ == "__main__":
    main()
