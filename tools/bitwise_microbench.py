import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from aby3_b200 import abi
lib = abi.lib
ctx = abi.Ctx(0)
n = 1 << 23
rb = int(lib.aby3cu_bin_row_bytes(n))
bufs = [ctx.alloc(8 * n) for _ in range(6)]
for i, b in enumerate(bufs[:4]):
    abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, bytes([i + 1] * 16), 0, b.p, 8 * n))
kp, kn = bytes(range(16)), bytes(range(100, 116))
for bits in (64, 32, 8):
    for it in range(3):
        e0, e1 = ctx.event(), ctx.event()
        ctx.record(e0)
        abi.check(lib.aby3cu_bin_bitwise_rowmajor(ctx.h, 8, bufs[0].p, bufs[1].p, bufs[2].p, bufs[3].p, bufs[4].p, None, n, bits, rb, kp, kn, 0))
        ctx.record(e1)
        print("bits", bits, "ms", abi.elapsed_ms(e0, e1), flush=True)
