"""Per-kernel timing through the C ABI (CUDA events on the launching stream,
3 warm-ups, inputs larger than L2 where the kernel is HBM-bound).  Writes one
JSON line per kernel to stdout.  Usage: python tools/microbench.py [--gemm N]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aby3_b200 import abi  # noqa: E402

lib = abi.lib
KA, KB = bytes(range(16)), bytes(range(50, 66))
PEAK = 6542.1
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(ctx, fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = ctx.event(), ctx.event()
    ctx.record(a)
    for _ in range(iters):
        fn()
    ctx.record(b)
    return abi.elapsed_ms(a, b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 26)     # 512 MiB per int64 array
    ap.add_argument("--gemm", type=int, default=1024)
    ap.add_argument("--algo", type=int, default=abi.GEMM_IMAD)
    ap.add_argument("--only-gemm", action="store_true")
    args = ap.parse_args()
    ctx = abi.Ctx(0)
    n = args.n
    bufs = [ctx.alloc(8 * n) for _ in range(7)]
    for b in bufs:
        abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, KA, 0, b.p, 8 * n))
    ctx.sync()
    p = [b.p for b in bufs]

    if args.only_gemm:
        for g in (1024, 2048, 4096):
            for algo in ((1, 2) if g <= 2048 else (2,)):
                ms = timeit(ctx, lambda: abi.check(lib.aby3cu_gemm_cross(ctx.h, algo, p[0], p[1], p[2], p[3], g, g, g, p[4], 1)), iters=5, warm=2)
                print(json.dumps({"kernel": "gemm_cross algo=%d %d^3 (incl. limb pre-pass)" % (algo, g), "ms": round(ms, 3),
                                  "ring_MAC_per_s_one_party": float("%.4g" % (g ** 3 / ms * 1e3)),
                                  "int8_TOPS": round(144 * g ** 3 / ms / 1e9, 1) if algo == 2 else None}), flush=True)
        ctx.close()
        return

    def report(name, ms, bytes_per_elem):
        gbs = bytes_per_elem * n / ms / 1e6
        print(json.dumps({"kernel": name, "ms": round(ms, 4), "GBps": round(gbs, 1),
                          "frac_of_measured_hbm": round(gbs / PEAK, 3), "n": n, "bytes_per_elem": bytes_per_elem}), flush=True)

    report("aes_ctr_fill", timeit(ctx, lambda: abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, KA, 0, p[0], 8 * n))), 8)
    report("zero_share(+addend)", timeit(ctx, lambda: abi.check(lib.aby3cu_zero_share(ctx.h, KA, KB, 0, p[1], p[0], n, 0))), 16)
    report("mul_hadamard(masked)", timeit(ctx, lambda: abi.check(lib.aby3cu_mul_hadamard(ctx.h, p[0], p[1], p[2], p[3], KA, KB, 0, p[4], n))), 40)
    report("mul_hadamard(no mask)", timeit(ctx, lambda: abi.check(lib.aby3cu_mul_hadamard(ctx.h, p[0], p[1], p[2], p[3], None, None, 0, p[4], n))), 40)
    report("mul_hadamard_trunc", timeit(ctx, lambda: abi.check(lib.aby3cu_mul_hadamard_trunc(ctx.h, p[0], p[1], p[2], p[3], KA, 4, KB, 4, 16, p[4], p[5], p[6], n))), 56)
    report("trunc_tuple(negr)", timeit(ctx, lambda: abi.check(lib.aby3cu_trunc_tuple(ctx.h, KA, 4, KB, 4, 16, None, p[4], p[5], p[6], n))), 24)
    report("trunc_finish", timeit(ctx, lambda: abi.check(lib.aby3cu_trunc_finish(ctx.h, p[0], p[1], p[2], p[3], n, 16))), 40)
    report("share_add", timeit(ctx, lambda: abi.check(lib.aby3cu_share_op(ctx.h, 0, p[0], p[1], p[2], n))), 24)
    # GEMV-like cross term (logistic inference / SGD shapes): A planes streamed once, 16 B per (m, k)
    for (Mv, Kv) in ((n // 512, 512), (n // 1024, 1024)):
        ms = timeit(ctx, lambda: abi.check(lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_IMAD, p[0], p[1], p[2], p[3], Mv, Kv, 1, p[4], 1)))
        print(json.dumps({"kernel": "gemm_skinny %dx%dx1" % (Mv, Kv), "ms": round(ms, 4), "GBps": round(16 * Mv * Kv / ms / 1e6, 1),
                          "frac_of_measured_hbm": round(16 * Mv * Kv / ms / 1e6 / PEAK, 3), "bytes_per_mk": 16}), flush=True)
    # binary: transpose 2^24 x 64 and one 64-gate AND level over 2^24 instances
    width = 1 << 24
    rb = lib.aby3cu_bin_row_bytes(width)
    wires = 192
    mem0, mem1 = ctx.alloc(wires * rb), ctx.alloc(wires * rb)
    abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, KA, 0, mem0.p, wires * rb))
    abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, KB, 0, mem1.p, wires * rb))
    ms = timeit(ctx, lambda: abi.check(lib.aby3cu_bit_transpose(ctx.h, p[0], width, 64, 8, mem0.p, rb, None)))
    print(json.dumps({"kernel": "bit_transpose 2^24x64 fwd", "ms": round(ms, 4), "GBps": round(16 * width / ms / 1e6, 1),
                      "frac_of_measured_hbm": round(16 * width / ms / 1e6 / PEAK, 3)}), flush=True)
    ms = timeit(ctx, lambda: abi.check(lib.aby3cu_bit_transpose(ctx.h, mem0.p, 64, width, rb, p[1], 8, None)))
    print(json.dumps({"kernel": "bit_transpose 64x2^24 back", "ms": round(ms, 4), "GBps": round(16 * width / ms / 1e6, 1),
                      "frac_of_measured_hbm": round(16 * width / ms / 1e6 / PEAK, 3)}), flush=True)
    gates = np.array([[i, 64 + i, 128 + i, 8] for i in range(64)], dtype=np.uint32)
    dg = ctx.upload(gates)
    ms = timeit(ctx, lambda: abi.check(lib.aby3cu_bin_level(ctx.h, dg.p, 64, mem0.p, mem1.p, rb, KA, KB, 0)))
    words = 64 * rb / 8
    print(json.dumps({"kernel": "bin_level 64 AND x 2^24", "ms": round(ms, 4), "GBps": round(40 * words / ms / 1e6, 1),
                      "frac_of_measured_hbm": round(40 * words / ms / 1e6 / PEAK, 3), "bytes_per_gate_word": 40}), flush=True)
    ms = timeit(ctx, lambda: abi.check(lib.aby3cu_bin_level(ctx.h, dg.p, 64, mem0.p, mem1.p, rb, None, None, 0)))
    print(json.dumps({"kernel": "bin_level 64 AND x 2^24 (no z)", "ms": round(ms, 4), "GBps": round(40 * words / ms / 1e6, 1),
                      "frac_of_measured_hbm": round(40 * words / ms / 1e6 / PEAK, 3)}), flush=True)
    # GEMM
    g = args.gemm
    assert 8 * g * g <= 8 * n
    ms = timeit(ctx, lambda: abi.check(lib.aby3cu_gemm_cross(ctx.h, args.algo, p[0], p[1], p[2], p[3], g, g, g, p[4], 0)), iters=3, warm=1)
    print(json.dumps({"kernel": "gemm_cross algo=%d %d^3" % (lib.aby3cu_gemm_last_algo(ctx.h), g), "ms": round(ms, 3),
                      "ring_MAC_per_s_one_party": round(g ** 3 / ms * 1e3, 1),
                      "u64_MAC_per_s": round(2 * g ** 3 / ms * 1e3, 1)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
