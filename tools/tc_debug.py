"""Diagnostic for the tcgen05 limb GEMM: structured inputs that isolate the limb
pairs, tile edges and K chunking, compared with numpy.  Prints a compact report."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aby3_b200 import abi  # noqa: E402

lib = abi.lib
U64 = np.uint64


def ref(a0, a1, b0, b1):
    a0, a1, b0, b1 = (x.astype(object) for x in (a0, a1, b0, b1))
    r = a0.dot(b0 + b1) + a1.dot(b0)
    return np.array(r % (1 << 64), dtype=object).astype(U64)


def run(ctx, a0, a1, b0, b1, algo, acc=None):
    M, K = a0.shape
    N = b0.shape[1]
    d = [ctx.upload(x.astype(U64)) for x in (a0, a1, b0, b1)]
    c = ctx.upload(acc.astype(U64) if acc is not None else np.zeros((M, N), U64))
    abi.check(lib.aby3cu_gemm_cross(ctx.h, algo, d[0].p, d[1].p, d[2].p, d[3].p, M, K, N, c.p, int(acc is not None)))
    out = ctx.download(c, (M, N), U64)
    for x in d + [c]:
        x.free()
    return out


def report(name, got, exp):
    bad = np.argwhere(got != exp)
    if len(bad) == 0:
        print("OK   ", name)
        return True
    print("FAIL ", name, "mismatches", len(bad), "of", got.size)
    for (r, c) in bad[:6]:
        print("      [%d,%d] got %016x exp %016x" % (r, c, int(got[r, c]), int(exp[r, c])))
    rows, cols = np.unique(bad[:, 0]), np.unique(bad[:, 1])
    print("      bad rows", rows[:10], "... bad cols", cols[:10])
    return False


def main():
    ctx = abi.Ctx(0)
    rng = np.random.default_rng(0)
    TC = abi.GEMM_TCGEN05
    ok = True
    z = lambda m, n: np.zeros((m, n), U64)
    # 1. only limb 0 of everything, one tile, one k block
    M, K, N = 128, 32, 64
    a0 = rng.integers(0, 256, (M, K)).astype(U64); b0 = rng.integers(0, 256, (K, N)).astype(U64)
    ok &= report("limb0 x limb0, A0*B0 only", run(ctx, a0, z(M, K), b0, z(K, N), TC), ref(a0, z(M, K), b0, z(K, N)))
    ok &= report("limb0, A1*B0 only (second K half)", run(ctx, z(M, K), a0, b0, z(K, N), TC), ref(z(M, K), a0, b0, z(K, N)))
    ok &= report("limb0, A0*B1 only", run(ctx, a0, z(M, K), z(K, N), b0, TC), ref(a0, z(M, K), z(K, N), b0))
    # 2. each limb pair
    for i in range(8):
        for j in range(8 - i):
            a = (rng.integers(0, 256, (M, K)).astype(U64)) << U64(8 * i)
            b = (rng.integers(0, 256, (K, N)).astype(U64)) << U64(8 * j)
            if not report("limb pair (%d,%d)" % (i, j), run(ctx, a, z(M, K), b, z(K, N), TC), ref(a, z(M, K), b, z(K, N))):
                ok = False
    # 3. full random, growing shapes
    for (M, K, N) in [(128, 32, 64), (128, 64, 64), (128, 256, 64), (256, 96, 128), (130, 70, 66), (512, 512, 512),
                      (1000, 300, 200), (1024, 1024, 1024)]:
        a0, a1 = (rng.integers(0, 2**64, (M, K), dtype=U64) for _ in range(2))
        b0, b1 = (rng.integers(0, 2**64, (K, N), dtype=U64) for _ in range(2))
        exp = run(ctx, a0, a1, b0, b1, abi.GEMM_IMAD)
        if M * K * N <= 128 * 256 * 64:
            assert np.array_equal(exp, ref(a0, a1, b0, b1)), "IMAD reference wrong"
        ok &= report("random %dx%dx%d" % (M, K, N), run(ctx, a0, a1, b0, b1, TC), exp)
        acc = rng.integers(0, 2**64, (M, N), dtype=U64)
        ok &= report("random %dx%dx%d accumulate" % (M, K, N), run(ctx, a0, a1, b0, b1, TC, acc), exp + acc)
    # 4. worst case for the exactness bound: all limbs 0xFF, K at the per-launch bound, and beyond it
    for K in (8192, 8256 + 40):
        M, N = 128, 64
        a = np.full((M, K), 2**64 - 1, U64); b = np.full((K, N), 2**64 - 1, U64)
        # A0*(B0+B1) + A1*B0 with B1 = 0
        exp = run(ctx, a, a, b, z(K, N), abi.GEMM_IMAD)
        ok &= report("all-ones K=%d" % K, run(ctx, a, a, b, z(K, N), TC), exp)
    print("ALL OK" if ok else "SOME FAILED")
    ctx.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
