"""Small, fast invocations of the places where a race or an out-of-bounds access would be silent, meant to run under
compute-sanitizer (tools/sanitize.sh): the tcgen05 GEMM's hand-rolled mbarrier pipeline, the persistent SGD kernel's
split grid barrier, the early-free pool + second stream of the truncating product, the binary engine.  Every target
checks its result against numpy so that a sanitizer run is also a parity run.

  python tools/sanitize_targets.py {smoke|gemm|sgd|early|binary}
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

U64 = np.uint64


def rnd(seed, shape):
    return np.random.default_rng(seed).integers(-2**63, 2**63, shape, dtype=np.int64)


def t_smoke():
    import __graft_entry__ as g
    g.smoke()


def t_gemm():
    """one 512^3 tcgen05 product through the C ABI (pack + k_gemm_tc), compared with numpy"""
    from aby3_b200 import abi
    ctx = abi.Ctx(0)
    M = K = N = int(os.environ.get("SAN_GEMM_N", "512"))
    a0, a1, b0, b1 = (rnd(i, s) for i, s in ((1, (M, K)), (2, (M, K)), (3, (K, N)), (4, (K, N))))
    dA0, dA1, dB0, dB1 = (ctx.upload(x) for x in (a0, a1, b0, b1))
    c0 = rnd(5, (M, N))
    dC = ctx.upload(c0)
    abi.check(abi.lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_TCGEN05, dA0.p, dA1.p, dB0.p, dB1.p, M, K, N, dC.p, 1))
    got = ctx.download(dC, M * N).reshape(M, N)
    exp = (c0.view(U64) + (a0.view(U64) @ (b0.view(U64) + b1.view(U64))) + a1.view(U64) @ b0.view(U64)).view(np.int64)
    assert np.array_equal(got, exp), "tcgen05 product differs from numpy"
    ctx.close()
    print("gemm ok")


def fixed(x, D=16):
    return (np.asarray(x) * (1 << D)).astype(np.int64)


def t_sgd():
    """k_sgd_linear_slab (feature-resident persistent kernel) and k_sgd_linear (rows/slabs) against the facade loop"""
    from aby3_b200 import harness
    for N, F, B, iters in ((300, 64, 32, 6), (200, 10, 7, 4)):
        rng = np.random.default_rng(7)
        x = rng.normal(1, 1, (N, F))
        y = x[:, :3] @ np.array([[2.0], [-1.0], [0.5]])
        idx = rng.integers(0, N, iters * B).astype(np.uint64)
        res = []
        for fused in (False, True):
            s = harness.Session()
            X, Y, W = s.share_int(0, fixed(x)), s.share_int(0, fixed(y)), s.share_int(0, np.zeros((F, 1), dtype=np.int64))
            (s.linreg_fused if fused else s.linreg)(X, Y, W, idx, iters, B, 2.0 ** -6)
            res.append(s.get_shares(W))
            s.close()
        assert np.array_equal(res[0], res[1]), "fused SGD differs from the facade loop"
    print("sgd ok")


def t_early():
    """truncating products big enough (outputs >= 4 MiB) for the early-free pool and the second stream; several steps
    back to back so that blocks are recycled while readers are still in flight"""
    from aby3_b200 import abi, harness
    os.environ["ABY3_EARLY_TRUNCATION"] = "1"
    s = harness.Session()
    s.set_gemm_algo(abi.GEMM_TCGEN05)
    M, K, N, d = 1024, 128, 768, 16
    rng = np.random.default_rng(9)
    a = (rng.uniform(-4, 4, (M, K)) * (1 << d)).astype(np.int64)
    b = (rng.uniform(-4, 4, (K, N)) * (1 << d)).astype(np.int64)
    A, B = s.share_int(0, a), s.share_int(1, b)
    ref = (a @ b) >> d
    for _ in range(4):
        C = s.mul(A, B, shift=d)
        c = s.reveal(C, 0)
        assert np.max(np.abs(c - ref)) <= 4
        sh = s.get_shares(C)
        for p in range(3):
            assert np.array_equal(sh[(p + 1) % 3, 1], sh[p, 0])
        s.free(C)
    s.close()
    print("early ok")


def t_binary():
    from aby3_b200 import harness
    s = harness.Session()
    width = 5000
    x, y = rnd(5, (width, 1)), rnd(6, (width, 1))
    X, Y = s.share_bin(0, x, 64), s.share_bin(1, y, 64)
    out = s.bin_eval(harness.library_circuit("lt", 64), [X, Y])[0]
    assert np.array_equal(s.reveal(out, 0, binary=True) & 1, (x < y).astype(np.int64))
    xs, ys = x >> 2, y >> 2                                  # cipher_gt = MSB(b - a): keep the difference in range
    A, B = s.share_int(0, xs), s.share_int(1, ys)
    g = s.cipher_gt(A, B)
    assert np.array_equal(s.reveal(g, 0, binary=True) & 1, (xs > ys).astype(np.int64))
    s.close()
    print("binary ok")


if __name__ == "__main__":
    {"smoke": t_smoke, "gemm": t_gemm, "sgd": t_sgd, "early": t_early, "binary": t_binary}[sys.argv[1]]()
