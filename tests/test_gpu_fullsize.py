"""BASELINE-size checks through size-independent properties (the oracle would need
minutes at these sizes): Freivalds' check of the revealed product, share consistency
between neighbours, agreement of all three reveals, sampled rows against numpy; plus
the edge cases: empty inputs, K beyond the per-launch exactness bound, forced row-blocking."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as o
from aby3_b200 import abi, harness

pytestmark = pytest.mark.gpu
U64 = np.uint64
lib = abi.lib


def rnd(seed, shape):
    return np.random.default_rng(seed).integers(-2**63, 2**63, shape, dtype=np.int64)


def test_si64_matmul_4096_cubed_freivalds():
    """config 1/2 size: reveal(C) v == a (b v) mod 2^64 for random v; replicated-share consistency."""
    n = 4096
    s = harness.Session()
    try:
        a, b = rnd(1, (n, n)), rnd(2, (n, n))
        A, B = s.share_int(0, a), s.share_int(1, b)
        C = s.mul(A, B)
        assert lib.aby3cu_version() == 1
        sh = s.get_shares(C)
        for p in range(3):
            assert np.array_equal(sh[(p + 1) % 3, 1], sh[p, 0])
        c = (sh[0, 0].view(U64) + sh[1, 0].view(U64) + sh[2, 0].view(U64)).view(np.int64)
        for p in range(3):
            assert np.array_equal(s.reveal(C, p), c)
        for seed in range(3):
            v = rnd(10 + seed, (n, 1))
            assert np.array_equal(c @ v, a @ (b @ v))
    finally:
        s.close()


def test_sf64_matmul_trunc_4096_cubed_properties():
    """config 2: shares stay a consistent replicated sharing, all parties reveal the same matrix,
    sampled rows are within the protocol's +-4 ulp of the plaintext fixed-point product."""
    n, d = 4096, 16
    s = harness.Session()
    try:
        rng = np.random.default_rng(3)
        a = (rng.uniform(-4, 4, (n, n)) * (1 << d)).astype(np.int64)
        b = (rng.uniform(-4, 4, (n, n)) * (1 << d)).astype(np.int64)
        A, B = s.share_int(0, a), s.share_int(0, b)
        C = s.mul(A, B, shift=d)
        sh = s.get_shares(C)
        for p in range(3):
            assert np.array_equal(sh[(p + 1) % 3, 1], sh[p, 0])
        c = s.reveal(C, 0)
        assert np.array_equal(s.reveal(C, 1), c) and np.array_equal(s.reveal(C, 2), c)
        rows = rng.integers(0, n, 16)
        ref = (a[rows] @ b) >> d
        assert np.max(np.abs(c[rows] - ref)) <= 4
    finally:
        s.close()


def test_config1_1024_cubed_every_share_plane_bit_exact():
    """BASELINE configs[0]: si64Matrix 1024^3 with the seeds of aby3_tests/Sh3EvaluatorTests.cpp:41-47 (the session
    defaults) and operands drawn like its PRNG(ZeroBlock) fill (:57-63).  ALL six share planes of the product -- the
    tcgen05 path over 128 output tiles and 2^20 zero-share elements -- against the oracle, then the truncating form."""
    n = 1024
    s, r = harness.Session(), o.Session()
    try:
        s.set_gemm_algo(abi.GEMM_TCGEN05)
        ks = o.keystream(bytes(16), 0, 2 * n * n * 8).view(np.int64)
        a, b = ks[:n * n].reshape(n, n).copy(), ks[n * n:].reshape(n, n).copy()
        A, B = s.share_int(0, a), s.share_int(1, b)
        Ao, Bo = r.share_int(0, a), r.share_int(1, b)
        assert np.array_equal(s.get_shares(A), Ao) and np.array_equal(s.get_shares(B), Bo)
        C = s.mul(A, B)
        Co = r.mul(Ao, Bo, nthreads=6)
        assert np.array_equal(s.get_shares(C), Co)
        assert np.array_equal(s.reveal(C, 0), o.plain_mul(a, b, nthreads=4))
        T = s.mul(A, B, shift=16)
        assert np.array_equal(s.get_shares(T), r.mul_trunc(Ao, Bo, 16, nthreads=6))
        for p in range(3):
            assert list(s.cursors(p)) == list(r.cursors(p))
    finally:
        s.close()
        r.close()


@pytest.mark.parametrize("algo", [abi.GEMM_AUTO])
def test_4096_cubed_all_six_share_planes_on_sampled_rows(algo):
    """BASELINE configs[1] size: the input sharings in full and, on rows sampled from every 128-row tile band plus
    the matrix edges, all six share planes of BOTH products (zero-share form and truncating form) against the oracle
    evaluated lazily (tests/sampled_rows.py: row i needs A's row i, B, and the keystreams at element offset i * N).
    A tile-rasterisation or keystream-offset bug anywhere in the 2048 output tiles shows up here as a share mismatch;
    Freivalds above only sees the reconstructed value."""
    import sampled_rows as sr
    n, d = 4096, 16
    s, r = harness.Session(), o.Session()
    try:
        s.set_gemm_algo(algo)
        rng = np.random.default_rng(21)
        a, b = rnd(22, (n, n)), rnd(23, (n, n))
        A, B = s.share_int(0, a), s.share_int(2, b)
        Ao, Bo = r.share_int(0, a), r.share_int(2, b)
        assert np.array_equal(s.get_shares(A), Ao)
        assert np.array_equal(s.get_shares(B), Bo)
        rows = np.unique(np.concatenate([[0, 1, 127, 128, n - 129, n - 128, n - 1],
                                         np.arange(32) * 128 + rng.integers(0, 128, 32)]))
        cur = [s.cursors(p) for p in range(3)]
        assert all(list(cur[p]) == list(r.cursors(p)) for p in range(3))
        C = s.mul(A, B)
        got = s.get_shares(C)
        exp = sr.mul_rows(r, cur, Ao, Bo, rows)
        assert np.array_equal(got[:, :, rows], exp)
        for p in range(3):                                        # the other rows: replicated-share consistency
            assert np.array_equal(got[(p + 1) % 3, 1], got[p, 0])
        s.free(C)
        # the truncating form twice: the second call starts from advanced common-keystream cursors and, with the
        # early truncation pair, from recycled pool blocks
        for it in range(2):
            cur = [s.cursors(p) for p in range(3)]
            T = s.mul(A, B, shift=d)
            got = s.get_shares(T)
            exp = sr.mul_trunc_rows(r, cur, Ao, Bo, d, rows)
            assert np.array_equal(got[:, :, rows], exp), it
            for p in range(3):
                assert np.array_equal(got[(p + 1) % 3, 1], got[p, 0])
            s.free(T)
    finally:
        s.close()
        r.close()


def test_binary_and_layer_2_pow_22_instances():
    """config 5 scale (bitwiseAnd(64) over millions of instances): reveal == a & b, consistency."""
    width = 1 << 22
    s = harness.Session()
    try:
        x, y = rnd(5, (width, 1)), rnd(6, (width, 1))
        X, Y = s.share_bin(0, x, 64), s.share_bin(1, y, 64)
        out = s.bin_eval(harness.library_circuit("and", 64), [X, Y])[0]
        sh = s.get_shares(out, binary=True)
        for p in range(3):
            assert np.array_equal(sh[(p + 1) % 3, 1], sh[p, 0])
        assert np.array_equal(s.reveal(out, 0, binary=True), x & y)
    finally:
        s.close()


@pytest.mark.parametrize("name", ["lt", "add_msb"])
def test_binary_engine_share_level_at_the_wide_aes_size(name):
    """A comparison circuit over 300 000 instances: wire rows of 37 KiB, so the AND layers take the wide AES form
    (four tables, CTR folding over runs of 256 consecutive row chunks).  Every party's output shares against the oracle."""
    width = 300000
    s, r = harness.Session(), o.Session()
    try:
        x, y = rnd(31, (width, 1)), rnd(32, (width, 1))
        cir = harness.library_circuit(name, 64)
        X, Y = s.share_bin(0, x, 64), s.share_bin(2, y, 64)
        Xo, Yo = r.share_bin(0, x), r.share_bin(2, y)
        out = s.bin_eval(cir, [X, Y])[0]
        outo, _ = o.bin_eval(r, cir, width, [Xo, Yo])
        assert np.array_equal(s.get_shares(out, binary=True) & 1, outo[0] & 1)
        for p in range(3):
            assert list(s.cursors(p)) == list(r.cursors(p))
    finally:
        s.close()
        r.close()


def test_empty_inputs_are_no_ops(ctx):
    z = abi.C.c_void_p(None)
    before = ctx.launches
    key = bytes(16)
    abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, key, 0, z, 0))
    abi.check(lib.aby3cu_zero_share(ctx.h, key, key, 0, z, z, 0, 0))
    abi.check(lib.aby3cu_mul_hadamard(ctx.h, z, z, z, z, key, key, 0, z, 0))
    abi.check(lib.aby3cu_trunc_finish(ctx.h, z, z, z, z, 0, 16))
    abi.check(lib.aby3cu_share_op(ctx.h, 0, z, z, z, 0))
    abi.check(lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_AUTO, z, z, z, z, 0, 5, 7, z, 0))
    abi.check(lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_AUTO, z, z, z, z, 5, 5, 0, z, 0))
    abi.check(lib.aby3cu_bit_transpose(ctx.h, z, 0, 0, 8, z, 8, z))
    assert ctx.launches == before
    # K == 0: the product is the zero matrix (or leaves C alone when accumulating)
    c = ctx.upload(np.full(6, 9, dtype=np.int64))
    abi.check(lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_AUTO, z, z, z, z, 2, 0, 3, c.p, 1))
    assert np.array_equal(ctx.download(c, 6), np.full(6, 9))
    abi.check(lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_AUTO, z, z, z, z, 2, 0, 3, c.p, 0))
    assert np.array_equal(ctx.download(c, 6), np.zeros(6, dtype=np.int64))


def test_bad_arguments_fail_loudly(ctx):
    z = abi.C.c_void_p(None)
    with pytest.raises(abi.Aby3CudaError):
        abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, bytes(16), 4, ctx.alloc(64).p, 16))      # offset not a multiple of 8
    with pytest.raises(abi.Aby3CudaError):
        abi.check(lib.aby3cu_gemm_cross(ctx.h, 7, z, z, z, z, 1, 1, 1, z, 0))              # unknown algo
    with pytest.raises(abi.Aby3CudaError):
        abi.check(lib.aby3cu_trunc_finish(ctx.h, z, z, z, z, 4, 64))                       # shift out of range
    with pytest.raises(abi.Aby3CudaError):
        abi.check(lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_AUTO, z, z, z, z, 2, 2, 2, z, 0))  # null operands


def test_k_beyond_the_per_launch_exactness_bound(ctx):
    """K = 8300 > 8256: the host splits K and accumulates through C; worst-case limbs (all 0xFF)."""
    M, K, N = 128, 8300, 64
    a = np.full((M, K), -1, dtype=np.int64)
    b = np.full((K, N), -1, dtype=np.int64)
    zb = np.zeros((K, N), dtype=np.int64)
    d = [ctx.upload(x) for x in (a, a, b, zb)]
    c = ctx.alloc(8 * M * N)
    abi.check(lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_TCGEN05, d[0].p, d[1].p, d[2].p, d[3].p, M, K, N, c.p, 0))
    assert np.array_equal(ctx.download(c, (M, N)), o.cross_term(a, a, b, zb))
    ra, rb, rb1 = rnd(1, (M, K)), rnd(2, (K, N)), rnd(3, (K, N))
    d = [ctx.upload(x) for x in (ra, a, rb, rb1)]
    abi.check(lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_TCGEN05, d[0].p, d[1].p, d[2].p, d[3].p, M, K, N, c.p, 0))
    assert np.array_equal(ctx.download(c, (M, N)), o.cross_term(ra, a, rb, rb1))


def test_forced_row_blocking_of_the_limb_workspace():
    """ABY3CU_WS_LIMIT_MB=1 makes the tcgen05 path loop over several row blocks."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle_lib as o
from aby3_b200 import abi
ctx = abi.Ctx(0)
rng = np.random.default_rng(0)
M, K, N = 1200, 160, 200
a0, a1 = (rng.integers(-2**63, 2**63, (M, K), dtype=np.int64) for _ in range(2))
b0, b1 = (rng.integers(-2**63, 2**63, (K, N), dtype=np.int64) for _ in range(2))
d = [ctx.upload(x) for x in (a0, a1, b0, b1)]
c = ctx.alloc(8 * M * N)
abi.check(abi.lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_TCGEN05, d[0].p, d[1].p, d[2].p, d[3].p, M, K, N, c.p, 0))
assert np.array_equal(ctx.download(c, (M, N)), o.cross_term(a0, a1, b0, b1))
print("ROWBLOCK OK", ctx.launches)
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ABY3CU_WS_LIMIT_MB="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0 and "ROWBLOCK OK" in out.stdout, out.stdout
    launches = int(out.stdout.strip().split()[-1])
    assert launches >= 1 + 2 * 3          # pack_b + several (pack_a, gemm) pairs


def test_linreg_config3_shapes_match_oracle():
    """config 3 shapes (batch 128 x 1024 features): a few SGD_Linear iterations, w shares bit-exact vs the oracle's
    composition of the same products (the full 2^20-sample run differs only in which rows are gathered)."""
    import test_gpu_sh3 as t
    s, r = harness.Session(), o.Session()
    try:
        N, F, B, iters, lr, D = 2048, 1024, 128, 4, 2.0 ** -10, 16
        rng = np.random.default_rng(9)
        x = (rng.normal(1, 1, (N, F)) * (1 << D)).astype(np.int64)
        y = (rng.normal(1, 1, (N, 1)) * (1 << D)).astype(np.int64)
        w = np.zeros((F, 1), dtype=np.int64)
        idx = rng.integers(0, N, iters * B).astype(np.uint64)
        X, Y, W = s.share_int(0, x), s.share_int(0, y), s.share_int(0, w)
        Xo, Yo, Wo = r.share_int(0, x), r.share_int(0, y), r.share_int(0, w)
        s.linreg(X, Y, W, idx, iters, B, lr)
        assert np.array_equal(s.get_shares(W), t.oracle_linreg(r, Xo, Yo, Wo, idx, iters, B, lr, D))
    finally:
        s.close()
        r.close()


def test_logistic_inference_config4_shape_property():
    """config 4 shape (rows x 512 features, piecewise sigmoid): the revealed output is EXACTLY the plaintext piecewise
    function of the revealed linear part (the truncation noise sits in z, f(z) is deterministic given z)."""
    s = harness.Session()
    try:
        rows, F, D = 1 << 17, 512, 16
        rng = np.random.default_rng(10)
        x = (rng.uniform(-1, 1, (rows, F)) * (1 << D)).astype(np.int64)
        w = (rng.uniform(-0.1, 0.1, (F, 1)) * (1 << D)).astype(np.int64)
        X, W = s.share_int(0, x), s.share_int(1, w)
        z = s.mul(X, W, shift=D)
        yv = s.piecewise(z, [-0.5, 0.5], [[], [0.5, 1], [1]], D)
        zr, yr = s.reveal(z, 0), s.reveal(yv, 2)
        half = 1 << (D - 1)
        exp = np.where(zr < -half, 0, np.where(zr < half, zr + half, 1 << D))
        assert np.array_equal(yr, exp)
        assert np.max(np.abs(zr - ((x @ w) >> D))) <= 4
        sh = s.get_shares(yv)
        for p in range(3):
            assert np.array_equal(sh[(p + 1) % 3, 1], sh[p, 0])
    finally:
        s.close()


def test_cipher_gt_and_merge_config5_properties():
    """config 5: cipher_gt over 2^22 pairs equals the plaintext comparison; odd_even_merge of two sorted runs of 2^19
    keys (odd sizes too) returns a sorted permutation of the inputs."""
    s = harness.Session()
    try:
        n = 1 << 22
        rng = np.random.default_rng(11)
        a = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
        b = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
        b[:1000] = a[:1000]                                   # ties
        A, B = s.share_int(0, a), s.share_int(2, b)
        gt = s.cipher_gt(A, B)
        assert np.array_equal(s.reveal(gt, 1, binary=True) & 1, (a > b).astype(np.int64))
        for h in (gt, A, B):
            s.free(h)
        for l1, l2 in (((1 << 19), (1 << 19)), (1000, 777), (1, 1)):
            d1 = np.sort(rng.integers(-2**62, 2**62, l1)).reshape(-1, 1)
            d2 = np.sort(rng.integers(-2**62, 2**62, l2)).reshape(-1, 1)
            D1, D2 = s.share_bin(0, d1, 64), s.share_bin(1, d2, 64)
            m = s.reveal(s.odd_even_merge(D1, D2), 0, binary=True).reshape(-1)
            assert np.array_equal(m, np.sort(np.concatenate([d1[:, 0], d2[:, 0]])))
    finally:
        s.close()
