"""The boundary claim of INTEGRATION.md section A, tested: the reference's callers -- aby3-ML (aby3ML.cpp, Regression.h,
main-linear.cpp, LinearModelGen.cpp) and aby3-Basic (BoolBasic, ArithBasic, BuildingBlocks, Sort, Basic, debug .cpp) --
compile UNMODIFIED, from where they lie under /root/reference, against the B200 facade (compat/include forwards
<aby3/sh3/*.h>, <cryptoTools/...>, <Eigen/Dense>), link against libsh3 + libaby3cu, and give the right answers on the GPU.

CPU part (runs where the reference tree is): the build itself.  GPU part: the prebuilt library travels with the snapshot."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import compat_lib as c
import oracle_lib as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
U64 = np.uint64


@pytest.mark.skipif(not c.have_reference(), reason="needs the reference tree (the GPU box uses the prebuilt library)")
def test_reference_applications_compile_unmodified_against_the_facade():
    out = c.build()
    assert os.path.exists(c.PATH)
    # the objects come from the reference's own files: every compile line names a path under the reference root, and
    # nothing under compat/ or aby3_b200/ carries a copy of those sources
    mk = open(os.path.join(ROOT, "compat", "Makefile")).read()
    assert "vpath %.cpp $(REF)/aby3-ML $(REF)/aby3-Basic $(REF)/aby3_tests" in mk
    for f in c.APP_SOURCES:
        assert os.path.exists(os.path.join(c.REFERENCE, f)), f
        assert not os.path.exists(os.path.join(ROOT, "compat", os.path.basename(f)))
    # exported entry points + no unresolved symbol once libsh3 / libaby3cu are on the path
    r = subprocess.run(["ldd", "-r", c.PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert "undefined symbol" not in r.stdout, r.stdout[-2000:]
    nm = subprocess.run(["nm", "-D", "--defined-only", c.PATH], stdout=subprocess.PIPE, text=True).stdout
    for sym in ("cmp_sgd_linear", "cmp_main_linear", "cmp_basic_bool", "cmp_basic_odd_even_merge",
                "_Z18linear_main_3pc_shRN2oc3CLPE", "_Z14bool_cipher_lt", "cmp_role_test",
                "_Z16arith_basic_testRN2oc3CLPE", "_Z19odd_even_merge_testRN2oc3CLPE"):
        assert sym in nm, sym
    assert out is not None


def ref_batches(N, B, iters):
    """aby3-ML/Regression.h:24-40,127: mini-batches without replacement; pool order from PRNG(toBlock(234543234)) through
    libstdc++'s std::random_shuffle(first, last, rand): for i in 1..n-1: swap(a[i], a[rand(i + 1)]), rand(k) = get<u64>() % k."""
    need = (iters * B // N + 2) * N
    ks = o.keystream(o.to_block(0, 234543234), 0, 8 * need).view(np.uint64)
    pos = 0
    pool = list(range(N))
    it = N
    out = []
    for _ in range(iters):
        dest = []
        while True:
            step = min(N - it, B - len(dest))
            dest += pool[it:it + step]
            it += step
            if it == N:
                for i in range(1, N):
                    j = int(ks[pos] % np.uint64(i + 1))
                    pos += 1
                    pool[i], pool[j] = pool[j], pool[i]
                it = 0
            if len(dest) == B:
                break
        out += dest
    return np.asarray(out, dtype=np.uint64)


def ml_seeds():
    """aby3-ML/aby3ML.cpp:14-16 + Sh3ShareGen.h:25-31: party i draws (enc seed, eval seed) as the first two blocks of
    PRNG(toBlock(i)), sends its seed to the next party and uses the previous party's as prevSeed."""
    own = [o.keystream(o.to_block(0, i), 0, 32) for i in range(3)]
    enc = b"".join(bytes(own[(i + 2) % 3][:16]) + bytes(own[i][:16]) for i in range(3))
    ev = b"".join(bytes(own[(i + 2) % 3][16:]) + bytes(own[i][16:]) for i in range(3))
    return enc, ev


@pytest.mark.gpu
def test_reference_sgd_linear_on_the_facade_matches_the_facade_loop_and_the_oracle():
    """SGD_Linear straight from the reference's Regression.h, driven by its aby3ML engine, every product a facade
    asyncMul on the B200: the w shares after k iterations equal, bit for bit, those of the facade's own SGD_Linear
    (ml/Regression.h, harness.linreg) and of the oracle composition, on the same seeds and the same mini-batches."""
    import test_gpu_sh3 as t
    from aby3_b200 import harness
    N, F, B, iters, lr, D = 300, 24, 16, 45, 2.0 ** -6, 16
    rng = np.random.default_rng(5)
    x = rng.normal(1, 1, (N, F))
    y = x[:, :3] @ np.array([2.0, -1.0, 0.5])
    secs, w = c.sgd_linear(x, y, B, iters, lr)
    assert secs > 0
    for p in range(3):
        assert np.array_equal(w[(p + 1) % 3, 1], w[p, 0])
    enc, ev = ml_seeds()
    idx = ref_batches(N, B, iters)
    fx, fy = (x * (1 << D)).astype(np.int64), (y.reshape(-1, 1) * (1 << D)).astype(np.int64)
    fw = np.zeros((F, 1), dtype=np.int64)
    s, r = harness.Session(enc_seeds=enc, eval_seeds=ev), o.Session(enc, ev)
    try:
        X, Y, W = s.share_int(0, fx), s.share_int(0, fy), s.share_int(0, fw)
        Xo, Yo, Wo = r.share_int(0, fx), r.share_int(0, fy), r.share_int(0, fw)
        s.linreg(X, Y, W, idx, iters, B, lr)
        assert np.array_equal(s.get_shares(W), w), "reference Regression.h on the facade != facade SGD_Linear"
        assert np.array_equal(t.oracle_linreg(r, Xo, Yo, Wo, idx, iters, B, lr, D), w), "!= oracle composition"
        learnt = o.reveal(w).astype(np.float64) / (1 << D)
        assert learnt[0, 0] > 0.5 > learnt[3, 0]            # 45 steps in: moving towards the model (2, -1, 0.5, 0, ...)
    finally:
        s.close()
        r.close()


@pytest.mark.gpu
def test_reference_main_linear_runs_on_the_facade(capfd):
    c.main_linear("-N", 600, "-D", 48, "-B", 16, "-I", 40, "-testN", 50)
    out = capfd.readouterr().out
    assert "iters/s" in out and "IT:40 =>" in out        # (party threads interleave their prints)


@pytest.mark.gpu
def test_reference_aby3_basic_on_the_facade_matches_plaintext_and_the_cpu_reference():
    """aby3-Basic's own functions (BoolBasic.cpp, BuildingBlocks.cpp, Sort.cpp) on the facade: every reveal equals the
    plaintext function; where the circuit is canonical (bitwise and / or: gate i = bit i) and for the arithmetic
    element-wise product, every party's SHARES equal those of the same sources running on the CPU (oracle/_ref)."""
    import ref_lib as rl
    e, v = o.default_seeds()
    s = c.Session(e, v)
    rs = rl.Session(e, v) if rl.available() else None
    try:
        rng = np.random.default_rng(0)
        n = 700
        a = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
        b = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
        b[:20] = a[:20]
        # (both sessions run the same calls in the same order, so their keystream cursors stay in lockstep)
        Aa, Ba = s.share_int(0, a), s.share_int(2, b)
        prod = s.cipher_mul(Aa, Ba)
        assert np.array_equal(s.reveal_all(prod)[0], (a.view(U64) * b.view(U64)).view(np.int64))
        if rs is not None:
            assert np.array_equal(Aa, rs.share_int(0, a)) and np.array_equal(Ba, rs.share_int(2, b))
            assert np.array_equal(prod, rs.mul(Aa, Ba))                    # share level
        A, B = s.share_bin(0, a), s.share_bin(1, b)
        if rs is not None:
            assert np.array_equal(A, rs.share_bin(0, a)) and np.array_equal(B, rs.share_bin(1, b))
        plain = {"and": a & b, "or": a | b, "lt": (a < b).astype(np.int64), "eq": (a == b).astype(np.int64), "add": a + b,
                 "max": np.maximum(a, b), "min": np.minimum(a, b)}
        for op, exp in plain.items():
            out, _ = s.basic_bool(op, A, B)
            rev = s.reveal_all(out, binary=True)
            for p in range(3):
                got = rev[p] & 1 if op in ("lt", "eq") else rev[p]
                assert np.array_equal(got, exp), (op, p)
            if rs is not None:
                ro, _ = rs.basic_bool(op, A, B)
                if op in ("and", "or"):
                    assert np.array_equal(out, ro), op                     # share level
        out, _ = s.cipher_gt(Aa, Ba)
        assert np.array_equal(s.reveal_all(out, binary=True)[1] & 1, (a > b).astype(np.int64))
        mx, mn, _ = s.max_min_split(A, B)
        assert np.array_equal(s.reveal_all(mx, binary=True)[0], np.maximum(a, b))
        assert np.array_equal(s.reveal_all(mn, binary=True)[2], np.minimum(a, b))
        d1 = np.sort(rng.integers(-2**40, 2**40, 50)).reshape(-1, 1)       # SortTest.cpp:363
        d2 = np.sort(rng.integers(-2**40, 2**40, 98)).reshape(-1, 1)
        out, _ = s.odd_even_merge(s.share_bin(0, d1), s.share_bin(0, d2))
        assert np.array_equal(s.reveal_all(out, binary=True)[0].reshape(-1), np.sort(np.concatenate([d1[:, 0], d2[:, 0]])))
    finally:
        s.close()
        if rs is not None:
            rs.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(c.ROLE_TESTS))
def test_reference_role_tests_pass_on_the_facade(name):
    """The fork's own test programs (aby3_tests/Test.cpp, BoolTest.cpp, SortTest.cpp -- what `frontend -Arith`, `-Bool`,
    `-Sort` ... run, frontend/main.cpp:16-54), compiled unmodified against the facade, three party threads with `-role i`:
    every check_result() (aby3-Basic/debug.h) reports SUCCESS, and as many of them as on the reference's CPU build."""
    ok, bad = c.role_test(name)
    assert bad == 0, (name, ok, bad)
    assert ok == c.ROLE_TESTS[name], (name, ok)
