"""GPU parity tests of the sh3 facade (C++ Sh3Encryptor / Sh3Evaluator /
Sh3BinaryEvaluator over the C ABI) against the CPU oracle: same seeds, same
inputs, every party's two share planes compared bit for bit, plus the
reference's own reconstruction-level assertions."""
import numpy as np
import pytest

import oracle_lib as o
from aby3_b200 import abi, harness

pytestmark = pytest.mark.gpu
U64 = np.uint64


@pytest.fixture()
def pair():
    s, r = harness.Session(), o.Session()
    yield s, r
    s.close()
    r.close()


def rnd(seed, shape):
    return np.random.default_rng(seed).integers(-2**63, 2**63, shape, dtype=np.int64)


def assert_cursors(s, r):
    for p in range(3):
        assert list(s.cursors(p)) == list(r.cursors(p)), p


def test_init_cursors(pair):
    s, r = pair
    assert_cursors(s, r)


@pytest.mark.parametrize("shape", [(1, 1), (10, 10), (16, 1), (511, 3), (1, 513), (300, 700)])
def test_share_and_reveal_int(pair, shape):
    s, r = pair
    for owner in range(3):
        m = rnd(owner + shape[0], shape)
        h = s.share_int(owner, m)
        assert np.array_equal(s.get_shares(h), r.share_int(owner, m))
        for p in range(3):
            assert np.array_equal(s.reveal(h, p), m)
    assert_cursors(s, r)


def test_share_and_reveal_bin(pair):
    s, r = pair
    m = rnd(1, (77, 2))
    h = s.share_bin(1, m, 128)
    assert np.array_equal(s.get_shares(h, binary=True), r.share_bin(1, m))
    assert np.array_equal(s.reveal(h, 2, binary=True), m)


@pytest.mark.parametrize("algo", [abi.GEMM_IMAD, abi.GEMM_TCGEN05, abi.GEMM_AUTO])
@pytest.mark.parametrize("dims", [(10, 10, 10), (128, 64, 64), (129, 200, 70), (128, 1024, 1), (1024, 128, 1), (256, 256, 256)])
def test_mul_matmul_shares_bit_exact(pair, algo, dims):
    s, r = pair
    s.set_gemm_algo(algo)
    M, K, N = dims
    a, b = rnd(1, (M, K)), rnd(2, (K, N))
    A, B = s.share_int(0, a), s.share_int(1, b)
    Ao, Bo = r.share_int(0, a), r.share_int(1, b)
    C = s.mul(A, B)
    Co = r.mul(Ao, Bo, mode=0)
    assert np.array_equal(s.get_shares(C), Co)
    assert np.array_equal(s.reveal(C, 0), o.plain_mul(a, b))
    assert_cursors(s, r)


def test_Sh3_Evaluator_asyncMul_test_chain(pair):
    """aby3_tests/Sh3EvaluatorTests.cpp:20-135: ten chained 10x10 products with A = C + A."""
    s, r = pair
    n = 10
    a, b = rnd(3, (n, n)), rnd(4, (n, n))
    A, B = s.share_int(0, a), s.share_int(0, b)
    Ao, Bo = r.share_int(0, a), r.share_int(0, b)
    for _ in range(n):
        C = s.mul(A, B)
        A = s.add(C, A)
        Co = r.mul(Ao, Bo)
        Ao = (Co.view(U64) + Ao.view(U64)).view(np.int64)
        c = o.plain_mul(a, b)
        a = (c.view(U64) + a.view(U64)).view(np.int64)
    assert np.array_equal(s.get_shares(C), Co)
    for p in range(3):
        assert np.array_equal(s.reveal(C, p), c)


@pytest.mark.parametrize("n", [1, 16, 513, 100001])
def test_mul_hadamard_fork_semantics(pair, n):
    """n x 1 operands multiply element-wise (aby3_tests/Test.cpp:116,153,184)."""
    s, r = pair
    a, b = rnd(5, (n, 1)), rnd(6, (n, 1))
    if n == 16:
        a = np.arange(n, dtype=np.int64).reshape(n, 1)
        b = (n - np.arange(n, dtype=np.int64)).reshape(n, 1)
    A, B = s.share_int(0, a), s.share_int(2, b)
    Ao, Bo = r.share_int(0, a), r.share_int(2, b)
    if n == 1:
        return      # 1x1: cols == rows, the matmul rule applies and coincides
    C = s.mul(A, B)
    Co = r.mul(Ao, Bo, mode=1)
    assert np.array_equal(s.get_shares(C), Co)
    assert np.array_equal(s.reveal(C, 1), (a.view(U64) * b.view(U64)).view(np.int64))
    assert_cursors(s, r)


def test_truncation_tuple_matches_oracle_and_bound(pair):
    """Sh3EvaluatorTests.cpp:350-410."""
    s, r = pair
    d = 8
    for _ in range(5):
        ts = [s.trunc_tuple(p, 4, 4, d) for p in range(3)]
        to = [r.trunc_tuple(p, 16, d) for p in range(3)]
        for p in range(3):
            for k in range(3):
                assert np.array_equal(ts[p][k], to[p][k])
        tr = (ts[0][1].view(U64) + ts[1][1].view(U64) + ts[2][1].view(U64)).view(np.int64)
        rr = (ts[0][0].view(U64) + ts[1][0].view(U64) + ts[2][0].view(U64)).view(np.int64)
        assert np.all(np.abs(tr - (rr >> d)) < 4)
    assert_cursors(s, r)


def fixed(vals, d):
    return (vals * (1 << d)).astype(np.int64)


@pytest.mark.parametrize("algo", [abi.GEMM_IMAD, abi.GEMM_TCGEN05])
def test_Sh3_Evaluator_asyncMul_matrixFixed_test(pair, algo):
    """Sh3EvaluatorTests.cpp:413-589: D8, randomisation off, reveal within 1 ulp; shares bit-exact."""
    s, r = pair
    s.set_gemm_algo(algo)
    s.disable_randomization(True)
    r.disable_randomization(True)
    d = 8
    for size in (4, 130):
        rng = np.random.default_rng(size)
        a = fixed((rng.integers(0, 2**32, (size, size), dtype=np.uint64) >> U64(8)).astype(np.float64) / 100.0, d)
        b = fixed((rng.integers(0, 2**32, (size, size), dtype=np.uint64) >> U64(8)).astype(np.float64) / 100.0, d)
        A, B = s.share_int(0, a), s.share_int(0, b)
        Ao, Bo = r.share_int(0, a), r.share_int(0, b)
        C = s.mul(A, B, shift=d)
        Co = r.mul_trunc(Ao, Bo, d)
        assert np.array_equal(s.get_shares(C), Co)
        c = o.plain_mul(a, b) >> d
        for p in range(3):
            assert np.all(np.abs(s.reveal(C, p) - c) <= 1)


@pytest.mark.parametrize("dims,shift", [((9, 6, 5), 16), ((128, 1024, 1), 16), ((1024, 128, 1), 33), ((200, 300, 70), 16)])
def test_mul_trunc_randomised_shares_bit_exact(pair, dims, shift):
    s, r = pair
    M, K, N = dims
    rng = np.random.default_rng(M)
    a, b = fixed(rng.normal(0, 20, (M, K)), 16), fixed(rng.normal(0, 20, (K, N)), 16)
    A, B = s.share_int(0, a), s.share_int(1, b)
    Ao, Bo = r.share_int(0, a), r.share_int(1, b)
    C = s.mul(A, B, shift=shift)
    Co = r.mul_trunc(Ao, Bo, shift)
    assert np.array_equal(s.get_shares(C), Co)
    c = o.plain_mul(a, b) >> shift
    for p in range(3):
        assert np.all(np.abs(s.reveal(C, p) - c) <= 4)
    # a second product continues the PRNG streams where the first stopped
    C2 = s.mul(A, B, shift=shift)
    assert np.array_equal(s.get_shares(C2), r.mul_trunc(Ao, Bo, shift))
    assert_cursors(s, r)


def test_mul_trunc_hadamard(pair):
    s, r = pair
    n = 1000
    a, b = fixed(np.random.default_rng(1).normal(0, 5, (n, 1)), 16), fixed(np.random.default_rng(2).normal(0, 5, (n, 1)), 16)
    A, B = s.share_int(0, a), s.share_int(0, b)
    Ao, Bo = r.share_int(0, a), r.share_int(0, b)
    C = s.mul(A, B, shift=16)
    assert np.array_equal(s.get_shares(C), r.mul_trunc(Ao, Bo, 16, mode=1))


def test_result_may_alias_an_operand(pair):
    s, r = pair
    a, b = rnd(1, (40, 40)), rnd(2, (40, 40))
    A, B = s.share_int(0, a), s.share_int(0, b)
    Ao, Bo = r.share_int(0, a), r.share_int(0, b)
    s.mul(A, B, out=A)
    assert np.array_equal(s.get_shares(A), r.mul(Ao, Bo))


def test_bad_shapes_throw(pair):
    s, _ = pair
    A, B = s.share_int(0, rnd(1, (4, 5))), s.share_int(0, rnd(2, (4, 6)))
    with pytest.raises(harness.Sh3Error):
        s.mul(A, B)


BIN_CASES = [("and", 8, 256), ("and", 64, 5000), ("or", 64, 300), ("nor", 64, 333), ("xor", 33, 65), ("add", 8, 256), ("add_depth", 64, 3000),
             ("add_msb", 64, 2049), ("lt", 64, 4097), ("eq", 64, 100)]


@pytest.mark.parametrize("name,bits,width", BIN_CASES)
def test_binary_engine_matches_oracle(pair, name, bits, width):
    """Sh3BinaryEvaluatorTests.cpp:333-424 semantics + share-level parity with the oracle."""
    s, r = pair
    rng = np.random.default_rng(width)
    m = U64((1 << bits) - 1) if bits < 64 else U64(2**64 - 1)
    a = (rng.integers(0, 2**63, width, dtype=np.uint64) * U64(2) + rng.integers(0, 2, width, dtype=np.uint64)) & m
    b = (rng.integers(0, 2**63, width, dtype=np.uint64) * U64(2) + rng.integers(0, 2, width, dtype=np.uint64)) & m
    b[: width // 4] = a[: width // 4]
    cir = harness.library_circuit(name, bits)
    A = s.share_bin(0, a.view(np.int64).reshape(width, 1), bits)
    B = s.share_bin(1, b.view(np.int64).reshape(width, 1), bits)
    Ao = r.share_bin(0, a.view(np.int64).reshape(width, 1))
    Bo = r.share_bin(1, b.view(np.int64).reshape(width, 1))
    outs = s.bin_eval(cir, [A, B])
    outs_o, _ = o.bin_eval(r, cir, width, [Ao, Bo])
    obits = int(cir["output_bits"][0])
    om = np.int64(-1) if obits == 64 else np.int64((1 << obits) - 1)
    got = s.get_shares(outs[0], binary=True)
    assert np.array_equal(got & om, outs_o[0] & om)
    rev = s.reveal(outs[0], 0, binary=True) & om
    assert np.array_equal(rev, o.reveal(outs_o[0], 0, binary=True) & om)
    assert_cursors(s, r)


def oracle_linreg(r, X, Y, w, idx, iters, B, lr, D=16):
    """aby3-ML/Regression.h:142-171 on the oracle's share arrays."""
    import math
    aB = int(math.log2(1 / (lr / B)))
    for i in range(iters):
        bi = idx[i * B:(i + 1) * B].astype(np.int64)
        XX = np.ascontiguousarray(X[:, :, bi, :])
        YY = np.ascontiguousarray(Y[:, :, bi, :])
        err = r.mul_trunc(XX, w, D)
        err = (err.view(U64) - YY.view(U64)).view(np.int64)
        XXt = np.ascontiguousarray(np.swapaxes(XX, 2, 3))
        upd = r.mul_trunc(XXt, np.ascontiguousarray(err), D + aB)
        w = np.ascontiguousarray((w.view(U64) - upd.view(U64)).view(np.int64))
    return w


def test_linear_regression_sgd_matches_oracle(pair):
    """config 3 in miniature: SGD_Linear (aby3-ML/Regression.h:112-184) on sf64<D16>,
    w shares after k iterations bit-exact against the oracle, and the model is learnt."""
    s, r = pair
    N, F, B, iters, lr, D = 512, 24, 16, 120, 2.0 ** -6, 16
    rng = np.random.default_rng(5)
    model = np.zeros((F, 1)); model[:5, 0] = [3, -2, 1, 4, -1]
    x = rng.normal(1, 1, (N, F))
    y = x @ model + rng.normal(0, 0.01, (N, 1))
    fx, fy, fw = fixed(x, D), fixed(y, D), np.zeros((F, 1), dtype=np.int64)
    idx = np.concatenate([rng.permutation(N) for _ in range((iters * B + N - 1) // N)])[:iters * B].astype(np.uint64)
    X, Y, W = s.share_int(0, fx), s.share_int(0, fy), s.share_int(0, fw)
    Xo, Yo, Wo = r.share_int(0, fx), r.share_int(0, fy), r.share_int(0, fw)
    s.linreg(X, Y, W, idx, iters, B, lr)
    Wo = oracle_linreg(r, Xo, Yo, Wo, idx, iters, B, lr, D)
    assert np.array_equal(s.get_shares(W), Wo)
    learnt = s.reveal(W, 0).astype(np.float64) / (1 << D)
    assert np.linalg.norm(learnt - model) < 0.5 * np.linalg.norm(model)
    assert_cursors(s, r)


@pytest.mark.parametrize("name,bits,width", [("and", 64, 777), ("add_depth", 64, 4100), ("lt", 64, 64)])
def test_binary_engine_packed_io_matches_oracle(pair, name, bits, width):
    """setInput(i, sPackedBin) / getOutput(i, sPackedBin) (Sh3BinaryEvaluator.cpp:279-309, 1213-1283)
    give the same shares as the sbMatrix forms."""
    s, r = pair
    rng = np.random.default_rng(width)
    a, b = rng.integers(-2**63, 2**63, (width, 1), dtype=np.int64), rng.integers(-2**63, 2**63, (width, 1), dtype=np.int64)
    cir = harness.library_circuit(name, bits)
    A, B = s.share_bin(0, a, bits), s.share_bin(1, b, bits)
    Ao, Bo = r.share_bin(0, a), r.share_bin(1, b)
    outs = s.bin_eval(cir, [A, B], packed=True)
    outs_o, _ = o.bin_eval(r, cir, width, [Ao, Bo])
    obits = int(cir["output_bits"][0])
    om = np.int64(-1) if obits == 64 else np.int64((1 << obits) - 1)
    assert np.array_equal(s.get_shares(outs[0], binary=True) & om, outs_o[0] & om)
    assert_cursors(s, r)


@pytest.mark.parametrize("n", [1, 7, 1000, 70001])
def test_arith_times_shared_bit_matches_oracle(pair, n):
    """sh3_asyncArithBinMul_test / asyncPubArithBinMul (Sh3EvaluatorTests.cpp:780-1032): c = b*a
    exactly; every share plane and PRNG / OT cursor position equal to the oracle's."""
    s, r = pair
    rng = np.random.default_rng(n)
    a = rng.integers(-2**63, 2**63, (n, 1), dtype=np.int64)
    b = rng.integers(0, 2, (n, 1)).astype(np.int64)
    Bo = r.share_bin(0, b) & 1                       # one-bit sbMatrix: only bit 0 carries data
    _ = s.share_bin(0, b, 1)                          # keep the encryptor cursors of both sides aligned
    B = s.set_shares(Bo, binary=True, bit_count=1)
    A, Ao = s.share_int(1, a), r.share_int(1, a)
    for _ in range(2):                                # twice: the OT counters and PRNG cursors carry over
        C = s.mul_bit(A, B)
        Co = r.mul_bit(Ao, Bo)
        assert np.array_equal(s.get_shares(C), Co)
        assert np.array_equal(s.reveal(C, 0), a * b)
    P = s.mul_bit_pub(-12345, B)
    Po = r.mul_bit_pub(-12345, Bo)
    assert np.array_equal(s.get_shares(P), Po)
    assert np.array_equal(s.reveal(P, 2), -12345 * b)
    assert_cursors(s, r)


def test_piecewise_logistic_matches_oracle(pair):
    """aby3ML::logisticFunc (aby3ML.h:121-139) through Sh3Piecewise on the device: every share plane
    equal to the oracle's, reveal equal to the plaintext piecewise function."""
    import piecewise_ref as pw
    s, r = pair
    D, n = 16, 5000
    rng = np.random.default_rng(2)
    x = (rng.uniform(-2, 2, (n, 1)) * (1 << D)).astype(np.int64)
    th, coef = [-0.5, 0.5], [[], [0.5, 1], [1]]
    X, Xo = s.share_int(0, x), r.share_int(0, x)
    out = s.piecewise(X, th, coef, D)
    outo = pw.shared(r, Xo, th, coef, D, harness.library_circuit("piecewise2", 64))
    assert np.array_equal(s.get_shares(out), outo)
    assert np.array_equal(s.reveal(out, 1), pw.plain(x, th, coef, D))
    assert_cursors(s, r)


def test_basic_blocks_match_oracle(pair):
    """aby3-Basic on the device (basic/Basics.h): cipher_gt, bool_cipher_max_min_split and
    odd_even_merge (50 + 98 elements, SortTest.cpp:363) -- shares bit-exact against the same
    compositions on the oracle, reveals against plaintext."""
    import basic_ref as br
    s, r = pair
    rng = np.random.default_rng(0)
    n = 1000
    a, b = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64), rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
    A, B = s.share_int(0, a), s.share_int(1, b)
    Ao, Bo = r.share_int(0, a), r.share_int(1, b)
    gt = s.cipher_gt(A, B)
    gto = br.cipher_gt(r, Ao, Bo)
    assert np.array_equal(s.get_shares(gt, binary=True) & 1, gto & 1)
    assert np.array_equal(s.reveal(gt, 0, binary=True) & 1, (a > b).astype(np.int64))
    X, Y = s.share_bin(0, a, 64), s.share_bin(2, b, 64)
    Xo, Yo = r.share_bin(0, a), r.share_bin(2, b)
    mx, mn = s.max_min_split(X, Y)
    mxo, mno = br.max_min_split(r, Xo, Yo)
    assert np.array_equal(s.get_shares(mx, binary=True), mxo)
    assert np.array_equal(s.get_shares(mn, binary=True), mno)
    assert np.array_equal(s.reveal(mx, 1, binary=True), np.maximum(a, b))
    d1 = np.sort(rng.integers(-2**40, 2**40, 50)).reshape(-1, 1).astype(np.int64)
    d2 = np.sort(rng.integers(-2**40, 2**40, 98)).reshape(-1, 1).astype(np.int64)
    D1, D2 = s.share_bin(0, d1, 64), s.share_bin(0, d2, 64)
    D1o, D2o = r.share_bin(0, d1), r.share_bin(0, d2)
    m = s.odd_even_merge(D1, D2)
    mo = br.odd_even_merge(r, D1o, D2o)
    assert np.array_equal(s.get_shares(m, binary=True), mo)
    got = s.reveal(m, 0, binary=True).reshape(-1)
    assert np.array_equal(got, np.sort(np.concatenate([d1, d2]).reshape(-1)))
    assert_cursors(s, r)


def test_shared_stream_session_matches_oracle():
    """Co-located parties on ONE stream (no event per message): truncating matmul, SGD and a binary
    circuit give the same share planes as the oracle."""
    s, r = harness.Session(transport="shared_stream"), o.Session()
    try:
        rng = np.random.default_rng(21)
        a = (rng.normal(0, 20, (130, 70)) * (1 << 16)).astype(np.int64)
        b = (rng.normal(0, 20, (70, 50)) * (1 << 16)).astype(np.int64)
        A, B = s.share_int(0, a), s.share_int(1, b)
        Ao, Bo = r.share_int(0, a), r.share_int(1, b)
        for _ in range(3):
            Cs = s.mul(A, B, shift=16)
            assert np.array_equal(s.get_shares(Cs), r.mul_trunc(Ao, Bo, 16))
        assert np.array_equal(s.get_shares(s.mul(A, B)), r.mul(Ao, Bo))
        N, F, Bt, iters, lr, D = 256, 16, 8, 40, 2.0 ** -6, 16
        x = rng.normal(1, 1, (N, F))
        y = x[:, :1] * 2.0
        fx, fy, fw = fixed(x, D), fixed(y, D), np.zeros((F, 1), dtype=np.int64)
        idx = rng.integers(0, N, iters * Bt).astype(np.uint64)
        X, Y, W = s.share_int(0, fx), s.share_int(0, fy), s.share_int(0, fw)
        Xo, Yo, Wo = r.share_int(0, fx), r.share_int(0, fy), r.share_int(0, fw)
        s.linreg(X, Y, W, idx, iters, Bt, lr)
        assert np.array_equal(s.get_shares(W), oracle_linreg(r, Xo, Yo, Wo, idx, iters, Bt, lr, D))
        cir = harness.library_circuit("lt", 64)
        u, v = rnd(22, (700, 1)), rnd(23, (700, 1))
        U, V = s.share_bin(0, u, 64), s.share_bin(2, v, 64)
        Uo, Vo = r.share_bin(0, u), r.share_bin(2, v)
        out = s.bin_eval(cir, [U, V])[0]
        outo, _ = o.bin_eval(r, cir, 700, [Uo, Vo])
        assert np.array_equal(s.get_shares(out, binary=True) & 1, outo[0] & 1)
        assert_cursors(s, r)
    finally:
        s.close()
        r.close()


def test_converter_matches_oracle(pair):
    """Sh3Converter on the device (toBinaryMatrix(si64Matrix), bitInjection, toPackedBin / toBinaryMatrix(sPackedBin))
    against the oracle restatement, which tests/test_ref_parity.py pins against the reference's own code."""
    s, r = pair
    s.conv_init()
    r.conv_init()
    rng = np.random.default_rng(31)
    for rows, cols in ((43, 2), (1, 1), (1000, 1)):
        x = rnd(32 + rows, (rows, cols))
        X, Xo = s.share_int(0, x), r.share_int(0, x)
        Y = s.conv_a2b(X)
        Yo = r.conv_a2b(Xo, harness.library_circuit("a2b", 64 * cols))
        assert np.array_equal(s.get_shares(Y, binary=True), Yo)
        assert np.array_equal(s.reveal(Y, 2, binary=True), x)
    for rows, bits in ((43, 17), (5, 64), (20, 91), (3000, 1)):
        words = (bits + 63) // 64
        x = rnd(40 + bits, (rows, words))
        if bits % 64:
            x[:, -1] &= (1 << (bits % 64)) - 1
        B, Bo = s.share_bin(1, x, bits), r.share_bin(1, x)
        Y = s.conv_bit_injection(B)
        Yo = r.conv_bit_injection(Bo, bits)
        assert np.array_equal(s.get_shares(Y), Yo)
        exp = np.stack([(x[:, j // 64] >> (j % 64)) & 1 for j in range(bits)], axis=1)
        assert np.array_equal(s.reveal(Y, 0), exp)
    assert_cursors(s, r)
    for rows, bits in ((1, 1), (65, 63), (200, 130), (5000, 64)):
        words = (bits + 63) // 64
        x = rnd(50 + rows, (rows, words))
        if bits % 64:
            x[:, -1] &= (1 << (bits % 64)) - 1
        B = s.share_bin(0, x, bits)
        sh = s.get_shares(B, binary=True)
        if bits % 64:
            sh[..., -1] &= (1 << (bits % 64)) - 1        # only bitCount bits travel through the packed form
        back, packed = s.conv_packed_roundtrip(B, bits)
        assert np.array_equal(s.get_shares(back, binary=True), sh)
        simd = (rows + 63) // 64
        for p in range(3):
            for pl in range(2):
                t = o.bit_transpose(sh[p, pl].view(np.uint8).reshape(-1), rows, bits, words * 8, simd * 8)
                assert np.array_equal(t.view(np.int64).reshape(bits, simd), packed[p, pl])


def test_converter_two_round_bit_injection(pair):
    """bitInjection(twoRounds = true): party 1 receives its copy from party 0 instead of a second OT"""
    s, _ = pair
    x = rnd(60, (100, 1)) & 0x1FFFF
    B = s.share_bin(2, x, 17)
    Y = s.conv_bit_injection(B, two_rounds=True)
    exp = np.stack([(x[:, 0] >> j) & 1 for j in range(17)], axis=1)
    sh = s.get_shares(Y)
    for p in range(3):
        assert np.array_equal(s.reveal(Y, p), exp)
        assert np.array_equal(sh[(p + 1) % 3, 1], sh[p, 0])


@pytest.mark.parametrize("use_task", [False, True])
def test_packed_binary_sharing_matches_oracle(pair, use_task):
    """localPackedBinary / remotePackedBinary / revealAll(sPackedBin) on the device vs the oracle (pinned against
    the reference in tests/test_ref_parity.py)"""
    s, r = pair
    for owner, (rows, cols) in ((0, (1, 1)), (1, (65, 1)), (2, (200, 2)), (0, (5000, 1))):
        x = rnd(70 + rows, (rows, cols))
        hid, sh = s.share_packed(owner, x, use_task=use_task)
        assert np.array_equal(sh, r.share_packed(owner, x))
        for p in range(3):
            assert np.array_equal(s.reveal_packed(hid, p, rows, cols), x)
    assert_cursors(s, r)


def oracle_logreg(r, X, Y, w, idx, iters, B, lr, D=16):
    """aby3-ML/Regression.h:218-295 on the oracle's share arrays (piecewise sigmoid between the two products)."""
    import math
    import piecewise_ref as pw
    aB = int(math.log2(1 / (lr / B)))
    cir = harness.library_circuit("piecewise2", 64)
    for i in range(iters):
        bi = idx[i * B:(i + 1) * B].astype(np.int64)
        XX = np.ascontiguousarray(X[:, :, bi, :])
        YY = np.ascontiguousarray(Y[:, :, bi, :])
        xw = r.mul_trunc(XX, w, D)
        fxw = pw.shared(r, np.ascontiguousarray(xw), [-0.5, 0.5], [[], [0.5, 1], [1]], D, cir)
        err = (fxw.view(U64) - YY.view(U64)).view(np.int64)
        XXt = np.ascontiguousarray(np.swapaxes(XX, 2, 3))
        upd = r.mul_trunc(XXt, np.ascontiguousarray(err), D + aB)
        w = np.ascontiguousarray((w.view(U64) - upd.view(U64)).view(np.int64))
    return w


def test_logistic_regression_sgd_matches_oracle(pair):
    """aby3-ML SGD_Logistic on sf64<D16>: w shares after k iterations bit-exact against the oracle-side composition,
    and the learnt model separates the classes."""
    s, r = pair
    N, F, B, iters, lr, D = 512, 8, 32, 60, 2.0 ** -2, 16
    rng = np.random.default_rng(41)
    model = np.array([[1.5], [-2.0], [0.5], [0.0], [1.0], [-1.0], [0.0], [0.25]])
    x = rng.normal(0, 1, (N, F))
    y = ((x @ model) > 0).astype(np.float64)
    fx, fy, fw = fixed(x, D), fixed(y, D), np.zeros((F, 1), dtype=np.int64)
    idx = rng.integers(0, N, iters * B).astype(np.uint64)
    X, Y, W = s.share_int(0, fx), s.share_int(0, fy), s.share_int(0, fw)
    Xo, Yo, Wo = r.share_int(0, fx), r.share_int(0, fy), r.share_int(0, fw)
    s.logreg(X, Y, W, idx, iters, B, lr)
    Wo = oracle_logreg(r, Xo, Yo, Wo, idx, iters, B, lr, D)
    assert np.array_equal(s.get_shares(W), Wo)
    learnt = s.reveal(W, 0).astype(np.float64) / (1 << D)
    acc = np.mean(((x @ learnt) > 0).astype(np.float64) == y)
    assert acc > 0.85, acc
    assert_cursors(s, r)


def test_graph_sgd_matches_facade_and_oracle(pair):
    """ml/SgdGraph.h: SGD_Linear replayed as one CUDA graph per iteration gives the same w shares and PRNG cursors
    as the facade loop (three party threads) and as the oracle; a second call continues from the advanced cursors."""
    s, r = pair
    N, F, B, iters, lr, D = 700, 40, 16, 37, 2.0 ** -6, 16
    rng = np.random.default_rng(51)
    x = rng.normal(1, 1, (N, F))
    y = x[:, :3] @ np.array([[2.0], [-1.0], [0.5]])
    fx, fy, fw = fixed(x, D), fixed(y, D), np.zeros((F, 1), dtype=np.int64)
    idx = rng.integers(0, N, 2 * iters * B).astype(np.uint64)
    X, Y, W = s.share_int(0, fx), s.share_int(0, fy), s.share_int(0, fw)
    Xo, Yo, Wo = r.share_int(0, fx), r.share_int(0, fy), r.share_int(0, fw)
    s.linreg_graph(X, Y, W, idx[:iters * B], iters, B, lr)
    Wo = oracle_linreg(r, Xo, Yo, Wo, idx[:iters * B], iters, B, lr, D)
    assert np.array_equal(s.get_shares(W), Wo)
    assert_cursors(s, r)
    # continue: one more stretch through the graph, then one through the facade loop
    s.linreg_graph(X, Y, W, idx[iters * B:iters * B + 5 * B], 5, B, lr)
    Wo = oracle_linreg(r, Xo, Yo, Wo, idx[iters * B:iters * B + 5 * B], 5, B, lr, D)
    assert np.array_equal(s.get_shares(W), Wo)
    s.linreg(X, Y, W, idx[-3 * B:], 3, B, lr)
    Wo = oracle_linreg(r, Xo, Yo, Wo, idx[-3 * B:], 3, B, lr, D)
    assert np.array_equal(s.get_shares(W), Wo)
    s.linreg_graph(X, Y, W, idx[:B], 1, B, lr)                     # a single iteration: no graph at all
    Wo = oracle_linreg(r, Xo, Yo, Wo, idx[:B], 1, B, lr, D)
    assert np.array_equal(s.get_shares(W), Wo)
    assert_cursors(s, r)


@pytest.mark.parametrize("N,F,B,iters", [(700, 40, 16, 37), (3000, 1024, 128, 9), (500, 6, 200, 5), (64, 258, 3, 4), (2000, 1024, 130, 3), (900, 1184, 77, 6),
                                           (50, 8, 1, 3), (300, 1188, 128, 2), (4096, 4, 128, 1)])
def test_fused_sgd_matches_facade_and_oracle(pair, N, F, B, iters):
    """csrc/sgd_fused.cu: SGD_Linear as ONE persistent kernel gives the same w shares and PRNG cursors as the oracle; further
    stretches through the graph replay and the facade loop continue from there.  B <= 128 and F <= 8 * SMs run the
    feature-resident kernel (one grid barrier per iteration), the other shapes the row / slab kernel (two barriers)."""
    s, r = pair
    lr, D = 2.0 ** -6, 16
    rng = np.random.default_rng(52 + F)
    x = rng.normal(1, 1, (N, F))
    y = x[:, :3] @ np.array([[2.0], [-1.0], [0.5]])
    fx, fy, fw = fixed(x, D), fixed(y, D), fixed(rng.normal(0, 1, (F, 1)), D)
    idx = rng.integers(0, N, (iters + 5) * B).astype(np.uint64)
    X, Y, W = s.share_int(0, fx), s.share_int(1, fy), s.share_int(2, fw)
    Xo, Yo, Wo = r.share_int(0, fx), r.share_int(1, fy), r.share_int(2, fw)
    launches = s.launches
    s.linreg_fused(X, Y, W, idx[:iters * B], iters, B, lr)
    assert s.launches - launches == 1                                  # the whole run is one kernel
    Wo = oracle_linreg(r, Xo, Yo, Wo, idx[:iters * B], iters, B, lr, D)
    assert np.array_equal(s.get_shares(W), Wo)
    assert_cursors(s, r)
    s.linreg_graph(X, Y, W, idx[iters * B:(iters + 2) * B], 2, B, lr)
    s.linreg_fused(X, Y, W, idx[(iters + 2) * B:(iters + 4) * B], 2, B, lr)
    s.linreg(X, Y, W, idx[(iters + 4) * B:], 1, B, lr)
    Wo = oracle_linreg(r, Xo, Yo, Wo, idx[iters * B:], 5, B, lr, D)
    assert np.array_equal(s.get_shares(W), Wo)
    assert_cursors(s, r)


@pytest.mark.parametrize("seed,width", [(0, 1), (1, 64), (2, 65), (3, 300), (4, 2049), (5, 77), (6, 100000), (7, 31)])
def test_random_circuits_match_oracle(pair, seed, width):
    """Random circuits over every supported gate type with inverted outputs, ragged widths: device == oracle share for
    share (the oracle is pinned against the reference's evaluator on the same circuits in tests/test_ref_parity.py)."""
    import circuits_random as cr
    s, r = pair
    cir = cr.random_circuit(seed, n_gates=100 + 20 * seed)
    rng = np.random.default_rng(seed)
    ins = [rng.integers(0, 2 ** int(b), width, dtype=np.uint64) for b in cir["input_bits"]]
    S = [s.share_bin(k % 3, x.view(np.int64).reshape(width, 1), int(cir["input_bits"][k])) for k, x in enumerate(ins)]
    So = [r.share_bin(k % 3, x.view(np.int64).reshape(width, 1)) for k, x in enumerate(ins)]
    outs = s.bin_eval(cir, S)
    oo, _ = o.bin_eval(r, cir, width, So)
    exp = cr.plain_eval(cir, ins) if width <= 5000 else None
    for k in range(len(oo)):
        m = np.int64((1 << int(cir["output_bits"][k])) - 1)
        assert np.array_equal(s.get_shares(outs[k], binary=True) & m, oo[k] & m), k
        if exp is not None:
            assert np.array_equal(s.reveal(outs[k], 1, binary=True).reshape(width).view(np.uint64) & np.uint64(m), exp[k]), k
    assert_cursors(s, r)


def test_random_shapes_products_match_oracle(pair):
    """Ragged shapes through every product form (matrix / element-wise, with and without truncation, all GEMM paths
    as AUTO picks them): shares bit-exact against the oracle."""
    s, r = pair
    rng = np.random.default_rng(77)
    for _ in range(12):
        M, K, N = (int(x) for x in rng.integers(1, 200, 3))
        a, b = fixed(rng.normal(0, 30, (M, K)), 16), fixed(rng.normal(0, 30, (K, N)), 16)
        A, B = s.share_int(int(rng.integers(0, 3)), a), s.share_int(int(rng.integers(0, 3)), b)
        Ao, Bo = r.share_int(0, a), r.share_int(0, b)
        Ao, Bo = s.get_shares(A), s.get_shares(B)            # same starting shares on both sides
        if M == K and K == N:
            continue
        assert np.array_equal(s.get_shares(s.mul(A, B)), r.mul(Ao, Bo))
        assert np.array_equal(s.get_shares(s.mul(A, B, shift=16)), r.mul_trunc(Ao, Bo, 16))
    for n in (1, 3, 1000):
        a, b = fixed(rng.normal(0, 30, (n, 1)), 16), fixed(rng.normal(0, 30, (n, 1)), 16)
        A, B = s.share_int(0, a), s.share_int(1, b)
        Ao, Bo = s.get_shares(A), s.get_shares(B)
        r.share_int(0, a); r.share_int(1, b)                 # keep the encryptor cursors in step
        if n == 1:
            continue                                         # 1 x 1 is a (1 x 1) matrix product on both sides
        assert np.array_equal(s.get_shares(s.mul(A, B)), r.mul(Ao, Bo, mode=1))
        assert np.array_equal(s.get_shares(s.mul(A, B, shift=16)), r.mul_trunc(Ao, Bo, 16, mode=1))


def test_shadow_evaluator_finds_injected_faults(pair):
    """The device-side counterpart of the reference's BINARY_ENGINE_DEBUG checker: a clean evaluation has no disagreeing
    gate; a single flipped bit in one party's share of an internal wire is reported (by the gate that wrote the wire and
    by every gate that reads it), whichever party holds the fault."""
    import circuits_random as cr
    s, _ = pair
    for name in ("lt", "add_depth"):
        cir = harness.library_circuit(name, 64)
        x, y = rnd(80, (777, 1)), rnd(81, (777, 1))
        X, Y = s.share_bin(0, x, 64), s.share_bin(1, y, 64)
        assert s.bin_eval_check(cir, [X, Y]) == 0
        g = cir["gates"].reshape(-1, 4)
        and_out = int(g[g[:, 3] == 8][5, 2])                       # output wire of some AND gate
        readers = int(np.sum((g[:, 0] == and_out) | ((g[:, 1] == and_out) & (g[:, 3] != 10))))
        for party in range(3):
            bad = s.bin_eval_check(cir, [X, Y], tamper_wire=and_out, tamper_party=party)
            # the faulty share is seen by its holder and by the previous party (who fetches it as its third plane); the
            # next party reads its own stale copy.  The writing gate always disagrees, a reader only if the fault propagates.
            assert 2 <= bad <= 2 * (1 + readers), (name, party, bad, readers)
    rc = cr.random_circuit(3, n_gates=200)
    ins = [np.random.default_rng(5).integers(0, 2 ** int(b), 500, dtype=np.uint64) for b in rc["input_bits"]]
    S = [s.share_bin(k % 3, v.view(np.int64).reshape(500, 1), int(rc["input_bits"][k])) for k, v in enumerate(ins)]
    assert s.bin_eval_check(rc, S) == 0


def test_overlapped_transfers_give_the_same_product(pair):
    """eMatrix::prefetchDevice / fetchHostAsync (copy-stream h2d / d2h used by bench.py's streamed end-to-end path):
    inputs prefetched as row blocks, results revealed asynchronously -- shares equal the oracle's, reveals the product."""
    s, r = pair
    rng = np.random.default_rng(61)
    M, K, N, NB, D = 256, 96, 80, 4, 16
    a, b = fixed(rng.normal(0, 20, (M, K)), D), fixed(rng.normal(0, 20, (K, N)), D)
    pb, vb = s.plain(0, K, N)
    vb[...] = b
    rb = M // NB
    pa, pc = [], []
    for i in range(NB):
        pid, v = s.plain(0, rb, K)
        v[...] = a[i * rb:(i + 1) * rb]
        pa.append(pid)
        pc.append(s.plain(0, rb, N))
    for rep in range(2):                                   # the second round re-uploads into the same device mirrors
        s.plain_touch(0, pb)
        s.plain_prefetch(0, pb)
        for pid in pa:
            s.plain_touch(0, pid)
            s.plain_prefetch(0, pid)
        hb = s.share_plain(0, pb, K, N)
        Bo = r.share_int(0, b)
        for i in range(NB):
            ha = s.share_plain(0, pa[i], rb, K)
            Ao = r.share_int(0, a[i * rb:(i + 1) * rb])
            assert np.array_equal(s.get_shares(ha), Ao)
            hc = s.mul(ha, hb, shift=D)
            assert np.array_equal(s.get_shares(hc), r.mul_trunc(Ao, Bo, D))
            s.reveal_plain_async(hc, 0, pc[i][0])
        for i in range(NB):
            s.plain_wait(0, pc[i][0])
            ref = (a[i * rb:(i + 1) * rb] @ b) >> D
            assert np.max(np.abs(pc[i][1] - ref)) <= 4
    assert_cursors(s, r)


def test_converter_ragged_bit_count_and_trim(pair):
    """toBinaryMatrix into a destination of 91 bits per row (Sh3_convert_arithToBinaryMatrix_test: n = 43, m = 91): the
    last word is masked on the device (aby3cu_mask_last_word); the revealed low 91 bits are the input's."""
    s, _ = pair
    rows, bits = 43, 91
    x = rnd(95, (rows, 2))
    x[:, 1] &= (1 << (bits - 64)) - 1                      # aby3::details::trim(x, m)
    X = s.share_int(0, x)
    Y = s.conv_a2b(X, bits=bits)
    got = s.reveal(Y, 1, binary=True)
    got[:, 1] &= (1 << (bits - 64)) - 1
    assert np.array_equal(got, x)


def test_early_truncation_pair_stream_matches_oracle(pair):
    """Products of 4 MiB and more draw their truncation pair AHEAD on the party's second stream, into pool blocks that
    are free early (sh3/Gpu.h allocEarly, aby3cu_gemm_cross_after).  A chain of such products with the output reused,
    freed and aliased -- the pool rotates through every case -- gives the oracle's shares, reveals and PRNG cursors."""
    s, r = pair
    rng = np.random.default_rng(71)
    M, K, N, D = 768, 48, 768, 16                           # M * N * 8 = 4.5 MiB >= Context::kEarlyMin
    a, b = fixed(rng.normal(0, 20, (M, K)), D), fixed(rng.normal(0, 20, (K, N)), D)
    q = fixed(rng.normal(0, 1, (N, N)), D)
    A, B, Q = s.share_int(0, a), s.share_int(1, b), s.share_int(2, q)
    Ao, Bo, Qo = r.share_int(0, a), r.share_int(1, b), r.share_int(2, q)
    C = s.mul(A, B, shift=D)
    Co = r.mul_trunc(Ao, Bo, D)
    assert np.array_equal(s.get_shares(C), Co)
    for rep in range(4):                                    # same output handle: its old planes go back to the pool
        s.mul(A, B, shift=D, out=C)
        Co = r.mul_trunc(Ao, Bo, D)
        assert np.array_equal(s.get_shares(C), Co), rep
    E = s.mul(C, Q, shift=D)                                # a product of a product (M x N) * (N x N)
    Eo = r.mul_trunc(Co, Qo, D)
    assert np.array_equal(s.get_shares(E), Eo)
    s.mul(E, Q, shift=D, out=E)                             # output aliases an operand
    Eo = r.mul_trunc(Eo, Qo, D)
    assert np.array_equal(s.get_shares(E), Eo)
    s.free(C)
    F = s.mul(A, B, shift=D)                                # after a free: blocks parked with their release position
    assert np.array_equal(s.get_shares(F), r.mul_trunc(Ao, Bo, D))
    ref = (a @ b) >> D
    assert np.max(np.abs(s.reveal(F, 0) - ref)) <= 4
    assert_cursors(s, r)


@pytest.mark.parametrize("blocks,dims", [(4, (1024, 96, 200)), (3, (700, 64, 130)), (8, (256, 64, 64)), (4, (130, 64, 64))])
def test_block_wise_open_of_the_truncating_product_matches_oracle(pair, blocks, dims):
    """Sh3Evaluator::mOpenBlocks > 1 (the distributed placement's overlap of reshare and contraction): the opened xy - r
    travels in row blocks on the parties' communication streams, each as soon as aby3cu_gemm_cross_blocks has made its
    rows final.  Same shares and cursors as the one-message form, ragged last block and tiny shapes (no blocking) included."""
    s, r = pair
    s.set_gemm_algo(abi.GEMM_TCGEN05)
    s.set_open_blocks(blocks)
    M, K, N = dims
    rng = np.random.default_rng(M)
    a = (rng.uniform(-8, 8, (M, K)) * 65536).astype(np.int64)
    b = (rng.uniform(-8, 8, (K, N)) * 65536).astype(np.int64)
    A, B = s.share_int(0, a), s.share_int(2, b)
    Ao, Bo = r.share_int(0, a), r.share_int(2, b)
    for _ in range(3):
        C = s.mul(A, B, shift=16)
        assert np.array_equal(s.get_shares(C), r.mul_trunc(Ao, Bo, 16))
        s.free(C)
    assert_cursors(s, r)


def test_copying_reshare_and_two_run_selection_still_match():
    """The round-2 shortcuts for co-located parties -- shared planes in the binary engine (ABY3_BIN_SHARED_PLANES) and the
    one-pass compare-exchange selection (ABY3_FUSED_MAXMIN), the common GEMV launch (ABY3_RING_GEMV) -- are switches read once
    per process: with all of them OFF (the paths parties on different GPUs take: AND rows packed / sent / scattered, two
    bitwiseAnd runs, one cross-term launch per party) the same tests give the same share planes against the oracle."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, ABY3_BIN_SHARED_PLANES="0", ABY3_FUSED_MAXMIN="0", ABY3_RING_GEMV="0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                        "-k", "basic_blocks or binary_engine or piecewise_logistic or random_circuits or colocated_ring"],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    assert " passed" in r.stdout


def test_colocated_ring_gemv_shares_bit_exact(pair):
    """A GEMV-shaped truncating product big enough for the co-located parties' common launch (sh3/Colocated.h: 2^22 elements of
    A, every share plane read once): all six share planes and the PRNG cursors against the oracle, twice (recycled buffers)."""
    s, r = pair
    M, K, d = 8192, 512, 16
    a = (np.random.default_rng(71).normal(0, 3, (M, K)) * (1 << d)).astype(np.int64)
    w = (np.random.default_rng(72).normal(0, 0.5, (K, 1)) * (1 << d)).astype(np.int64)
    A, W = s.share_int(0, a), s.share_int(1, w)
    Ao, Wo = r.share_int(0, a), r.share_int(1, w)
    for _ in range(2):
        C = s.mul(A, W, shift=d)
        Co = r.mul_trunc(Ao, Wo, d)
        assert np.array_equal(s.get_shares(C), Co)
        assert np.all(np.abs(s.reveal(C, 2) - (o.plain_mul(a, w) >> d)) <= 4)
        s.free(C)
    assert_cursors(s, r)

