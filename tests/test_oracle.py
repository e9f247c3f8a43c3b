"""CPU tests of the oracle: AES known-answer vectors, the oc::PRNG restatement,
and the reference's own reconstruction-level checks re-expressed on the oracle
(aby3_tests/Sh3EvaluatorTests.cpp, Sh3EncryptorTests.cpp).  These are what pin
the oracle; raw share / keystream values are "parity unpinned" (see oracle.h)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as o

U64 = np.uint64


def _hex(s):
    return bytes.fromhex(s.replace(" ", ""))


# FIPS-197 Appendix B and C.1, NIST SP 800-38A F.1.1 (ECB-AES128.Encrypt)
KATS = [
    ("2b7e151628aed2a6abf7158809cf4f3c", "3243f6a8885a308d313198a2e0370734", "3925841d02dc09fbdc118597196a0b32"),
    ("000102030405060708090a0b0c0d0e0f", "00112233445566778899aabbccddeeff", "69c4e0d86a7b0430d8cdb78070b4c55a"),
    ("2b7e151628aed2a6abf7158809cf4f3c", "6bc1bee22e409f96e93d7e117393172a", "3ad77bb40d7a3660a89ecaf32466ef97"),
    ("2b7e151628aed2a6abf7158809cf4f3c", "ae2d8a571e03ac9c9eb76fac45af8e51", "f5d3d58503b9699de785895a96fdbaaf"),
    ("2b7e151628aed2a6abf7158809cf4f3c", "30c81c46a35ce411e5fbc1191a0a52ef", "43b1cd7f598ece23881b00e3ed030688"),
    ("2b7e151628aed2a6abf7158809cf4f3c", "f69f2445df4f9b17ad2b417be66c3710", "7b0c785e27e8ad3f8223207104725dd4"),
]


@pytest.mark.parametrize("soft", [0, 1])
def test_aes_known_answers(soft):
    for k, p, c in KATS:
        out = np.zeros(16, dtype=np.uint8)
        o.lib.orc_aes128_encrypt(_hex(k), _hex(p), o.ptr(out), soft)
        assert out.tobytes() == _hex(c)


def test_selftest_and_soft_equals_aesni():
    assert o.lib.orc_selftest() == 0
    rng = np.random.default_rng(1)
    for _ in range(64):
        k = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
        p = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
        a, b = np.zeros(16, np.uint8), np.zeros(16, np.uint8)
        o.lib.orc_aes128_encrypt(k, p, o.ptr(a), 0)
        o.lib.orc_aes128_encrypt(k, p, o.ptr(b), 1)
        assert a.tobytes() == b.tobytes()


def test_ctr_block_is_aes_of_little_endian_counter():
    key = _hex("000102030405060708090a0b0c0d0e0f")
    out = np.zeros(32, np.uint8)
    o.lib.orc_aes_ctr_blocks(key, 5, 2, o.ptr(out))
    for i in range(2):
        ref = np.zeros(16, np.uint8)
        o.lib.orc_aes128_encrypt(key, (5 + i).to_bytes(8, "little") + bytes(8), o.ptr(ref), 1)
        assert out[16 * i:16 * i + 16].tobytes() == ref.tobytes()


def test_prng_is_one_contiguous_keystream():
    """oc::PRNG: whatever the draw sizes (16-byte keys, bulk 8n-byte fills crossing the
    256-block buffer, direct-encrypt path), the bytes are AES-CTR from counter 0."""
    seed = o.to_block(0, 7)
    p = o.lib.orc_prng_new(seed, 256)
    sizes = [16, 16, 8, 4096 - 40, 8, 8 * 5000, 128, 8 * 3, 16 * 300 + 8, 8]
    got = []
    for n in sizes:
        b = np.zeros(n, np.uint8)
        o.lib.orc_prng_get(p, o.ptr(b), n)
        got.append(b)
    total = sum(sizes)
    assert o.lib.orc_prng_bytes_consumed(p) == total
    o.lib.orc_prng_free(p)
    assert np.array_equal(np.concatenate(got), o.keystream(seed, 0, total))


def test_session_init_cursors():
    s = o.Session()
    for p in range(3):
        c = s.cursors(p)
        # enc: one key block from each common PRNG; eval: key block + SharedOT seed block
        assert list(c) == [0, 0, 32, 32, 16, 16]


def test_share_int_matches_stream_formula_and_reveals():
    s = o.Session()
    rng = np.random.default_rng(2)
    for n in [1, 2, 511, 512, 513, 1500]:
        c0 = s.cursors(0)[0]
        m = rng.integers(-2**63, 2**63, n, dtype=np.int64)
        sh = s.share_int(0, m)
        for p in range(3):
            kp = o.keystream(s.seed("enc", p, 0), 0, 16).tobytes()
            kn = o.keystream(s.seed("enc", p, 1), 0, 16).tobytes()
            z = o.stream_u64(kp, int(c0), n) - o.stream_u64(kn, int(c0), n)
            exp = z + (m.view(U64) if p == 0 else U64(0))
            assert np.array_equal(sh[p, 0].view(U64), exp)
            assert np.array_equal(sh[(p + 1) % 3, 1], sh[p, 0])      # replicated consistency
        for p in range(3):
            assert np.array_equal(o.reveal(sh, p), m)
        assert s.cursors(0)[0] == c0 + n


def test_share_bin_reveals():
    s = o.Session()
    m = np.random.default_rng(3).integers(-2**63, 2**63, (40, 2), dtype=np.int64)
    sh = s.share_bin(1, m)
    for p in range(3):
        assert np.array_equal(o.reveal(sh, p, binary=True), m)


def _rand(prng_seed, shape):
    return np.random.default_rng(prng_seed).integers(-2**63, 2**63, shape, dtype=np.int64)


def test_Sh3_Evaluator_asyncMul_test():
    """aby3_tests/Sh3EvaluatorTests.cpp:20-135: 10x10, ten chained products with A = C + A."""
    s = o.Session()
    n = 10
    for t in range(3):
        a, b = _rand(10 + t, (n, n)), _rand(20 + t, (n, n))
        A, B = s.share_int(0, a), s.share_int(0, b)
        c = None
        for _ in range(n):
            Cs = s.mul(A, B, mode=0)
            A = (Cs.view(U64) + A.view(U64)).view(np.int64)
            c = o.plain_mul(a, b)
            a = (c.view(U64) + a.view(U64)).view(np.int64)
        for p in range(3):
            assert np.array_equal(o.reveal(Cs, p), c)
            assert np.array_equal(Cs[(p + 1) % 3, 1], Cs[p, 0])


def test_mul_hadamard_fork_semantics():
    """aby3_tests/Test.cpp:116,153,184: n x 1 'matrices' multiply element-wise."""
    s = o.Session()
    n = 16
    a = np.arange(n, dtype=np.int64).reshape(n, 1)
    b = (n - np.arange(n, dtype=np.int64)).reshape(n, 1)
    A, B = s.share_int(0, a), s.share_int(1, b)
    Cs = s.mul(A, B, mode=1)
    assert np.array_equal(o.reveal(Cs, 2), a * b)


def test_mul_zero_share_sums_to_zero_and_uses_eval_stream():
    s = o.Session()
    a, b = _rand(5, (7, 5)), _rand(6, (5, 3))
    A, B = s.share_int(0, a), s.share_int(0, b)
    Cs = s.mul(A, B)
    for p in range(3):
        kp = o.keystream(s.seed("eval", p, 0), 0, 16).tobytes()
        kn = o.keystream(s.seed("eval", p, 1), 0, 16).tobytes()
        z = (o.stream_u64(kp, 0, 21) - o.stream_u64(kn, 0, 21)).reshape(7, 3)
        cross = o.cross_term(A[p, 0], A[p, 1], B[p, 0], B[p, 1])
        assert np.array_equal(Cs[p, 0].view(U64), cross.view(U64) + z)
    assert s.cursors(1)[1] == 21


def test_Sh3_Evaluator_truncationPai_test():
    """Sh3EvaluatorTests.cpp:350-410: |reveal(RTrunc) - (sum R >> d)| < 4, D8, 4x4."""
    s = o.Session()
    d, n = 8, 16
    for _ in range(200):
        t = [s.trunc_tuple(p, n, d) for p in range(3)]
        # RTrunc is a replicated sharing: plane 0 of party p == plane 1 of party p+1
        for p in range(3):
            assert np.array_equal(t[p][1], t[(p + 1) % 3][2])
        tr = (t[0][1].view(U64) + t[1][1].view(U64) + t[2][1].view(U64)).view(np.int64)
        r = (t[0][0].view(U64) + t[1][0].view(U64) + t[2][0].view(U64)).view(np.int64)
        exp = r >> d
        assert np.all((tr > exp - 4) & (tr < exp + 4))


def test_trunc_tuple_stream_formula():
    s = o.Session()
    n, d = 37, 16
    R, T0, T1 = s.trunc_tuple(2, n, d)
    kn, kp = s.seed("eval", 2, 1), s.seed("eval", 2, 0)
    t0 = o.stream_u64(kn, 4, n).view(np.int64)       # cursor stands at byte 32 = element 4
    t1 = o.stream_u64(kp, 4, n).view(np.int64)
    assert np.array_equal(R, t0 >> 2)
    assert np.array_equal(T0, t0 >> (d + 2))
    assert np.array_equal(T1, t1 >> (d + 2))
    assert list(s.cursors(2)[2:4]) == [32 + 8 * n, 32 + 8 * n]


def _fixed(vals, d):
    return (vals * (1 << d)).astype(np.int64)       # fp::operator=(double): truncation toward zero


def test_Sh3_Evaluator_asyncMul_matrixFixed_test():
    """Sh3EvaluatorTests.cpp:413-589: D8, randomisation disabled, |c - reveal(C)| <= 1 ulp."""
    d, size = 8, 4
    s = o.Session()
    s.disable_randomization(True)
    for tt in range(20):
        rng = np.random.default_rng(tt)
        a = _fixed((rng.integers(0, 2**32, (size, size), dtype=np.uint64) >> U64(8)).astype(np.float64) / 100.0, d)
        b = _fixed((rng.integers(0, 2**32, (size, size), dtype=np.uint64) >> U64(8)).astype(np.float64) / 100.0, d)
        c = o.plain_mul(a, b) >> d
        A, B = s.share_int(0, a), s.share_int(0, b)
        Cs = s.mul_trunc(A, B, d)
        for p in range(3):
            cc = o.reveal(Cs, p)
            assert np.all(np.abs(cc - c) <= 1)
            assert np.array_equal(Cs[(p + 1) % 3, 1], Cs[p, 0])


def test_mul_trunc_randomised_is_close_and_consistent():
    """With the truncation pair on, the result is within a few ulp."""
    d = 16
    s = o.Session()
    rng = np.random.default_rng(9)
    a = _fixed(rng.normal(0, 50, (9, 6)), d)
    b = _fixed(rng.normal(0, 50, (6, 5)), d)
    A, B = s.share_int(0, a), s.share_int(2, b)
    Cs = s.mul_trunc(A, B, d)
    c = o.plain_mul(a, b) >> d
    for p in range(3):
        assert np.all(np.abs(o.reveal(Cs, p) - c) <= 4)
        assert np.array_equal(Cs[(p + 1) % 3, 1], Cs[p, 0])


def test_threaded_oracle_equals_sequential():
    a, b = _rand(1, (33, 17)), _rand(2, (17, 29))
    outs = []
    for nt in (1, 3, 6):
        s = o.Session()
        A, B = s.share_int(0, a), s.share_int(0, b)
        outs.append((s.mul(A, B, nthreads=nt), s.mul_trunc(A, B, 16, nthreads=nt)))
    for x in outs[1:]:
        assert np.array_equal(x[0], outs[0][0]) and np.array_equal(x[1], outs[0][1])


def test_bit_transpose_definition():
    """aby3_tests/Sh3ConverterTests.cpp:12-43: out bit (r,c) == in bit (c,r), LSB first."""
    rng = np.random.default_rng(4)
    for rows, cols in [(1, 1), (7, 64), (100, 13), (256, 256), (65, 130)]:
        ins, outs = (cols + 7) // 8 + 3, (rows + 7) // 8 + 1
        m = rng.integers(0, 256, rows * ins, dtype=np.uint8)
        t = o.bit_transpose(m, rows, cols, ins, outs)
        bits_in = np.unpackbits(m.reshape(rows, ins), axis=1, bitorder="little")[:, :cols]
        bits_out = np.unpackbits(t.reshape(cols, outs), axis=1, bitorder="little")[:, :rows]
        assert np.array_equal(bits_out, bits_in.T)


def _bit_shares(s, bits):
    """a one-bit sbMatrix the way the binary engine hands it out: only bit 0 carries data"""
    sh = s.share_bin(0, bits.reshape(-1, 1).astype(np.int64))
    return sh & 1


def test_sh3_asyncArithBinMul_test():
    """aby3_tests/Sh3EvaluatorTests.cpp:780-900: c = b * a exactly, and a consistent sharing."""
    s = o.Session()
    rng = np.random.default_rng(3)
    for n in (1, 7, 128, 1000):
        a = rng.integers(-2**63, 2**63, (n, 1), dtype=np.int64)
        b = rng.integers(0, 2, n)
        A, B = s.share_int(0, a), _bit_shares(s, b)
        Cs = s.mul_bit(A, B)
        for p in range(3):
            assert np.array_equal(o.reveal(Cs, p), a * b.reshape(n, 1))
            assert np.array_equal(Cs[(p + 1) % 3, 1], Cs[p, 0])


def test_sh3_asyncPubArithBinMul_test():
    """Sh3EvaluatorTests.cpp:903-1032: c = b * a for a public constant a."""
    s = o.Session()
    rng = np.random.default_rng(4)
    for n, a in ((5, 3), (300, -77), (64, 1 << 40)):
        b = rng.integers(0, 2, n)
        Cs = s.mul_bit_pub(a, _bit_shares(s, b))
        for p in range(3):
            assert np.array_equal(o.reveal(Cs, p).reshape(n), a * b)
            assert np.array_equal(Cs[(p + 1) % 3, 1], Cs[p, 0])


def test_sampled_rows_helper_is_the_oracle_evaluated_lazily():
    """tests/sampled_rows.py (used by the GPU tests at 4096^3) against the C oracle's full result at small sizes,
    across two products each so that non-zero keystream cursors are covered."""
    import sampled_rows as sr
    r = o.Session()
    rng = np.random.default_rng(0)
    for (M, K, N) in ((37, 20, 11), (5, 3, 700)):
        a = rng.integers(-2**63, 2**63, (M, K), dtype=np.int64)
        b = rng.integers(-2**63, 2**63, (K, N), dtype=np.int64)
        A, B = r.share_int(0, a), r.share_int(1, b)
        rows = np.array([0, M - 1, M // 2, 0])
        for _ in range(2):
            cur = [r.cursors(p) for p in range(3)]
            C = r.mul(A, B)
            assert np.array_equal(sr.mul_rows(r, cur, A, B, rows), C[:, :, rows])
            cur = [r.cursors(p) for p in range(3)]
            C = r.mul_trunc(A, B, 16)
            assert np.array_equal(sr.mul_trunc_rows(r, cur, A, B, 16, rows), C[:, :, rows])
    r.close()
