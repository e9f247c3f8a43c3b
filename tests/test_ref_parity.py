"""Pins the oracle (oracle/oracle.cpp, the CPU restatement every GPU parity test compares with) against
the REFERENCE ITSELF: oracle/_ref/libaby3ref.so is the reference's own sh3 sources (Sh3Runtime,
Sh3Encryptor, Sh3ShareGen, Sh3Evaluator, SharedOT, Sh3BinaryEvaluator, Sh3Piecewise, CircuitLibrary)
compiled UNMODIFIED from /root/reference against stand-in headers for the absent third-party libraries
(oracle/shim/README.md).  Same seeds, same inputs -> every party's two share planes must be identical.

What this does NOT pin (restated in oracle/shim, not inspectable here): the cryptoTools primitives
(oc::PRNG / oc::AES keystream layout, oc::transpose bit order beyond what Sh3ConverterTests fixes,
BetaLibrary gate order).  Runs on CPU; on a box without /root/reference the prebuilt library is used."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as o
import ref_lib as r

pytestmark = pytest.mark.skipif(not r.available(), reason="oracle/_ref not built and /root/reference absent")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pair(enc=None, ev=None):
    e, v = o.default_seeds()
    return o.Session(enc or e, ev or v), r.Session(enc or e, ev or v)


def other_seeds():
    rng = np.random.default_rng(99)
    return bytes(rng.integers(0, 256, 96, dtype=np.uint8)), bytes(rng.integers(0, 256, 96, dtype=np.uint8))


def consistent_seeds():
    """party i's next seed == party i+1's prev seed, as a real deployment has them"""
    rng = np.random.default_rng(5)
    k = [bytes(rng.integers(0, 256, 16, dtype=np.uint8)) for _ in range(6)]
    enc = b"".join(k[i] + k[(i + 1) % 3] for i in range(3))
    ev = b"".join(k[3 + i] + k[3 + (i + 1) % 3] for i in range(3))
    return enc, ev


@pytest.mark.parametrize("seeds", ["default", "random"])
def test_share_and_reveal_int(seeds):
    so, sr = pair() if seeds == "default" else pair(*consistent_seeds())
    rng = np.random.default_rng(1)
    for owner, shape in ((0, (1, 1)), (1, (7, 5)), (2, (33, 17)), (0, (300, 3))):      # crosses the 256-block refill
        a = rng.integers(-2**63, 2**63, shape, dtype=np.int64)
        Ao, Ar = so.share_int(owner, a), sr.share_int(owner, a)
        assert np.array_equal(Ao, Ar)
        rev = sr.reveal_all(Ar)
        for p in range(3):
            assert np.array_equal(rev[p], a)
            assert np.array_equal(o.reveal(Ao, p), a)


def test_share_and_reveal_bin():
    so, sr = pair()
    rng = np.random.default_rng(2)
    for owner, shape in ((0, (5, 1)), (1, (64, 2)), (2, (700, 1))):
        a = rng.integers(-2**63, 2**63, shape, dtype=np.int64)
        Ao, Ar = so.share_bin(owner, a), sr.share_bin(owner, a)
        assert np.array_equal(Ao, Ar)
        rev = sr.reveal_all(Ar, binary=True)
        for p in range(3):
            assert np.array_equal(rev[p], a)
            assert np.array_equal(o.reveal(Ao, p, binary=True), a)


def test_asyncMul_si64Matrix_shares():
    """Sh3Evaluator.cpp:92-116 (this fork's element-wise cross term + zero share + reshare); chained
    products keep the zero-share cursor in step (Sh3EvaluatorTests.cpp: 10 chained rounds)."""
    so, sr = pair()
    rng = np.random.default_rng(3)
    a = rng.integers(-2**63, 2**63, (10, 10), dtype=np.int64)
    b = rng.integers(-2**63, 2**63, (10, 10), dtype=np.int64)
    Ao, Ar = so.share_int(0, a), sr.share_int(0, a)
    Bo, Br = so.share_int(1, b), sr.share_int(1, b)
    exp = a.copy()
    for _ in range(10):
        Ao, Ar = so.mul(Ao, Bo, mode=1), sr.mul(Ar, Br)
        assert np.array_equal(Ao, Ar)
        exp = exp * b
    assert np.array_equal(sr.reveal_all(Ar)[2], exp)
    # a size that needs several refills of the 256-block share buffer
    x = rng.integers(-2**63, 2**63, (40, 40), dtype=np.int64)
    Xo, Xr = so.share_int(2, x), sr.share_int(2, x)
    assert np.array_equal(so.mul(Xo, Xo, mode=1), sr.mul(Xr, Xr))


def test_getTruncationTuple_shares():
    """Sh3Evaluator.cpp:503-566, every party, several shifts; the draws advance the common PRNGs"""
    so, sr = pair()
    for d in (8, 16, 33):
        for p in range(3):
            for rows, cols in ((1, 1), (7, 5), (64, 33)):
                to, tr = so.trunc_tuple(p, rows * cols, d), sr.trunc_tuple(p, rows, cols, d)
                for x, y in zip(to, tr):
                    assert np.array_equal(x, y)


@pytest.mark.parametrize("disable", [False, True])
def test_asyncMul_truncating_shares(disable):
    """Sh3Evaluator.cpp:651-730 on square operands (the fork's element-wise overwrite, :664-665)"""
    so, sr = pair()
    so.disable_randomization(disable)
    sr.disable_randomization(disable)
    rng = np.random.default_rng(4)
    for n, d in ((6, 16), (24, 8), (40, 16)):
        a = (rng.normal(0, 100, (n, n)) * (1 << d)).astype(np.int64)
        b = (rng.normal(0, 100, (n, n)) * (1 << d)).astype(np.int64)
        Ao, Ar = so.share_int(2, a), sr.share_int(2, a)
        Bo, Br = so.share_int(0, b), sr.share_int(0, b)
        Co, Cr = so.mul_trunc(Ao, Bo, d, mode=1), sr.mul_trunc(Ar, Br, d)
        assert np.array_equal(Co, Cr)
        assert np.max(np.abs(sr.reveal_all(Cr)[0] - ((a * b) >> d))) <= (1 if disable else 4)


def _bits(s, b):
    return s.share_bin(0, b.reshape(-1, 1).astype(np.int64)) & 1          # Sh3EvaluatorTests.cpp:846-854


def test_asyncMul_bit_times_arithmetic_shares():
    """Sh3Evaluator.cpp:119-263 + aby3/OT/SharedOT.cpp; pinned by sh3_asyncArithBinMul_test (:780-900)"""
    so, sr = pair()
    rng = np.random.default_rng(5)
    for n in (1, 100, 300):
        a = rng.integers(-2**31, 2**31, (n, 1), dtype=np.int64)
        b = rng.integers(0, 2, n)
        Ao, Ar = so.share_int(0, a), sr.share_int(0, a)
        Bo, Br = _bits(so, b), _bits(sr, b)
        assert np.array_equal(Bo, Br)
        Co, Cr = so.mul_bit(Ao, Bo), sr.mul_bit(Ar, Br)
        assert np.array_equal(Co, Cr)
        assert np.array_equal(sr.reveal_all(Cr)[1], a * b.reshape(n, 1))


def test_asyncMul_bit_times_public_shares():
    """Sh3Evaluator.cpp:418-501; pinned by sh3_asyncPubArithBinMul_test (:903-1032)"""
    so, sr = pair()
    rng = np.random.default_rng(6)
    for n, a in ((5, 3), (300, -77), (64, 1 << 40)):
        b = rng.integers(0, 2, n)
        Bo, Br = _bits(so, b), _bits(sr, b)
        Co, Cr = so.mul_bit_pub(a, Bo), sr.mul_bit_pub(a, Br)
        assert np.array_equal(Co, Cr)
        assert np.array_equal(sr.reveal_all(Cr)[0].reshape(n), a * b)


def _lib_circuit(name, bits):
    from aby3_b200 import harness          # host-only use: the facade's circuit library as DATA for both engines
    return harness.library_circuit(name, bits)


@pytest.mark.parametrize("name,bits,width", [("and", 64, 1), ("and", 64, 300), ("or", 64, 77), ("nor", 64, 91), ("xor", 33, 65), ("add", 16, 100),
                                             ("add_depth", 64, 257), ("add_msb", 64, 2049), ("lt", 64, 130), ("eq", 64, 64),
                                             ("piecewise2", 64, 96)])
def test_binary_engine_shares(name, bits, width):
    """Sh3BinaryEvaluator.cpp setCir / setInput / roundCallback / getShares / getOutput on the same
    levelised circuit: output share planes identical (AND masks, reshare order, transposes, inversion flags)."""
    so, sr = pair()
    cir = _lib_circuit(name, bits)
    rng = np.random.default_rng(7)
    ins_o, ins_r, plain = [], [], []
    for k, nb in enumerate(cir["input_bits"]):
        nb = int(nb)
        words = (nb + 63) // 64
        v = rng.integers(-2**63, 2**63, (width, words), dtype=np.int64)
        if nb % 64:
            v[:, -1] &= (1 << (nb % 64)) - 1
        plain.append(v)
        ins_o.append(so.share_bin(k % 3, v))
        ins_r.append(sr.share_bin(k % 3, v))
    outs_o, _ = o.bin_eval(so, cir, width, ins_o)
    outs_r = sr.bin_eval(cir, width, ins_r)
    for k, (x, y) in enumerate(zip(outs_o, outs_r)):
        nb = int(cir["output_bits"][k])
        mask = np.int64(-1) if nb % 64 == 0 else np.int64((1 << (nb % 64)) - 1)
        x, y = x.copy(), y.copy()
        x[..., -1] &= mask                    # bits beyond bitCount are unspecified in both (sbMatrix::trim)
        y[..., -1] &= mask
        assert np.array_equal(x, y), "output %d of %s" % (k, name)
    if name == "and":
        assert np.array_equal(sr.reveal_all(outs_r[0], binary=True)[0], plain[0] & plain[1])
    if name == "add_depth":
        assert np.array_equal(sr.reveal_all(outs_r[0], binary=True)[0], plain[0] + plain[1])
    # both engines consumed the same number of bytes from the common PRNGs (one key block each)
    to, tr = so.trunc_tuple(0, 4, 16), sr.trunc_tuple(0, 4, 1, 16)
    assert all(np.array_equal(a, b) for a, b in zip(to, tr))


@pytest.mark.parametrize("seed,width", [(0, 1), (1, 64), (2, 65), (3, 300), (4, 2049), (5, 77), (6, 4096), (7, 31)])
def test_random_circuits_all_gate_types(seed, width):
    """Random circuits over every gate type the engine supports (Xor, And, Nor, Or, Nxor, copy, na_And) with random
    inverted outputs: oracle == the reference's evaluator share for share, and both reveal the plaintext evaluation."""
    import circuits_random as cr
    so, sr = pair()
    cir = cr.random_circuit(seed, n_gates=100 + 20 * seed)
    rng = np.random.default_rng(seed)
    ins = [rng.integers(0, 2 ** int(b), width, dtype=np.uint64) for b in cir["input_bits"]]
    So = [so.share_bin(k % 3, x.view(np.int64).reshape(width, 1)) for k, x in enumerate(ins)]
    Sr = [sr.share_bin(k % 3, x.view(np.int64).reshape(width, 1)) for k, x in enumerate(ins)]
    oo, _ = o.bin_eval(so, cir, width, So)
    rr = sr.bin_eval(cir, width, Sr)
    exp = cr.plain_eval(cir, ins)
    for k in range(len(oo)):
        m = np.int64((1 << int(cir["output_bits"][k])) - 1)
        assert np.array_equal(oo[k] & m, rr[k] & m), k
        assert np.array_equal(o.reveal(oo[k], 0, binary=True).reshape(width).view(np.uint64) & np.uint64(m), exp[k]), k


def test_piecewise_plain_matches_reference():
    """Sh3Piecewise::eval(i64Matrix) -- the plaintext evaluator aby3_tests/Sh3PiecewiseTests.cpp:13-80 pins"""
    import piecewise_ref as pw
    rng = np.random.default_rng(8)
    D = 16
    x = (rng.uniform(-2, 2, 500) * (1 << D)).astype(np.int64)
    for th, coef in (([-0.5, 0.5], [[], [0.5, 1], [1]]), ([0.0], [[], [0, 1]]), ([-1.0, 0.0, 1.0], [[1], [0.25, 2], [], [3, -1]])):
        assert np.array_equal(pw.plain(x, th, coef, D).reshape(-1), r.piecewise_plain(x, th, coef, D))


def test_piecewise_step_function_reveals_like_reference():
    """The reference's three-party Sh3Piecewise::eval (Sh3Piecewise.cpp:184-567: getInputRegions, the region
    circuit from its own CircuitLibrary, asyncMul(i64, sbMatrix) per region) against the oracle-side composition,
    on piecewise-CONSTANT functions.  The region circuit comes from each side's own circuit library, so share
    planes may differ; reconstructed outputs may not.

    Degree-1 regions cannot be checked this way: the reference passes functionOutputs[c] as both A and C of
    asyncMul(si64Matrix, sbMatrix) (Sh3Piecewise.cpp:296-300) and party 0 overwrites c[0](i), c[1](i) before it
    reads A[0](i) + A[1](i) (Sh3Evaluator.cpp:155-162), so its three-party result is garbage there (its own
    three-party test is skipped, aby3_tests/Sh3PiecewiseTests.cpp:129).  For those the pinned semantics are the
    plaintext evaluator's (test above), which the oracle and the device path reproduce."""
    import piecewise_ref as pw
    so, sr = pair()
    rng = np.random.default_rng(9)
    D, n = 16, 200
    x = (rng.uniform(-1.5, 1.5, (n, 1)) * (1 << D)).astype(np.int64)
    cir2, cir3 = _lib_circuit("piecewise2", 64), _lib_circuit("piecewise3", 64)
    for th, coef, cir in (([-0.5, 0.5], [[], [0.25], [1]], cir2), ([0.0, 1.0], [[3], [], [-2]], cir2),
                          ([-1.0, 0.0, 1.0], [[1], [2], [], [7]], cir3)):
        Xo, Xr = so.share_int(0, x), sr.share_int(0, x)
        Yr = sr.piecewise(Xr, th, coef, D)
        Yo = pw.shared(so, Xo, th, coef, D, cir)
        exp = pw.plain(x, th, coef, D)
        assert np.array_equal(sr.reveal_all(Yr)[0], exp)
        assert np.array_equal(o.reveal(Yo, 0), exp)
        for p in range(3):
            assert np.array_equal(Yr[(p + 1) % 3, 1], Yr[p, 0])          # a consistent replicated sharing


def test_reference_piecewise_aliasing_is_what_breaks_degree_one_regions():
    """Documents the finding above: same inputs, logisticFunc coefficients -- the reference's three-party
    result differs from its own plaintext evaluator exactly on the rows of the degree-1 region."""
    import piecewise_ref as pw
    _, sr = pair()
    rng = np.random.default_rng(10)
    D, n = 16, 64
    x = (rng.uniform(-1.5, 1.5, (n, 1)) * (1 << D)).astype(np.int64)
    th, coef = [-0.5, 0.5], [[], [0.5, 1], [1]]
    got = sr.reveal_all(sr.piecewise(sr.share_int(0, x), th, coef, D))[0]
    exp = r.piecewise_plain(x, th, coef, D).reshape(n, 1)
    middle = (x >= -(1 << (D - 1))) & (x < (1 << (D - 1)))
    assert np.array_equal(got[~middle], exp[~middle])
    assert middle.any() and not np.array_equal(got[middle], exp[middle])


def test_converter_arith_to_binary_shares():
    """Sh3Converter::toBinaryMatrix(si64Matrix) (Sh3Converter.cpp:63-209; Sh3_convert_arithToBinaryMatrix_test):
    oracle restatement vs the reference's code, share planes and reveals."""
    so, sr = pair()
    so.conv_init()
    sr.conv_init()
    rng = np.random.default_rng(11)
    for rows, cols in ((43, 2), (1, 1), (130, 1)):
        x = rng.integers(-2**63, 2**63, (rows, cols), dtype=np.int64)
        Xo, Xr = so.share_int(0, x), sr.share_int(0, x)
        cir = _lib_circuit("a2b", 64 * cols)
        Yo, Yr = so.conv_a2b(Xo, cir), sr.conv_a2b(Xr)
        assert np.array_equal(sr.reveal_all(Yr, binary=True)[0], x)
        assert np.array_equal(o.reveal(Yo, 1, binary=True), x)
        assert np.array_equal(Yo, Yr)


def test_converter_bit_injection_shares():
    """Sh3Converter::bitInjection (Sh3Converter.cpp:211-370; Sh3_convert_BitInjection_test: n = 43, m = 17)"""
    so, sr = pair()
    so.conv_init()
    sr.conv_init()
    rng = np.random.default_rng(12)
    for rows, bits in ((43, 17), (5, 64), (20, 91), (300, 1)):
        words = (bits + 63) // 64
        x = rng.integers(-2**63, 2**63, (rows, words), dtype=np.int64)
        if bits % 64:
            x[:, -1] &= (1 << (bits % 64)) - 1
        Bo, Br = so.share_bin(1, x), sr.share_bin(1, x)
        Yo, Yr = so.conv_bit_injection(Bo, bits), sr.conv_bit_injection(Br, bits)
        exp = np.zeros((rows, bits), dtype=np.int64)
        for j in range(bits):
            exp[:, j] = (x[:, j // 64] >> (j % 64)) & 1
        assert np.array_equal(sr.reveal_all(Yr)[2], exp)
        assert np.array_equal(o.reveal(Yo, 0), exp)
        assert np.array_equal(Yo, Yr)
    # the common PRNGs advanced identically
    assert all(np.array_equal(a, b) for a, b in zip(so.trunc_tuple(2, 3, 16), sr.trunc_tuple(2, 3, 1, 16)))


def test_converter_packed_layout():
    """toPackedBin / toBinaryMatrix(sPackedBin) (Sh3Converter.cpp:12-61; Sh3_convert_b64Matrix_PackedBin_test):
    the reference's packed layout equals the oracle's bit transpose, and the round trip is the identity."""
    rng = np.random.default_rng(13)
    for rows, bits in ((1, 1), (64, 64), (65, 63), (200, 130), (256, 256)):
        words = (bits + 63) // 64
        planes = rng.integers(-2**63, 2**63, (2, rows, words), dtype=np.int64)
        if bits % 64:
            planes[:, :, -1] &= (1 << (bits % 64)) - 1
        packed, back = r.conv_packed_roundtrip(planes, bits)
        assert np.array_equal(back, planes)
        simd = (rows + 63) // 64
        for p in range(2):
            t = o.bit_transpose(planes[p].view(np.uint8).reshape(-1), rows, bits, words * 8, simd * 8)
            assert np.array_equal(t.view(np.int64).reshape(bits, simd), packed[p])


def test_packed_binary_sharing_and_reveal():
    """Sh3Encryptor::localPackedBinary / remotePackedBinary / revealAll(sPackedBin) (Sh3Encryptor.cpp:342-425, 627-724)"""
    so, sr = pair()
    rng = np.random.default_rng(14)
    for owner, (rows, cols) in ((0, (1, 1)), (1, (65, 1)), (2, (200, 2)), (0, (64, 3))):
        x = rng.integers(-2**63, 2**63, (rows, cols), dtype=np.int64)
        sh_r, rev = sr.share_reveal_packed(owner, x)
        sh_o = so.share_packed(owner, x)
        assert np.array_equal(sh_o, sh_r)
        for p in range(3):
            assert np.array_equal(rev[p], x)


def test_scheduler_orders_hold_for_the_reference_runtime(tmp_path):
    """tests/cpp/test_runtime.cpp (the assertions of aby3_tests/Sh3RuntimeTests.cpp, which the facade's
    Sh3Runtime passes in test_cpp_runtime.py) compiled against the reference's own Sh3Runtime."""
    if not os.path.isdir(os.path.join(r.REFERENCE, "aby3", "sh3")):
        pytest.skip("needs the reference headers")
    exe = str(tmp_path / "test_runtime_ref")
    orc = os.path.join(ROOT, "oracle")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-w", "-maes", "-msse4.1", "-mavx2", "-DREF_RUNTIME",
                           "-I", os.path.join(orc, "shim"), "-I", r.REFERENCE, "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "test_runtime.cpp"), os.path.join(orc, "shim", "shim.cpp"),
                           os.path.join(orc, "_ref", "Sh3Runtime.o"), os.path.join(orc, "_ref", "Sh3Types.o")])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stdout
    assert "ALL OK" in out.stdout


# ---- the reference's APPLICATIONS (aby3-Basic, aby3-ML) compiled unmodified into oracle/_ref --------------------------
# These are the expected values of tests/test_compat.py (the same sources on the B200 facade) and bench.py's CPU figures
# for BASELINE configs 3 and 5.

def test_reference_aby3_basic_on_cpu():
    """aby3-Basic/BoolBasic.cpp, BuildingBlocks.cpp, Sort.cpp as the reference's own tests use them
    (aby3_tests/BoolTest.cpp:60-283, SortTest.cpp:354-486): reveals equal the plaintext functions."""
    e, v = o.default_seeds()
    s = r.Session(e, v)
    rng = np.random.default_rng(0)
    n = 700
    a = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
    b = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
    b[:20] = a[:20]
    A, B = s.share_bin(0, a), s.share_bin(1, b)
    plain = {"lt": (a < b).astype(np.int64), "eq": (a == b).astype(np.int64), "and": a & b, "or": a | b, "add": a + b,
             "max": np.maximum(a, b), "min": np.minimum(a, b)}
    for op, exp in plain.items():
        out, _ = s.basic_bool(op, A, B)
        rev = s.reveal_all(out, binary=True)
        for p in range(3):
            got = rev[p] & 1 if op in ("lt", "eq") else rev[p]
            assert np.array_equal(got, exp), (op, p)
    out, _ = s.cipher_gt(s.share_int(0, a), s.share_int(2, b))
    assert np.array_equal(s.reveal_all(out, binary=True)[1] & 1, (a > b).astype(np.int64))
    mx, mn, _ = s.max_min_split(A, B)
    assert np.array_equal(s.reveal_all(mx, binary=True)[0], np.maximum(a, b))
    assert np.array_equal(s.reveal_all(mn, binary=True)[2], np.minimum(a, b))
    d1 = np.sort(rng.integers(-2**40, 2**40, 50)).reshape(-1, 1)           # SortTest.cpp:363: 50 + 98 elements
    d2 = np.sort(rng.integers(-2**40, 2**40, 98)).reshape(-1, 1)
    out, _ = s.odd_even_merge(s.share_bin(0, d1), s.share_bin(0, d2))
    assert np.array_equal(s.reveal_all(out, binary=True)[0].reshape(-1), np.sort(np.concatenate([d1[:, 0], d2[:, 0]])))
    s.close()


def test_reference_aby3_ml_linear_on_cpu(capfd):
    """aby3-ML/main-linear.cpp's own entry point (three party threads over loopback sessions, LinearModelGen data,
    SGD_Linear) runs to completion on the stand-in headers and prints its iters/s line (main-linear.cpp:147-149);
    ref_lib.sgd_linear (the same engine + Regression.h on caller-supplied data) returns replicated w shares."""
    r.main_linear("-N", 600, "-D", 48, "-B", 16, "-I", 40, "-testN", 50)
    out = capfd.readouterr().out
    assert "iters/s" in out and "IT:40 =>" in out        # (party threads interleave their prints)
    rng = np.random.default_rng(1)
    x = rng.normal(1, 1, (300, 40))
    y = x[:, :3] @ np.array([2.0, -1.0, 0.5])
    t, w = r.sgd_linear(x, y, 16, 25, lr=2.0 ** -6)
    assert t > 0
    for p in range(3):
        assert np.array_equal(w[(p + 1) % 3, 1], w[p, 0])


def test_reference_role_tests_on_cpu():
    """aby3_tests/Test.cpp, BoolTest.cpp, SortTest.cpp compiled unmodified into oracle/_ref: the number of check_result()
    SUCCESS lines each writes is what tests/test_compat.py expects of the same programs on the GPU facade."""
    import compat_lib
    for name, n in compat_lib.ROLE_TESTS.items():
        if name.startswith("quick_sort"):
            continue                                   # 2.6 s / 66 s of CPU: their counts (1, 0) were taken from a run of oracle/_ref and are pinned by the GPU-side test
        assert r.role_test(name) == (n, 0), name
