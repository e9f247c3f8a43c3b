"""CPU tests: the oracle's binary engine evaluated on the facade's circuit library
(host-only circuit export) reproduces the reference's checks --
aby3_tests/Sh3BinaryEvaluatorTests.cpp:333-424 (AND, add, add_msb: revealed
c(j) == op(a(j), b(j))), plus lt / eq / or as used by aby3-Basic."""
import numpy as np
import pytest

import oracle_lib as o
from aby3_b200 import harness

U64 = np.uint64


def mask(bits):
    return (1 << bits) - 1 if bits < 64 else (1 << 64) - 1


def signed(x, bits):
    x = int(x) & mask(bits)
    return x - (1 << bits) if x >> (bits - 1) else x


def expected(name, a, b, bits):
    m = mask(bits)
    if name == "and":
        return a & b
    if name == "or":
        return a | b
    if name == "xor":
        return a ^ b
    if name == "nor":
        return ~(a | b) & m
    if name in ("add", "add_depth"):
        return (a + b) & m
    if name == "add_msb":
        return ((a + b) & m) >> (bits - 1)
    if name == "lt":
        return int(signed(a, bits) < signed(b, bits))
    if name == "eq":
        return int((a & m) == (b & m))
    raise KeyError(name)


CASES = [("and", 8), ("and", 64), ("or", 64), ("nor", 64), ("nor", 9), ("xor", 17), ("add", 8), ("add", 64), ("add_depth", 8), ("add_depth", 64),
         ("add_msb", 8), ("add_msb", 64), ("lt", 8), ("lt", 64), ("eq", 64), ("eq", 13)]


@pytest.mark.parametrize("name,bits", CASES)
def test_library_circuit_on_oracle_engine(name, bits):
    width = 256 if bits <= 8 else 70              # Sh3BinaryEvaluatorTests.cpp: width 256, 8-bit operands
    rng = np.random.default_rng(hash((name, bits)) % 2**32)
    a = rng.integers(0, 2**63, width, dtype=np.uint64) * U64(2) + rng.integers(0, 2, width, dtype=np.uint64)
    b = rng.integers(0, 2**63, width, dtype=np.uint64) * U64(2) + rng.integers(0, 2, width, dtype=np.uint64)
    if name in ("lt", "eq"):
        b[: width // 4] = a[: width // 4]         # exercise the equal case
    a &= U64(mask(bits)); b &= U64(mask(bits))
    cir = harness.library_circuit(name, bits)
    assert cir["level_gates"].sum() == len(cir["gates"]) // 4
    s = o.Session()
    A = s.share_bin(0, a.view(np.int64).reshape(width, 1))
    B = s.share_bin(1, b.view(np.int64).reshape(width, 1))
    outs, _ = o.bin_eval(s, cir, width, [A, B])
    obits = int(cir["output_bits"][0])
    for p in range(3):
        got = o.reveal(outs[0], p, binary=True).reshape(width).view(U64) & U64(mask(obits))
        exp = np.array([expected(name, int(x), int(y), bits) for x, y in zip(a, b)], dtype=U64)
        assert np.array_equal(got, exp), (name, bits, p)
        assert np.array_equal(outs[0][(p + 1) % 3, 1], outs[0][p, 0])


def test_depth_of_depth_optimised_circuits():
    assert len(harness.library_circuit("and", 64)["level_gates"]) == 1
    assert len(harness.library_circuit("add_depth", 64)["level_gates"]) <= 8
    assert len(harness.library_circuit("lt", 64)["level_gates"]) <= 8
    assert len(harness.library_circuit("eq", 64)["level_gates"]) <= 7


def test_binary_engine_consumes_one_key_block_from_each_common_prng():
    s = o.Session()
    cir = harness.library_circuit("and", 8)
    z = np.zeros((3, 2, 16, 1), dtype=np.int64)
    before = s.cursors(0).copy()
    o.bin_eval(s, cir, 16, [z, z])
    after = s.cursors(0)
    assert after[2] == before[2] + 16 and after[3] == before[3] + 16


def test_and_gate_mask_is_the_documented_keystream():
    """z for nonlinear gate #g = KS_prev ^ KS_next at bytes [g*rowBytes, (g+1)*rowBytes)
    (Sh3BinaryEvaluator.cpp:1406-1442) -- checked through the wire memory on zero inputs."""
    s = o.Session()
    width, bits = 100, 8
    cir = harness.library_circuit("and", bits)
    z = np.zeros((3, 2, width, 1), dtype=np.int64)
    _, mem = o.bin_eval(s, cir, width, [z, z])
    rb = mem.shape[3]
    for p in range(3):
        kp = o.keystream(s.seed("eval", p, 0), 32, 16).tobytes()     # setCir draws at cursor 32
        kn = o.keystream(s.seed("eval", p, 1), 32, 16).tobytes()
        for g in range(bits):
            out_wire = int(cir["gates"][4 * g + 2])
            exp = o.keystream(kp, g * rb, rb) ^ o.keystream(kn, g * rb, rb)
            assert np.array_equal(mem[p, 0, out_wire], exp)


def test_piecewise_logistic_on_oracle():
    """aby3-ML's logistic approximation (aby3ML.h:121-139): f(x) = 0 | x + 0.5 | 1 on the oracle,
    exact against the plaintext evaluator (Sh3_Piecewise_plain_test semantics)."""
    import piecewise_ref as pw
    D, n = 16, 300
    rng = np.random.default_rng(0)
    x = (rng.uniform(-2, 2, (n, 1)) * (1 << D)).astype(np.int64)
    x[:4, 0] = [-(1 << 15), (1 << 15), -(1 << 15) - 1, (1 << 15) - 1]      # on and next to the thresholds
    th, coef = [-0.5, 0.5], [[], [0.5, 1], [1]]
    s = o.Session()
    X = s.share_int(0, x)
    cir = harness.library_circuit("piecewise2", 64)
    out = pw.shared(s, X, th, coef, D, cir)
    exp = pw.plain(x, th, coef, D)
    xf = x.astype(np.float64) / (1 << D)
    assert np.array_equal(exp.reshape(-1), (np.clip(xf + 0.5, 0, 1) * (1 << D)).astype(np.int64).reshape(-1))
    for p in range(3):
        assert np.array_equal(o.reveal(out, p), exp)
        assert np.array_equal(out[(p + 1) % 3, 1], out[p, 0])


def test_piecewise_relu_on_oracle():
    """Sh3PiecewiseTests.cpp:13-80: max(0, x) as a one-threshold piecewise function."""
    import piecewise_ref as pw
    D, n = 16, 100
    x = (np.random.default_rng(1).uniform(-5, 5, (n, 1)) * (1 << D)).astype(np.int64)
    s = o.Session()
    out = pw.shared(s, s.share_int(1, x), [0], [[], [0, 1]], D, harness.library_circuit("piecewise1", 64))
    assert np.array_equal(o.reveal(out, 0), np.maximum(x, 0))


def test_basic_blocks_on_oracle():
    """aby3_tests/Test.cpp / BoolTest.cpp / SortTest.cpp semantics on the oracle: gt, max/min split,
    odd-even merge of two sorted runs (50 + 98 elements as in SortTest.cpp:363)."""
    import basic_ref as br
    rng = np.random.default_rng(0)
    s = o.Session()
    n = 16
    a = np.arange(n, dtype=np.int64).reshape(n, 1)
    b = (n - np.arange(n, dtype=np.int64)).reshape(n, 1)
    gt = br.cipher_gt(s, s.share_int(0, a), s.share_int(0, b))
    assert np.array_equal(o.reveal(gt, 0, binary=True) & 1, (a > b).astype(np.int64))
    x = rng.integers(-2**62, 2**62, (40, 1), dtype=np.int64)
    y = rng.integers(-2**62, 2**62, (40, 1), dtype=np.int64)
    mx, mn = br.max_min_split(s, s.share_bin(0, x), s.share_bin(1, y))
    assert np.array_equal(o.reveal(mx, 0, binary=True), np.maximum(x, y))
    assert np.array_equal(o.reveal(mn, 2, binary=True), np.minimum(x, y))
    d1 = np.sort(rng.integers(-2**40, 2**40, 50)).reshape(-1, 1).astype(np.int64)
    d2 = np.sort(rng.integers(-2**40, 2**40, 98)).reshape(-1, 1).astype(np.int64)
    merged = br.odd_even_merge(s, s.share_bin(0, d1), s.share_bin(0, d2))
    got = o.reveal(merged, 1, binary=True).reshape(-1)
    assert sorted(got.tolist()) == sorted(np.concatenate([d1, d2]).reshape(-1).tolist())
    assert np.all(np.diff(got) >= 0), "the merge network of Sort.cpp:327-406 must sort two sorted runs"
