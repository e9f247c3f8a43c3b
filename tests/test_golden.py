"""The oracle against golden vectors produced by the REFERENCE'S OWN CODE (tests/golden/make_golden.py ->
tests/golden/ref_vectors.npz; the generator runs where /root/reference exists).  Same seeds, same inputs, same order of
calls -> every share plane identical.  CPU only; needs neither the reference tree nor oracle/_ref."""
import numpy as np

import oracle_lib as o
from golden_cases import GOLD, cases, circuit, eq


def test_oracle_reproduces_reference_golden_vectors():
    s = o.Session()
    c = cases()
    A, B = s.share_int(0, c["a"]), s.share_int(1, c["b"])
    eq("share_a", A)
    eq("share_b", B)
    eq("mul_hadamard", s.mul(A, B, mode=1))
    for p in range(3):
        R, T0, T1 = s.trunc_tuple(p, 35, 16)
        eq("trunc_R_%d" % p, R)
        eq("trunc_T0_%d" % p, T0)
        eq("trunc_T1_%d" % p, T1)
    FA, FB = s.share_int(2, c["fa"]), s.share_int(0, c["fb"])
    eq("share_fa", FA)
    eq("mul_trunc_hadamard_16", s.mul_trunc(FA, FB, 16, mode=1))
    Bb = s.share_bin(0, c["bits"]) & 1
    Av = s.share_int(1, c["av"])
    eq("bit_shares", Bb)
    eq("mul_bit", s.mul_bit(Av, Bb))
    eq("mul_bit_pub", s.mul_bit_pub(-12345, Bb))
    X, Y = s.share_bin(0, c["x"]), s.share_bin(2, c["y"])
    eq("share_x", X)
    for name in ("and", "add_depth", "lt"):
        outs, _ = o.bin_eval(s, circuit(name), 150, [X, Y])
        got = outs[0]
        if name == "lt":
            got, exp = got & 1, GOLD["bin_lt"] & 1          # one output bit; the rest of the word is unspecified
            assert np.array_equal(got, exp)
        else:
            eq("bin_" + name, got)
    s.conv_init()
    CV = s.share_int(0, c["conv"])
    eq("share_conv", CV)
    eq("conv_a2b", s.conv_a2b(CV, _a2b_circuit()))
    IJ = s.share_bin(1, c["inj"])
    eq("conv_bit_injection_17", s.conv_bit_injection(IJ, 17))
    eq("packed_shares", s.share_packed(2, c["pk"]))
    eq("final_trunc_R_0", s.trunc_tuple(0, 4, 16)[0])


def _a2b_circuit():
    from aby3_b200 import harness       # host-only: the adder circuit as data
    return harness.library_circuit("a2b", 128)


def test_reference_reveals_in_golden_file_are_the_plaintexts():
    c = cases()
    assert np.array_equal(GOLD["packed_reveal"][0], c["pk"])
    assert np.array_equal(o.reveal(GOLD["mul_hadamard"], 0), c["a"] * c["b"])
    assert np.array_equal(o.reveal(GOLD["bin_and"], 1, binary=True), c["x"] & c["y"])
    assert np.array_equal(o.reveal(GOLD["bin_add_depth"], 2, binary=True), c["x"] + c["y"])
    assert np.array_equal(o.reveal(GOLD["bin_lt"], 0, binary=True) & 1, (c["x"] < c["y"]).astype(np.int64))
    assert np.array_equal(o.reveal(GOLD["conv_a2b"], 0, binary=True), c["conv"])
    assert np.array_equal(o.reveal(GOLD["mul_bit"], 0), c["av"] * c["bits"])
    exp = np.stack([(c["inj"][:, 0] >> j) & 1 for j in range(17)], axis=1)
    assert np.array_equal(o.reveal(GOLD["conv_bit_injection_17"], 1), exp)
