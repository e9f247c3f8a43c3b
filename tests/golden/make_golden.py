"""Generates tests/golden/ref_vectors.npz: outputs of the REFERENCE'S OWN CODE (oracle/_ref = the reference's sh3
sources compiled from /root/reference against oracle/shim) on fixed seeds and inputs.  Run here, where the reference
tree exists:  python tests/golden/make_golden.py
The vectors let the oracle (tests/test_golden.py) and the CUDA path (tests/test_gpu_golden.py) be checked against the
reference on machines where neither /root/reference nor the prebuilt oracle/_ref library exists."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle_lib as o          # noqa: E402  (seed layout helpers only)
import ref_lib as r             # noqa: E402


def cases():
    """Deterministic inputs shared with the tests (which regenerate them from the same seeds)."""
    rng = np.random.default_rng(20261018)
    c = {}
    c["a"] = rng.integers(-2**63, 2**63, (12, 12), dtype=np.int64)
    c["b"] = rng.integers(-2**63, 2**63, (12, 12), dtype=np.int64)
    c["fa"] = (rng.normal(0, 50, (12, 12)) * (1 << 16)).astype(np.int64)
    c["fb"] = (rng.normal(0, 50, (12, 12)) * (1 << 16)).astype(np.int64)
    c["bits"] = rng.integers(0, 2, (70, 1), dtype=np.int64)
    c["av"] = rng.integers(-2**31, 2**31, (70, 1), dtype=np.int64)
    c["x"] = rng.integers(-2**63, 2**63, (150, 1), dtype=np.int64)
    c["y"] = rng.integers(-2**63, 2**63, (150, 1), dtype=np.int64)
    c["conv"] = rng.integers(-2**63, 2**63, (33, 2), dtype=np.int64)
    c["inj"] = rng.integers(0, 2**17, (21, 1), dtype=np.int64)
    c["pk"] = rng.integers(-2**63, 2**63, (70, 1), dtype=np.int64)
    return c


def main():
    from aby3_b200 import harness       # host-only: circuits as data
    e, v = o.default_seeds()
    s = r.Session(e, v)
    c = cases()
    out = {}
    A, B = s.share_int(0, c["a"]), s.share_int(1, c["b"])
    out["share_a"], out["share_b"] = A, B
    out["mul_hadamard"] = s.mul(A, B)                                    # Sh3Evaluator.cpp:92-116
    for p in range(3):
        R, T0, T1 = s.trunc_tuple(p, 5, 7, 16)                           # :503-566
        out["trunc_R_%d" % p], out["trunc_T0_%d" % p], out["trunc_T1_%d" % p] = R, T0, T1
    FA, FB = s.share_int(2, c["fa"]), s.share_int(0, c["fb"])
    out["share_fa"], out["share_fb"] = FA, FB
    out["mul_trunc_hadamard_16"] = s.mul_trunc(FA, FB, 16)               # :651-730 (square operands)
    Bb = s.share_bin(0, c["bits"]) & 1
    Av = s.share_int(1, c["av"])
    out["bit_shares"], out["share_av"] = Bb, Av
    out["mul_bit"] = s.mul_bit(Av, Bb)                                   # :119-263
    out["mul_bit_pub"] = s.mul_bit_pub(-12345, Bb)                       # :418-501
    X, Y = s.share_bin(0, c["x"]), s.share_bin(2, c["y"])
    out["share_x"], out["share_y"] = X, Y
    for name in ("and", "add_depth", "lt"):
        cir = harness.library_circuit(name, 64)
        out["bin_" + name] = s.bin_eval(cir, 150, [X, Y])[0]             # Sh3BinaryEvaluator.cpp
        for k, val in cir.items():
            if isinstance(val, np.ndarray):
                out["cir_%s_%s" % (name, k)] = val
        out["cir_%s_wire_count" % name] = np.array([cir["wire_count"]])
    s.conv_init()
    CV = s.share_int(0, c["conv"])
    out["share_conv"] = CV
    out["conv_a2b"] = s.conv_a2b(CV)                                     # Sh3Converter.cpp:63-209
    IJ = s.share_bin(1, c["inj"])
    out["share_inj"] = IJ
    out["conv_bit_injection_17"] = s.conv_bit_injection(IJ, 17)          # :211-370
    sh, rev = s.share_reveal_packed(2, c["pk"])                          # Sh3Encryptor.cpp:342-425
    out["packed_shares"], out["packed_reveal"] = sh, rev
    out["final_trunc_R_0"] = s.trunc_tuple(0, 4, 1, 16)[0]               # where every common PRNG ended up
    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %.1f KiB)" % (path, len(out), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
