"""Replays the sequence of tests/golden/make_golden.py on any engine with the oracle_lib.Session / harness.Session
vocabulary, comparing every step with the committed outputs of the reference's own code (tests/golden/ref_vectors.npz)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import cases  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ref_vectors.npz"))


def circuit(name):
    keys = ("gates", "level_gates", "input_first", "input_bits", "output_off", "output_bits", "output_wires", "output_invert")
    cir = {k: np.ascontiguousarray(GOLD["cir_%s_%s" % (name, k)]) for k in keys}
    cir["wire_count"] = int(GOLD["cir_%s_wire_count" % name][0])
    return cir


def eq(name, got):
    exp = GOLD[name]
    assert np.array_equal(np.asarray(got).reshape(exp.shape), exp), "differs from the reference's output: " + name
