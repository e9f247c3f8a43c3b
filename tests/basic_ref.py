"""aby3-Basic building blocks composed from the oracle's primitives (test infrastructure):
same engine calls, in the same order, as aby3_b200/basic/Basics.h and the reference
(aby3-Basic/BoolBasic.cpp, BuildingBlocks.cpp, Sort.cpp)."""
import math

import numpy as np

import oracle_lib as o
from aby3_b200 import harness

U64 = np.uint64
_cache = {}


def cir(name, bits=64):
    if (name, bits) not in _cache:
        _cache[(name, bits)] = harness.library_circuit(name, bits)
    return _cache[(name, bits)]


def run(r, name, A, B):
    n = A.shape[2]
    outs, _ = o.bin_eval(r, cir(name), n, [np.ascontiguousarray(A), np.ascontiguousarray(B)])
    return outs[0]


def cipher_gt(r, A, B):
    """MSB(B - A): x0+x1 from party 0 (copied to party 1 unmasked), x2 from parties 1 / 2."""
    D = (B.view(U64) - A.view(U64)).view(np.int64)
    n = D.shape[2]
    in0 = np.zeros((3, 2, n, 1), dtype=np.int64)
    v = (D[0, 0].view(U64) + D[0, 1].view(U64)).view(np.int64)
    in0[0, 0] = v
    in0[1, 1] = v
    in1 = np.zeros((3, 2, n, 1), dtype=np.int64)
    in1[1, 0] = D[1, 0]
    in1[2, 1] = D[2, 1]
    return run(r, "add_msb", in0, in1)


def bool_not(X):
    Y = X.copy()
    Y[1, 0] = ~Y[1, 0]
    Y[2, 1] = ~Y[2, 1]
    return Y


def max_min_split(r, A, B):
    n = A.shape[2]
    comp = run(r, "lt", A, B)
    mask = (-(comp & 1)).astype(np.int64)                 # share-wise 0 / -1
    ext_comp = np.concatenate([mask, mask], axis=2)
    ext_ab = np.concatenate([A, B], axis=2)
    t1 = run(r, "and", ext_comp, ext_ab)
    t2 = run(r, "and", bool_not(ext_comp), ext_ab)
    mn = t1[:, :, :n] ^ t2[:, :, n:]
    mx = t1[:, :, n:] ^ t2[:, :, :n]
    return mx, mn


def odd_even_merge(r, D1, D2):
    l1, l2 = D1.shape[2], D2.shape[2]
    length = max(l1, l2)
    mx, _ = max_min_split(r, np.ascontiguousarray(D1[:, :, l1 - 1:l1]), np.ascontiguousarray(D2[:, :, l2 - 1:l2]))
    res = np.empty((3, 2, 2 * length, 1), dtype=np.int64)
    res[:] = mx
    res[:, :, 0:2 * l1:2] = D1
    res[:, :, 1:2 * l2:2] = D2
    t = int(math.ceil(math.log2(length) + 1))
    q, d, r0 = 1 << (t - 1), 1, 0
    while d > 0:
        xs = np.arange(r0, 2 * length - d, 2)
        if len(xs):
            ys = xs + d
            m1, m0 = max_min_split(r, np.ascontiguousarray(res[:, :, xs]), np.ascontiguousarray(res[:, :, ys]))
            res[:, :, xs] = m0
            res[:, :, ys] = m1
        d = q - 1
        q >>= 1
        r0 = 1
    return np.ascontiguousarray(res[:, :, :l1 + l2])
