"""Sh3Piecewise on the oracle's primitives (test infrastructure): the same message flow and the
same order of PRNG draws as aby3/sh3/Sh3Piecewise.cpp:184-567, built from oracle_lib calls."""
import numpy as np

import oracle_lib as o

U64 = np.uint64


def fixed(v, D):
    return int(v) * (1 << D) if isinstance(v, (int, np.integer)) else int(v * (1 << D))


def plain(x, thresholds, coefficients, D):
    """plaintext evaluation (Sh3Piecewise.cpp:60-183) for degree <= 1, integer slopes"""
    x = np.asarray(x, dtype=np.int64).reshape(-1)
    th = [x < fixed(t, D) for t in thresholds]
    regions = [th[0]] + [(~th[t - 1]) & th[t] for t in range(1, len(th))] + [~th[-1]]
    out = np.zeros_like(x)
    for reg, coef in zip(regions, coefficients):
        if not coef:
            continue
        f = np.full_like(x, fixed(coef[0], D))
        if len(coef) > 1:
            f = f + int(coef[1]) * x
        out = out + reg.astype(np.int64) * f
    return out.reshape(-1, 1)


def shared(r, X, thresholds, coefficients, D, cir):
    n = X.shape[2]
    c0 = np.zeros((3, 2, n, 1), dtype=np.int64)
    v = (X[0, 0].view(U64) + X[0, 1].view(U64)).view(np.int64)       # P0: x0 + x1
    c0[0, 0] = v
    c0[1, 1] = v                                                       # sent to P1 without a mask
    c1 = np.zeros((3, 2, n, 1), dtype=np.int64)
    c1[1, 0] = X[1, 0]                                                 # x of party 1 ...
    c1[2, 1] = X[2, 1]                                                 # ... which party 2 holds as its prev share
    ins = []
    for t in thresholds:
        ct = c0.copy()
        ct[0, 0] -= fixed(t, D)
        ct[1, 1] -= fixed(t, D)
        ins.append(ct)
    ins.append(c1)
    regions, _ = o.bin_eval(r, cir, n, ins)
    out = np.zeros((3, 2, n, 1), dtype=np.int64)
    for c, coef in enumerate(coefficients):
        if not coef:
            continue
        if len(coef) > 1:
            f = (X.view(U64) * U64(int(coef[1]) & (2**64 - 1))).view(np.int64)
            k = fixed(coef[0], D)
            f[0, 0] += k
            f[1, 1] += k
            res = r.mul_bit(np.ascontiguousarray(f), np.ascontiguousarray(regions[c]))
        else:
            res = r.mul_bit_pub(fixed(coef[0], D), np.ascontiguousarray(regions[c]))
        out = (out.view(U64) + res.view(U64)).view(np.int64)
    return out
