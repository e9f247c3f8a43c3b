"""Parties on three different GPUs (distributed placement): the reshare crosses
NVLink either as in-process peer copies (transport "local") or as ncclSend/ncclRecv
(transport "nccl").  Same bit-exact oracle comparison as the co-located tests."""
import numpy as np
import pytest

import oracle_lib as o
from aby3_b200 import abi, harness

pytestmark = pytest.mark.gpu
U64 = np.uint64

needs3 = pytest.mark.skipif(abi.device_count() < 3, reason="needs three GPUs")


def rnd(seed, shape):
    return np.random.default_rng(seed).integers(-2**63, 2**63, shape, dtype=np.int64)


@needs3
@pytest.mark.parametrize("transport", ["local", "nccl"])
def test_three_gpu_mul_and_trunc(transport):
    s, r = harness.Session(devices=(0, 1, 2), transport=transport), o.Session()
    try:
        a, b = rnd(1, (300, 200)), rnd(2, (200, 130))
        A, B = s.share_int(0, a), s.share_int(1, b)
        Ao, Bo = r.share_int(0, a), r.share_int(1, b)
        C = s.mul(A, B)
        Co = r.mul(Ao, Bo)
        assert np.array_equal(s.get_shares(C), Co)
        for p in range(3):
            assert np.array_equal(s.reveal(C, p), o.plain_mul(a, b))
        fa = (np.random.default_rng(3).normal(0, 10, (257, 64)) * 65536).astype(np.int64)
        fb = (np.random.default_rng(4).normal(0, 10, (64, 96)) * 65536).astype(np.int64)
        A, B = s.share_int(0, fa), s.share_int(2, fb)
        Ao, Bo = r.share_int(0, fa), r.share_int(2, fb)
        for _ in range(3):
            C = s.mul(A, B, shift=16)
            Co = r.mul_trunc(Ao, Bo, 16)
            assert np.array_equal(s.get_shares(C), Co)
        for p in range(3):
            assert list(s.cursors(p)) == list(r.cursors(p))
    finally:
        s.close()
        r.close()


@needs3
def test_three_gpu_binary_engine_nccl():
    s, r = harness.Session(devices=(0, 1, 2), transport="nccl"), o.Session()
    try:
        width = 5000
        x, y = rnd(5, (width, 1)), rnd(6, (width, 1))
        cir = harness.library_circuit("lt", 64)
        X, Y = s.share_bin(0, x, 64), s.share_bin(1, y, 64)
        Xo, Yo = r.share_bin(0, x), r.share_bin(1, y)
        out = s.bin_eval(cir, [X, Y])[0]
        outo, _ = o.bin_eval(r, cir, width, [Xo, Yo])
        assert np.array_equal(s.get_shares(out, binary=True) & 1, outo[0] & 1)
        assert np.array_equal(s.reveal(out, 0, binary=True) & 1, (x < y).astype(np.int64))
    finally:
        s.close()
        r.close()
