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


@pytest.mark.skipif(abi.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("devices,transport", [((0, 1, 2), "local"), ((0, 1, 2), "nccl"), ((0, 0, 1), "local"), ((1, 0, 0), "local")])
def test_compare_exchange_blocks_across_gpus(devices, transport):
    """aby3-Basic's selection and merge with the parties on different GPUs (and on a MIXED placement, where only one of a
    party's neighbours shares its GPU): the binary engine must take the copying reshare there (no shared planes), the one-pass
    selection sends its result plane over the channel -- same share planes as the oracle."""
    import basic_ref as br
    if max(devices) >= abi.device_count():
        pytest.skip("needs %d GPUs" % (max(devices) + 1))
    s, r = harness.Session(devices=devices, transport=transport), o.Session()
    try:
        rng = np.random.default_rng(11)
        n = 3001
        a, b = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64), rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
        X, Y = s.share_bin(0, a, 64), s.share_bin(2, b, 64)
        Xo, Yo = r.share_bin(0, a), r.share_bin(2, b)
        mx, mn = s.max_min_split(X, Y)
        mxo, mno = br.max_min_split(r, Xo, Yo)
        assert np.array_equal(s.get_shares(mx, binary=True), mxo)
        assert np.array_equal(s.get_shares(mn, binary=True), mno)
        d1 = np.sort(rng.integers(-2**40, 2**40, 50)).reshape(-1, 1).astype(np.int64)
        d2 = np.sort(rng.integers(-2**40, 2**40, 98)).reshape(-1, 1).astype(np.int64)
        m = s.odd_even_merge(s.share_bin(0, d1, 64), s.share_bin(0, d2, 64))
        mo = br.odd_even_merge(r, r.share_bin(0, d1), r.share_bin(0, d2))
        assert np.array_equal(s.get_shares(m, binary=True), mo)
        for p in range(3):
            assert list(s.cursors(p)) == list(r.cursors(p))
    finally:
        s.close()
        r.close()

