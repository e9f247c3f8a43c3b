"""CPU-side checks of the C ABI: the library loads, exports every symbol that
include/aby3cu.h declares, the host-only key-draw helper matches the oracle, and
compute entry points fail loudly (no CPU fallback) when no B200 is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as o
from aby3_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "aby3cu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aby3cu_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(abi.lib, s), "libaby3cu.so does not export %s" % s
        assert s in abi.PROTOTYPES, "abi.py has no prototype for %s" % s
    assert sorted(abi.PROTOTYPES) == syms


def test_version():
    assert abi.lib.aby3cu_version() == 1


def test_host_keystream_matches_oracle():
    rng = np.random.default_rng(0)
    for off, n in [(0, 16), (16, 16), (8, 24), (5, 37), (4090, 100), (0, 4096)]:
        key = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
        assert np.array_equal(abi.host_keystream(key, off, n), o.keystream(key, off, n))


def test_host_keystream_is_bounded():
    with pytest.raises(abi.Aby3CudaError):
        abi.host_keystream(bytes(16), 0, 4097)


def test_bin_row_bytes():
    for w in [1, 2047, 2048, 2049, 1 << 24]:
        assert abi.lib.aby3cu_bin_row_bytes(w) == o.lib.orc_bin_row_bytes(w) == 256 * ((w + 2047) // 2048)


@pytest.mark.skipif(abi.device_count() > 0, reason="only meaningful on a box without a GPU")
def test_no_gpu_means_loud_failure():
    with pytest.raises(abi.Aby3CudaError) as e:
        abi.Ctx(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_touch_oracle():
    """Nothing under aby3_b200/ may import, link or name the oracle."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "aby3_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                t = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"liboracle|oracle_lib|orc_[a-z]+\(|#include\s+\"[^\"]*oracle", t):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
