import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native libraries once (no-op when they are up to date).  On the
    GPU box the prebuilt .so files travel with the snapshot."""
    from aby3_b200 import build
    try:
        build.build_all()
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    except Exception as e:  # pragma: no cover - surfaced by the tests that need the libs
        print("build step failed:", e)
    yield


def has_gpu():
    from aby3_b200 import abi
    return abi.device_count() > 0


@pytest.fixture(scope="session")
def ctx():
    from aby3_b200 import abi
    c = abi.Ctx(0)
    yield c
    c.close()
