"""Memory-safety / ordering evidence without compute-sanitizer (closed on this GPU pool: profiles/r2_sanitizer_closed.log).
What the sanitizer would have looked for is checked by the product's own debug mode ABY3_POOL_GUARD=1 (sh3/Gpu.h):
canary bytes around every pool block (out-of-bounds writes by any kernel), poison fill of every block that changes
hands AFTER the waits on its recorded readers (a reader the early-free pool / second stream failed to order reads
poison and the share-level comparison with the oracle fails) -- the analogue of the reference's own poison check
(aby3/sh3/Sh3BinaryEvaluator.cpp:578-621).  Plus repeat-run determinism of the hand-synchronised kernels."""
import os
import subprocess
import sys

import numpy as np
import pytest

from aby3_b200 import abi, harness

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
U64 = np.uint64


def _guarded(args, timeout=1500):
    env = dict(os.environ, ABY3_POOL_GUARD="1")
    return subprocess.run([sys.executable] + args, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=timeout)


def test_guard_catches_an_overrun_and_passes_a_clean_block():
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from aby3_b200 import harness\n"
            "s = harness.Session()\n"
            "print('CLEAN', s.guard_selftest(100000, 0), 'OVERRUN', s.guard_selftest(100000, 1), 'BIG', s.guard_selftest(9 << 20, 300))\n"
            "s.close()\n" % ROOT)
    out = _guarded(["-c", code])
    assert out.returncode == 0, out.stdout
    assert "CLEAN 0 OVERRUN 1 BIG 1" in out.stdout, out.stdout
    assert harness.Session().guard_selftest(1000, 1) == -1          # guard mode is off in this process


@pytest.mark.parametrize("target", ["smoke", "gemm", "sgd", "early", "binary"])
def test_targets_under_pool_guard(target):
    """tools/sanitize_targets.py (every target checks its own result) with canaries + poison-on-reuse on"""
    out = _guarded([os.path.join(ROOT, "tools", "sanitize_targets.py"), target])
    assert out.returncode == 0, out.stdout[-3000:]
    assert "pool guard:" not in out.stdout, out.stdout[-3000:]
    # (the "gemm" target drives the C ABI directly: no facade pool, no summary line)
    assert target == "gemm" or "[pool guard]" in out.stdout, out.stdout[-3000:]
    for line in out.stdout.splitlines():
        if line.startswith("[pool guard]"):
            assert line.rstrip().endswith(" 0 damaged"), line


def test_share_level_parity_suite_under_pool_guard():
    """the facade's share-level parity tests touching the pool's early-free path, the second stream, zero-copy opens
    (SharedBuffer / Borrowed) and the overlapped transfers -- all bit-exact against the oracle with poison-on-reuse on"""
    out = _guarded(["-m", "pytest", "tests/test_gpu_sh3.py", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", "-k",
                    "trunc or early or overlapped or chain or matmul_shares or fused_sgd or graph_sgd or basic_blocks or alias"])
    assert out.returncode == 0, out.stdout[-4000:]
    assert "pool guard:" not in out.stdout, out.stdout[-4000:]


def test_repeat_runs_are_bit_identical_under_concurrent_load():
    """k_gemm_tc (mbarrier pipeline, TMEM hand-over) and the persistent SGD kernel (split grid barrier, REDs): the same
    inputs 30 times, while another context keeps the GPU busy with keystream and GEMM work of its own, must give the
    same bits every time (a missing barrier shows up as a run-to-run difference long before it shows up as a crash)."""
    ctx, other = abi.Ctx(0), abi.Ctx(0)
    M, K, N = 384, 640, 320
    rng = np.random.default_rng(3)
    ops = [rng.integers(-2**63, 2**63, s, dtype=np.int64) for s in ((M, K), (M, K), (K, N), (K, N))]
    d = [ctx.upload(x) for x in ops]
    od = [other.upload(x) for x in ops]
    oc, fill = other.alloc(8 * M * N), other.alloc(1 << 24)
    c = ctx.alloc(8 * M * N)
    exp = (ops[0].view(U64) @ (ops[2].view(U64) + ops[3].view(U64)) + ops[1].view(U64) @ ops[2].view(U64)).view(np.int64)
    for it in range(30):
        abi.check(abi.lib.aby3cu_aes_ctr_fill(other.h, bytes(16), 0, fill.p, 1 << 24))
        abi.check(abi.lib.aby3cu_gemm_cross(other.h, abi.GEMM_TCGEN05, od[0].p, od[1].p, od[2].p, od[3].p, M, K, N, oc.p, 0))
        abi.check(abi.lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_TCGEN05, d[0].p, d[1].p, d[2].p, d[3].p, M, K, N, c.p, 0))
        assert np.array_equal(ctx.download(c, (M, N)), exp), it
    other.sync()
    ctx.close()
    other.close()
    x = rng.normal(1, 1, (400, 64))
    y = x[:, :3] @ np.array([[2.0], [-1.0], [0.5]])
    idx = rng.integers(0, 400, 40 * 32).astype(np.uint64)
    first = None
    for it in range(8):
        s = harness.Session()
        X, Y = s.share_int(0, (x * 65536).astype(np.int64)), s.share_int(0, (y * 65536).astype(np.int64))
        W = s.share_int(0, np.zeros((64, 1), dtype=np.int64))
        s.linreg_fused(X, Y, W, idx, 40, 32, 2.0 ** -6)
        w = s.get_shares(W)
        s.close()
        first = w if first is None else first
        assert np.array_equal(w, first), it
