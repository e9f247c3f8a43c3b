"""Where does an SGD_Linear iteration (BASELINE configs[2]) spend its time?  Host trace per statement
(ABY3_SGD_TRACE) + wall / launches; run under ncu for the per-kernel device times."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ABY3_SGD_TRACE", "1")
from aby3_b200 import harness  # noqa: E402


def main():
    for transport in ("local", "shared_stream"):
        run(transport)


def run(transport):
    samples, features, batch, iters = 1 << 16, 1024, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 300
    sess = harness.Session(transport=transport)
    rng = np.random.default_rng(11)
    x = (rng.normal(1.0, 1.0, (samples, features)) * 65536).astype(np.int64)
    y = (rng.normal(1.0, 1.0, (samples, 1)) * 65536).astype(np.int64)
    X, Y = sess.share_int(0, x), sess.share_int(0, y)
    W = sess.share_int(0, np.zeros((features, 1), dtype=np.int64))
    idx = rng.integers(0, samples, iters * batch).astype(np.uint64)
    sess.linreg(X, Y, W, idx[:20 * batch], 20, batch, 2.0 ** -10)
    sess.sync()
    l0 = sess.launches
    sess.timer_begin()
    t0 = time.perf_counter()
    sess.linreg(X, Y, W, idx, iters, batch, 2.0 ** -10)
    wall = time.perf_counter() - t0
    ms = sess.timer_end()
    print("[%s] linreg: %.1f us/iter wall, %.1f us/iter device span, %.1f launches/iter" % (transport, wall * 1e6 / iters, ms * 1e3 / iters, (sess.launches - l0) / iters), flush=True)
    # the graph-replayed trainer on the same data
    W2 = sess.share_int(0, np.zeros((features, 1), dtype=np.int64))
    sess.linreg_graph(X, Y, W2, idx[:20 * batch], 20, batch, 2.0 ** -10)
    sess.sync()
    l0 = sess.launches
    t0 = time.perf_counter()
    sess.linreg_graph(X, Y, W2, idx, iters, batch, 2.0 ** -10)
    sess.sync()
    wall = time.perf_counter() - t0
    print("[%s] linreg_graph: %.1f us/iter wall, %.1f kernels/iter" % (transport, wall * 1e6 / iters, (sess.launches - l0) / iters), flush=True)
    sess.close()


if __name__ == "__main__":
    main()
