"""Run-to-run spread of the merge of 2 x n keys: device time, wall time and driver allocations of each of `reps` runs."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aby3_b200 import harness  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8388605
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    s = harness.Session()
    rng = np.random.default_rng(0)
    d1 = np.sort(rng.integers(-2**62, 2**62, n, dtype=np.int64)).reshape(-1, 1)
    d2 = np.sort(rng.integers(-2**62, 2**62, n, dtype=np.int64)).reshape(-1, 1)
    D1, D2 = s.share_bin(0, d1, 64), s.share_bin(0, d2, 64)
    for rep in range(reps):
        s.sync()
        p0 = s.pool_stats
        s.timer_begin()
        t0 = time.perf_counter()
        m = s.odd_even_merge(D1, D2)
        ms = s.timer_end()
        wall = (time.perf_counter() - t0) * 1e3
        p1 = s.pool_stats
        print("run %2d: device %8.2f ms  wall (enqueue) %8.2f ms  driver allocations %3d (%.2f GiB)  frees %d"
              % (rep, ms, wall, p1[0] - p0[0], (p1[1] - p0[1]) / 2**30, p1[2] - p0[2]), flush=True)
        s.free(m)
    s.close()


if __name__ == "__main__":
    main()
