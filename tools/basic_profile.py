"""Where does a compare-exchange stage spend its time?  Times the pieces of
bool_cipher_max_min_split at merge scale through the harness (wall clock + device)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aby3_b200 import harness  # noqa: E402


def timed(sess, name, fn, reps=3):
    fn()
    sess.sync()
    l0 = sess.launches
    sess.timer_begin()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    ms = sess.timer_end()
    wall = (time.perf_counter() - t0) * 1e3
    print("%-34s device %9.3f ms  wall %9.3f ms  launches %5d" % (name, ms / reps, wall / reps, (sess.launches - l0) // reps), flush=True)
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 21) - 3
    s = harness.Session()
    rng = np.random.default_rng(0)
    a = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
    b = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
    A, B = s.share_bin(0, a, 64), s.share_bin(0, b, 64)
    if os.environ.get("ABY3_BASIC_ONLY") == "split":      # for an ncu launch list of one compare-exchange
        for _ in range(2):
            [s.free(h) for h in s.max_min_split(A, B)]
            s.sync()
        s.close()
        return
    for name in ("and", "lt", "add_msb", "eq"):
        cir = harness.library_circuit(name, 64)
        print(name, "gates", len(cir["gates"]) // 4, "nonlinear", cir["nonlinear"], "levels", len(cir["level_gates"]), "wires", cir["wire_count"])
        timed(s, "bin_eval %s n=%d" % (name, n), lambda: [s.free(h) for h in s.bin_eval(cir, [A, B])])
    timed(s, "max_min_split n=%d" % n, lambda: [s.free(h) for h in s.max_min_split(A, B)])
    timed(s, "max_min_split n=%d (10 reps)" % n, lambda: [s.free(h) for h in s.max_min_split(A, B)], reps=10)
    # one merge of two sorted halves of n elements each: where do the stages spend their time?
    d1 = np.sort(a[:, 0]).reshape(-1, 1)
    d2 = np.sort(b[:, 0]).reshape(-1, 1)
    D1, D2 = s.share_bin(0, d1, 64), s.share_bin(0, d2, 64)
    s.sync()
    p0 = s.pool_stats
    l0 = s.launches
    s.timer_begin()
    t0 = time.perf_counter()
    m = s.odd_even_merge(D1, D2)
    ms = s.timer_end()
    wall = (time.perf_counter() - t0) * 1e3
    p1 = s.pool_stats
    print("odd_even_merge 2 x %d: device %.1f ms wall %.1f ms launches %d; pool: %d mallocs (%.1f GiB), %d frees"
          % (n, ms, wall, s.launches - l0, p1[0] - p0[0], (p1[1] - p0[1]) / 2**30, p1[2] - p0[2]), flush=True)
    s.free(m)
    s.sync()
    p0 = s.pool_stats
    s.timer_begin()
    m = s.odd_even_merge(D1, D2)
    ms = s.timer_end()
    p1 = s.pool_stats
    print("odd_even_merge again:       device %.1f ms; pool: %d mallocs (%.1f GiB), %d frees"
          % (ms, p1[0] - p0[0], (p1[1] - p0[1]) / 2**30, p1[2] - p0[2]), flush=True)
    s.close()


if __name__ == "__main__":
    main()
