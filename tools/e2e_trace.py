"""Kernel END-time trace (ABY3CU_TRACE=1, no profiler) of the end-to-end leg of bench.py: row-block streamed, steps pipelined
two deep, 4096^3.  Prints one steady-state step: every kernel end on every stream, the gap since the previous end on that
stream, and how much of the step lies between consecutive GEMM ends.  Usage: python tools/e2e_trace.py [row_blocks] [steps]"""
import collections
import csv
import os
import sys

os.environ["ABY3CU_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from aby3_b200 import abi, harness  # noqa: E402

SHIFT = 16


def main():
    NB = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    M = K = N = 4096
    sess = harness.Session(devices=(0, 0, 0))
    rng = np.random.default_rng(0)
    a = (rng.uniform(-4, 4, (M, K)) * (1 << SHIFT)).astype(np.int64)
    b = (rng.uniform(-4, 4, (K, N)) * (1 << SHIFT)).astype(np.int64)
    rb = M // NB

    def make_set():
        st = {"pb": sess.plain(0, K, N), "pa": [], "pc": []}
        st["pb"][1][...] = b
        for i in range(NB):
            pid, view = sess.plain(0, rb, K)
            view[...] = a[i * rb:(i + 1) * rb]
            st["pa"].append(pid)
            st["pc"].append(sess.plain(0, rb, N))
        return st

    sets = [make_set(), make_set()]

    def upload(st):
        sess.plain_touch(0, st["pb"][0])
        for pid in st["pa"]:
            sess.plain_touch(0, pid)
        sess.plain_prefetch(0, st["pb"][0])
        for pid in st["pa"]:
            sess.plain_prefetch(0, pid)

    def compute(st):
        hb = sess.share_plain(0, st["pb"][0], K, N)
        live = [hb]
        for i in range(NB):
            ha = sess.share_plain(0, st["pa"][i], rb, K)
            hc = sess.mul(ha, hb, shift=SHIFT)
            sess.reveal_plain_async(hc, 0, st["pc"][i][0])
            live += [ha, hc]
        return live

    def finish(st, live):
        for i in range(NB):
            sess.plain_wait(0, st["pc"][i][0])
        for h in live:
            sess.free(h)

    def run(nsteps):
        upload(sets[0])
        prev = None
        for k in range(nsteps):
            live = compute(sets[k % 2])
            if prev is not None:
                finish(*prev)
            prev = (sets[k % 2], live)
            if k + 1 < nsteps:
                upload(sets[(k + 1) % 2])
        finish(*prev)

    run(3)
    run(3)
    sess.sync()
    probe = abi.Ctx(0)
    abi.check(abi.lib.aby3cu_trace_begin(probe.h))
    import time
    t0 = time.perf_counter()
    run(steps)
    sess.sync()
    wall = (time.perf_counter() - t0) * 1e3 / steps
    out = "gpurun_out/e2e_trace_nb%d.csv" % NB
    abi.check(abi.lib.aby3cu_trace_dump(out.encode()))
    rows = sorted(csv.DictReader(open(out)), key=lambda r: float(r["end_ms"]))
    streams = {}
    for r in rows:
        streams.setdefault(r["stream"], len(streams))
    gemm_ends = [float(r["end_ms"]) for r in rows if "gemm_tc" in r["kernel"] and "ready" not in r["kernel"]]
    per_step = 3 * NB
    print("row blocks %d: %.3f ms per step (wall), %d GEMM launches traced" % (NB, wall, len(gemm_ends)))
    # one steady-state step in the middle
    mid = (steps // 2) * per_step
    lo, hi = gemm_ends[mid - 1], gemm_ends[mid + per_step - 1]
    print("steady-state step: %.3f ms between GEMM end #%d and #%d" % (hi - lo, mid - 1, mid + per_step - 1))
    last = collections.defaultdict(float)
    for r in rows:
        t = float(r["end_ms"])
        s = streams[r["stream"]]
        if lo - 0.5 <= t <= hi + 0.2:
            print("%9.3f s%-2d %-26s +%.3f" % (t - lo, s, r["kernel"][:26], t - last[s]))
        last[s] = t
    sess.close()
    probe.close()


if __name__ == "__main__":
    main()
