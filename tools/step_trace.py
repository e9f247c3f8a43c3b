"""Kernel timeline of the 4096^3 truncating product (three co-located parties) without a profiler: ABY3CU_TRACE=1 makes
every launch record an event behind itself; this prints, for one steady-state step, the END time of every kernel on
every stream (ms since the step's first kernel end minus nothing: absolute since trace_begin) and the gap to the
previous kernel end on the same stream.  Usage: python tools/step_trace.py [size] [steps]"""
import collections
import csv
import os
import sys

os.environ["ABY3CU_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from aby3_b200 import abi, harness  # noqa: E402


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/step_trace.csv"
    sess = harness.Session(devices=(0, 0, 0))
    rng = np.random.default_rng(0)
    a = (rng.normal(0, 30, (size, size)) * 65536).astype(np.int64)
    b = (rng.normal(0, 30, (size, size)) * 65536).astype(np.int64)
    A, B = sess.share_int(0, a), sess.share_int(0, b)
    C = sess.mul(A, B, shift=16)
    for _ in range(3):
        sess.mul(A, B, shift=16, out=C)
    sess.sync()
    probe = abi.Ctx(0)
    abi.check(abi.lib.aby3cu_trace_begin(probe.h))
    for _ in range(steps):
        sess.mul(A, B, shift=16, out=C)
    sess.sync()
    abi.check(abi.lib.aby3cu_trace_dump(out.encode()))
    rows = list(csv.DictReader(open(out)))
    streams = {}
    for r in rows:
        streams.setdefault(r["stream"], len(streams))
    last = collections.defaultdict(float)
    print("%8s %3s %-18s %8s" % ("end_ms", "str", "kernel", "since_prev_on_stream"))
    for r in sorted(rows, key=lambda r: float(r["end_ms"])):
        s = streams[r["stream"]]
        t = float(r["end_ms"])
        print("%8.3f %3d %-18s %8.3f" % (t, s, r["kernel"], t - last[s]))
        last[s] = t
    sess.close()
    probe.close()


if __name__ == "__main__":
    main()
