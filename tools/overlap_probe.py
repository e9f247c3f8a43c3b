"""Does an AES-bound pre-kernel (truncation pair, 64 KiB of T-tables per CTA) really run UNDER the tcgen05 GEMM of another
party (148 persistent CTAs of 145 KiB)?  Two contexts (two streams) on one GPU: the GEMM alone, k truncation pairs alone,
both started from one event.  Prints one JSON line.  ABY3CU_NO_CARVEOUT=1 shows the driver's default carve-out."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aby3_b200 import abi  # noqa: E402

lib = abi.lib
KA, KB = bytes(range(16)), bytes(range(50, 66))


def main():
    g = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n = g * g
    c1, c2 = abi.Ctx(0), abi.Ctx(0)
    if os.environ.get('PROBE_CORUN', '1') == '1':
        abi.check(lib.aby3cu_ctx_set_corun(c2.h, 1))
    A = [c1.alloc(8 * n) for _ in range(5)]
    T = [c2.alloc(8 * n) for _ in range(3)]
    for b in A:
        abi.check(lib.aby3cu_aes_ctr_fill(c1.h, KA, 0, b.p, 8 * n))
    c1.sync()

    def gemm():
        abi.check(lib.aby3cu_gemm_cross(c1.h, abi.GEMM_TCGEN05, A[0].p, A[1].p, A[2].p, A[3].p, g, g, g, A[4].p, 1))

    def pairs():
        for _ in range(reps):
            abi.check(lib.aby3cu_trunc_tuple(c2.h, KA, 4, KB, 4, 16, None, T[0].p, T[1].p, T[2].p, n))

    def run(do_gemm, do_pairs):
        s, e1, e2 = c1.event(), c1.event(), c2.event()
        c1.sync(); c2.sync()
        c1.record(s)
        abi.check(lib.aby3cu_event_wait(c2.h, s))
        if do_gemm:
            gemm()
        if do_pairs:
            pairs()
        c1.record(e1)
        c2.record(e2)
        return abi.elapsed_ms(s, e1), abi.elapsed_ms(s, e2)

    for _ in range(2):
        run(True, True)
    out = {"size": g, "pairs": reps, "carveout_pref": os.environ.get("ABY3CU_NO_CARVEOUT", "0") != "1",
           "corun_ctas": os.environ.get("PROBE_CORUN", "1") == "1"}
    out["gemm_alone_ms"] = min(run(True, False)[0] for _ in range(3))
    out["pairs_alone_ms"] = min(run(False, True)[1] for _ in range(3))
    both = [run(True, True) for _ in range(3)]
    out["together_gemm_ms"] = min(b[0] for b in both)
    out["together_pairs_ms"] = min(b[1] for b in both)
    out["together_makespan_ms"] = min(max(b) for b in both)
    out["serial_sum_ms"] = out["gemm_alone_ms"] + out["pairs_alone_ms"]
    print(json.dumps(out), flush=True)
    c1.close(); c2.close()


if __name__ == "__main__":
    main()
