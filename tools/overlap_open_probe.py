"""Does pulling finished row blocks of xy - r over NVLink WHILE the contraction is still running cost the contraction?
GPU 0: aby3cu_gemm_cross_blocks (4096^3, one launch, a progress event per 1024-row block); GPU 1: a stream that waits for
event b and peer-copies block b.  Reports: product alone, product then copies (serial), overlapped.  Needs two GPUs."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aby3_b200 import abi  # noqa: E402

lib = abi.lib


def main():
    n, nb = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 4
    c0, c1 = abi.Ctx(0), abi.Ctx(1)
    bufs = [c0.alloc(8 * n * n) for _ in range(5)]
    for i, b in enumerate(bufs):
        abi.check(lib.aby3cu_aes_ctr_fill(c0.h, bytes([i + 1] * 16), 0, b.p, 8 * n * n))
    dst = c1.alloc(8 * n * n)
    p = [b.p for b in bufs]
    rows = n // nb
    evs = (abi.C.c_void_p * nb)(*[c0.event() for _ in range(nb)])
    done = c1.event()

    def product(blocks):
        if blocks:
            abi.check(lib.aby3cu_gemm_cross_blocks(c0.h, abi.GEMM_TCGEN05, p[0], p[1], p[2], p[3], n, n, n, p[4], 1, None, rows, evs, nb))
        else:
            abi.check(lib.aby3cu_gemm_cross(c0.h, abi.GEMM_TCGEN05, p[0], p[1], p[2], p[3], n, n, n, p[4], 1))

    def copies(per_block):
        sl = 8 * n * rows
        for b in range(nb):
            if per_block:
                abi.check(lib.aby3cu_event_wait(c1.h, evs[b]))
            abi.check(lib.aby3cu_d2d(c1.h, dst.at(b * sl), 1, bufs[4].at(b * sl), 0, sl))

    out = {}
    for name, fn in (("product_alone", lambda: product(False)),
                     ("product_with_progress_events", lambda: product(True)),
                     ("serial", lambda: (product(True), abi.check(lib.aby3cu_event_wait(c1.h, evs[nb - 1])), copies(False))),
                     ("overlapped", lambda: (product(True), copies(True)))):
        for _ in range(3):
            fn()
        c0.sync(); c1.sync()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
            c0.sync(); c1.sync()
        out[name + "_ms"] = (time.perf_counter() - t0) / 10 * 1e3
    # when do the block events fire? (host time stamps of event completion after one launch)
    c0.sync(); c1.sync()
    t0 = time.perf_counter()
    product(True)
    stamps = []
    for b in range(nb):
        abi.check(lib.aby3cu_event_sync(evs[b]))
        stamps.append((time.perf_counter() - t0) * 1e3)
    c0.sync()
    out["block_event_ms_after_launch"] = stamps
    out["product_end_ms_after_launch"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    for _ in range(10):
        copies(False)
    c1.sync()
    out["copies_alone_ms"] = (time.perf_counter() - t0) / 10 * 1e3
    out["blocks"] = nb
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
