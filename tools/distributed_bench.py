"""Distributed placement: party p on GPU p, the ring reshare and the truncation opens cross NVLink.
Times the sf64<D16> 4096^3 product + truncation for both transports ("local" = cudaMemcpyPeerAsync inside one process,
"nccl" = ncclSend/ncclRecv, one group per protocol step) against the co-located placement.  Needs three GPUs.
Timing: host wall clock over K steps with all three devices drained on both sides (the parties' streams live on
different devices, so one pair of CUDA events cannot bracket them)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aby3_b200 import harness  # noqa: E402


def run(devices, transport, n=4096, steps=10, shift=16, blocks=None):
    s = harness.Session(devices=devices, transport=transport)
    if blocks is not None:
        s.set_open_blocks(blocks)
    rng = np.random.default_rng(0)
    a = (rng.uniform(-4, 4, (n, n)) * (1 << shift)).astype(np.int64)
    b = (rng.uniform(-4, 4, (n, n)) * (1 << shift)).astype(np.int64)
    A, B = s.share_int(0, a), s.share_int(0, b)
    C = s.mul(A, B, shift=shift)
    for _ in range(3):
        s.mul(A, B, shift=shift, out=C)
    s.sync()
    sent0 = s.bytes_sent
    t0 = time.perf_counter()
    for _ in range(steps):
        s.mul(A, B, shift=shift, out=C)
    s.sync()
    dt = (time.perf_counter() - t0) / steps
    sent = (s.bytes_sent - sent0) / steps
    c = s.reveal(C, 0)
    err = int(np.max(np.abs(c[:8] - ((a[:8] @ b) >> shift))))
    s.close()
    return {"devices": list(devices), "transport": transport, "open_blocks": blocks, "ms_per_step": dt * 1e3, "ring_mac_per_s": n ** 3 / dt,
            "bytes_between_parties_per_step": sent, "reshare_GBps": sent / dt / 1e9, "max_abs_err_ulp": err}


def main():
    out = [run((0, 0, 0), "local")]
    for blocks in (1, 2, 4, 8):
        out.append(run((0, 1, 2), "local", blocks=blocks))
        out.append(run((0, 1, 2), "nccl", blocks=blocks))
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
