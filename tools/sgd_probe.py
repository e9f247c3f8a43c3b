"""Where does an iteration of the persistent SGD kernel (csrc/sgd_fused.cu) spend its time?  Runs the kernel through the C
ABI on synthetic shares with ABY3CU_SGD_STAMPS=1 and prints, per stretch between the stamps of CTA 0, the median SM-clock
cycles over iterations 8..63, plus the event-timed microseconds per iteration."""
import ctypes as C
import json
import os
import sys

os.environ["ABY3CU_SGD_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from aby3_b200 import abi  # noqa: E402

lib = abi.lib


def main():
    rows, F, B, iters = 1 << 18, 1024, 128, 2000
    ctx = abi.Ctx(0)
    rng = np.random.default_rng(0)
    key = [bytes(rng.integers(0, 256, 16, dtype=np.uint8)) for _ in range(6)]
    X = [ctx.alloc(rows * F * 8) for _ in range(6)]
    for i, b in enumerate(X):
        abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, key[i], 0, b.p, rows * F * 8))
    Y = [ctx.upload(rng.integers(-2**40, 2**40, rows)) for _ in range(6)]
    W = [ctx.upload(rng.integers(-2**20, 2**20, F)) for _ in range(6)]
    idx = ctx.upload(rng.integers(0, rows, iters * B).astype(np.uint64))
    wb = lib.aby3cu_sgd_linear_colocated_work_bytes(B)
    work = ctx.alloc(wb)
    P6 = C.c_void_p * 6
    P3 = C.c_char_p * 3
    U3 = C.c_uint64 * 3
    args = (ctx.h, P6(*[b.p.value for b in X]), P6(*[b.p.value for b in Y]), P6(*[b.p.value for b in W]), idx.p, F, B, iters, 16, 33,
            P3(*key[:3]), U3(0, 0, 0), P3(*key[3:]), U3(0, 0, 0), work.p)
    abi.check(lib.aby3cu_sgd_linear_colocated(*args))
    ctx.sync()
    a, b = ctx.event(), ctx.event()
    ctx.record(a)
    abi.check(lib.aby3cu_sgd_linear_colocated(*args))
    ctx.record(b)
    ms = abi.elapsed_ms(a, b)
    full = ctx.download(work, (64, 16), dtype=np.int64, byte_off=(15 * B + 8) * 8)[8:]
    variant = os.environ.get("ABY3CU_SGD_VARIANT", "slab")
    if variant.startswith("r"):
        names = {"phase A (w loads, dot, reduce, write)": (0, 1), "arrive 1: red.release": (8, 9), "pre-issue 1": (9, 2), "wait 1": (2, 3),
                 "phase B (V1/E from L2)": (3, 4), "phase C (mac, reduce, open, w)": (4, 5), "arrive 2: red.release": (10, 11),
                 "pre-issue 2": (11, 6), "wait 2": (6, 7)}
    else:
        names = {"first product + REDs + owner duty": (0, 2), "arrive (red.release)": (2, 3), "keystream words": (3, 4), "wait": (4, 5),
                 "issue of the S / E loads": (5, 8), "issue of the next pieces, offsets, prefetches": (8, 9),
                 "S / E arrive, shared": (9, 6), "second product, open, w": (6, 7)}
    out = {"variant": variant, "us_per_iter": ms * 1e3 / iters, "iters_per_s": iters / ms * 1e3,
           "median_cycles": {n: int(np.median(full[:, j] - full[:, i])) for n, (i, j) in names.items()}}
    out["median_cycles"]["whole iteration"] = int(np.median(full[1:, 0] - full[:-1, 0]))
    print(json.dumps(out, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
