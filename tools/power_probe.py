"""Is the tcgen05 GEMM bound by the tensor pipe or by the board's power limit?

The same `aby3cu_gemm_cross` (4096^3, limb pre-pass included) is timed two ways through the C ABI with CUDA events:
  * spaced    -- one product at a time with an idle gap before it (the GPU is at its maximum clock when the launch starts);
  * sustained -- products back to back for a few seconds (what a step of the headline workload looks like to the board),
while a thread samples SM clock, power draw and the throttle reasons through NVML every few milliseconds.  One JSON line
per mode.  Usage: python tools/power_probe.py [--n 4096] [--seconds 3] [--gap-ms 40]"""
import argparse
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aby3_b200 import abi  # noqa: E402

lib = abi.lib
KA = bytes(range(16))


class Sampler:
    def __init__(self, index=0, period=0.004):
        import pynvml as nv
        self.nv = nv
        nv.nvmlInit()
        self.h = nv.nvmlDeviceGetHandleByIndex(index)
        self.period = period
        self.rows = []
        self.on = False

    def start(self):
        self.rows = []
        self.on = True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def _loop(self):
        nv = self.nv
        while self.on:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self.on = False
        self.t.join()
        if not self.rows:
            return {}
        mhz = sorted(r[0] for r in self.rows)
        w = sorted(r[1] for r in self.rows)
        reasons = 0
        for r in self.rows:
            reasons |= r[2]
        nv = self.nv
        names = [n for n, bit in (("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap), ("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                                  ("sw_thermal", nv.nvmlClocksEventReasonSwThermalSlowdown), ("hw_thermal", nv.nvmlClocksEventReasonHwThermalSlowdown))
                 if reasons & bit]
        return {"samples": len(mhz), "sm_mhz_median": mhz[len(mhz) // 2], "sm_mhz_min": mhz[0], "sm_mhz_max": mhz[-1],
                "power_w_median": round(w[len(w) // 2], 1), "power_w_max": round(w[-1], 1), "reasons": names,
                "power_limit_w": self.nv.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1000.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--gap-ms", type=float, default=40.0)
    args = ap.parse_args()
    g = args.n
    ctx = abi.Ctx(0)
    bufs = [ctx.alloc(8 * g * g) for _ in range(5)]
    for b in bufs:
        abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, KA, 0, b.p, 8 * g * g))
    p = [b.p for b in bufs]

    def product():
        abi.check(lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_TCGEN05, p[0], p[1], p[2], p[3], g, g, g, p[4], 1))

    for _ in range(3):
        product()
    ctx.sync()
    ops = 144.0 * g ** 3
    try:
        smp = Sampler()
    except Exception as e:           # no NVML: timings only
        smp = None
        print(json.dumps({"nvml": "unavailable: %s" % e}), flush=True)

    # ---- spaced: the board idles before every product
    times = []
    a, b = ctx.event(), ctx.event()
    for _ in range(25):
        ctx.sync()
        time.sleep(args.gap_ms / 1e3)
        ctx.record(a)
        product()
        ctx.record(b)
        ctx.sync()
        times.append(abi.elapsed_ms(a, b))
    times.sort()
    ms = times[len(times) // 2]
    print(json.dumps({"mode": "spaced", "gap_ms": args.gap_ms, "n": g, "ms_per_product_median": round(ms, 4), "ms_min": round(times[0], 4),
                      "int8_TOPS": round(ops / ms / 1e9, 1)}), flush=True)

    # ---- sustained: back to back for --seconds, timed in windows of 20 products
    time.sleep(0.5)
    if smp:
        smp.start()
    t_end = time.time() + args.seconds
    windows = []
    while time.time() < t_end:
        ctx.record(a)
        for _ in range(20):
            product()
        ctx.record(b)
        ctx.sync()
        windows.append(abi.elapsed_ms(a, b) / 20)
    clk = smp.stop() if smp else {}
    tail = sorted(windows[len(windows) // 2:])
    ms_s = tail[len(tail) // 2]
    print(json.dumps({"mode": "sustained", "seconds": args.seconds, "n": g, "ms_per_product_first_window": round(windows[0], 4),
                      "ms_per_product_second_half_median": round(ms_s, 4), "int8_TOPS": round(ops / ms_s / 1e9, 1),
                      "slowdown_vs_spaced": round(ms_s / ms, 3), "clocks": clk}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
