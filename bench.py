#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json):

  metric   3PC si64/sf64 matrix multiplication, ring-MAC/s, device-timed
  workload sf64Matrix<D16> fixed-point matmul WITH truncation, 4096 x 4096 x 4096
           (BASELINE.json configs[1]); one "step" = one complete three-party
           product: every party's cross-term GEMM (tcgen05 limb GEMM), truncation
           pair, the open of xy-r to parties 0/1, and the open-and-truncate pass,
           driven through the sh3 facade (eval.asyncMul(rt, A, B, C, shift).get()
           on three party threads, one stream each, all on this rank's GPU).
  N > 1    weak scaling: every rank owns an independent 4096-row block of a
           (4096*N) x 4096 x 4096 product (B replicated); no data-path collective.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`--impl reference` times the reference's CPU algorithm (the oracle port, all host
threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

# before anything creates the CUDA context (torch included): see aby3_b200/__init__.py
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZE = 4096
SHIFT = 16          # sf64<D16>
METRIC = "3pc_sf64_matmul_trunc_ring_mac_per_s"
UNIT = "ring-MAC/s"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def synth_inputs(M, K, N, seed):
    """Synthetic sf64<D16> operands, uniform in [-4, 4).  (The reference's 4x4 unit-test recipe
    `(u32 >> 8) / 100.0`, aby3_tests/Sh3EvaluatorTests.cpp:469-470, overflows 64 bits once K = 4096
    products are summed at D16, and the truncation protocol needs |x*y| << 2^63; the kernels' run
    time does not depend on the values.)"""
    # a: this rank's row block (seeded per rank); b: the same on every rank (B is replicated)
    a = (np.random.default_rng(seed).uniform(-4.0, 4.0, (M, K)) * (1 << SHIFT)).astype(np.int64)
    b = (np.random.default_rng(999).uniform(-4.0, 4.0, (K, N)) * (1 << SHIFT)).astype(np.int64)
    return a, b


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md): an NVML
    polling thread in this process (a sample every ~2 ms; nvidia-smi -lms takes longer to start
    than a short run lasts and is only the fallback)."""

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.stop_flag, self.th, self.smi = gpu_index, [], False, None, None

    def _nvml_loop(self, nv, h):
        bad = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
               "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append((float(sm), pw, [k for k, v in bad.items() if mask & v]))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.th.start()
        except Exception:
            self.th = None
            self._start_smi()

    def _start_smi(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.smi = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                         "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.smi = None

    def mark(self):
        """Start of the timed region: earlier samples (warm-up) are dropped."""
        self.rows_from = len(self.rows)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml"}
        if self.th is not None:
            self.stop_flag = True
            self.th.join(timeout=2)
            rows = self.rows[getattr(self, "rows_from", 0):] or self.rows
            out["sm_max_mhz"] = self.sm_max
        elif self.smi is not None:
            out["source"] = "nvidia-smi"
            try:
                self.smi.terminate()
                self.smi.wait(timeout=5)
            except Exception:
                pass
            rows = []
            try:
                for l in open(self.path):
                    r = l.strip().split(", ")
                    if len(r) >= 8:
                        out["sm_max_mhz"] = float(r[2])
                        rows.append((float(r[1]), float(r[3]), [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                                                    "sw_power_cap"), r[4:8]) if v.strip().lower().startswith("active")]))
                os.unlink(self.path)
            except Exception:
                pass
        else:
            return out
        if rows:
            sm = np.array([r[0] for r in rows])
            pw = np.array([r[1] for r in rows])
            out["sm_mhz"] = float(np.median(sm))
            out["sm_min_mhz"] = float(np.min(sm))
            out["power_w_max"] = float(np.max(pw))
            out["reasons"] = sorted({x for r in rows for x in r[2]})
            out["samples"] = len(rows)
        return out


def _port_baseline(nthreads, rows, steps=1, warmup=0):
    """The oracle port of the reference algorithm on a bounded sample: the first `rows` rows of A of the
    4096^3 workload -> (rows x 4096) * (4096 x 4096), truncation included."""
    import oracle_lib as o
    a, b = synth_inputs(rows, SIZE, SIZE, 7)
    s = o.Session()
    A, B = s.share_int(0, a), s.share_int(0, b)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        s.mul_trunc(A, B, SHIFT, mode=0, nthreads=nthreads)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    return {"value": rows * SIZE * SIZE / t, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": "first %d rows of A: (%dx%d)*(%dx%d) with truncation, %d step(s), %.2f s/step" % (rows, rows, SIZE, SIZE, SIZE, steps, t)}, t


def _ref_baseline(replicas, rows, steps=1, warmup=0):
    """The REFERENCE's own Sh3Evaluator::asyncMul(A, B, C, shift) (oracle/_ref: its sources compiled from
    /root/reference against the stand-in third-party headers of oracle/shim; the Eigen stand-in's int64
    product is a blocked scalar/AVX2 loop).  One three-party instance = three threads, one per party, as the
    reference runs (frontend/aby3Tutorial.cpp:393-394); `replicas` independent instances run side by side, each
    on its own `rows`-row block of A (the reference's own way to use more cores, BuildingBlocks.cpp:111-147)."""
    import threading
    import oracle_lib as o
    import ref_lib as r
    e, v = o.default_seeds()
    sessions = [r.Session(e, v) for _ in range(replicas)]
    times = []
    for i in range(warmup + steps):
        out = [None] * replicas

        def work(k):
            out[k] = sessions[k].time_mul_trunc(rows, SIZE, SIZE, SHIFT, 1)

        th = [threading.Thread(target=work, args=(k,)) for k in range(replicas)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    for s in sessions:
        s.close()
    t = float(np.mean(times))
    return {"value": replicas * rows * SIZE * SIZE / t, "unit": UNIT, "cores": 3 * replicas, "kind": "reference",
            "sample": "%d x (%dx%d)*(%dx%d) row blocks with truncation through the reference's Sh3Evaluator::asyncMul (3 threads each), "
                      "%d step(s), %.2f s/step" % (replicas, rows, SIZE, SIZE, SIZE, steps, t)}, t


def cpu_baseline(replicas, rows, steps=1, warmup=0):
    """CPU baseline on the host cores: the reference's own code when oracle/_ref exists, else the oracle port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import ref_lib as r
        have_ref = r.available()
    except Exception:
        have_ref = False
    if have_ref:
        return _ref_baseline(replicas, rows, steps, warmup)
    return _port_baseline(3 * replicas, rows, steps, warmup)


def _ref_session():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as o
    import ref_lib as r
    if not r.available():
        return None, None, None
    e, v = o.default_seeds()
    return r, o, r.Session(e, v)


def cpu_reference_side_figures(which):
    """CPU figures for the side workloads (C1, C3, C4, C5), each from the reference's OWN code in
    oracle/_ref on a bounded sample, 3 threads = one per party, as the reference runs (frontend/aby3Tutorial.cpp:393-394).
    NB this fork's matrix asyncMul computes the three Eigen products and then overwrites the result element-wise
    (Sh3Evaluator.cpp:662-668), so its regression does the GEMM-sized work without learning the model; C1 (the plain
    si64 matrix product, which the fork's asyncMul does not compute at all, :96-105) is timed on the oracle port."""
    out = {}
    try:
        r, o, rs = _ref_session()
    except Exception as e:
        return {"error": str(e)}
    if rs is None:
        return {"error": "oracle/_ref not available"}
    rng = np.random.default_rng(3)
    try:
        if "linreg" in which:
            N, F, B, iters = 1 << 13, 1024, 128, 150
            x = rng.normal(1.0, 1.0, (N, F))
            y = x[:, :10] @ rng.integers(0, 10, 10).astype(np.float64)
            t, _ = r.sgd_linear(x, y, B, iters, lr=2.0 ** -10, want_shares=False)
            out["linreg"] = {"value": iters / t, "unit": "iters/s", "cores": 3, "kind": "reference",
                             "sample": "aby3-ML aby3ML engine + Regression.h SGD_Linear, B=%d F=%d, %d iterations over 2^13 samples (%.2f s)" % (B, F, iters, t)}
        if "c1" in which:
            n = 1024
            a = rng.integers(-2**63, 2**63, (n, n), dtype=np.int64)
            b = rng.integers(-2**63, 2**63, (n, n), dtype=np.int64)
            os_ = o.Session()
            A, Bm = os_.share_int(0, a), os_.share_int(1, b)
            t0 = time.perf_counter()
            os_.mul(A, Bm, nthreads=3)
            t = time.perf_counter() - t0
            os_.close()
            out["c1"] = {"value": n ** 3 / t, "unit": "ring-MAC/s", "cores": 3, "kind": "port",
                         "sample": "si64Matrix 1024^3 matrix product, full run on the oracle port (%.2f s)" % t}
        if "logistic" in which:
            rows, F = 8192, 512
            t = rs.time_logistic(rows, F)
            out["logistic"] = {"value": rows / t, "unit": "rows/s", "cores": 3, "kind": "reference",
                               "sample": "asyncMul(X %dx%d, W, D16) + Sh3Piecewise::eval (aby3ML::logisticFunc) on %d rows (%.3f s)" % (rows, F, rows, t)}
        if "basic" in which:
            n = 1 << 17
            a = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
            b = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
            _, t_gt = rs.cipher_gt(rs.share_int(0, a), rs.share_int(0, b))
            h = 1 << 14
            d1, d2 = np.sort(a[:h, 0]).reshape(-1, 1), np.sort(b[:h, 0]).reshape(-1, 1)
            _, t_m = rs.odd_even_merge(rs.share_bin(0, d1), rs.share_bin(0, d2))
            out["gt"] = {"value": n / t_gt, "unit": "elements/s", "cores": 3, "kind": "reference",
                         "sample": "aby3-Basic cipher_gt on 2^17 shared int64 (%.2f s)" % t_gt}
            out["merge"] = {"value": 2 * h / t_m, "unit": "elements/s", "cores": 3, "kind": "reference",
                            "sample": "aby3-Basic odd_even_merge of 2 x 2^14 keys (%.2f s); cost per element grows with log2(n)" % t_m}
    except Exception as e:
        out["error"] = str(e)
    rs.close()
    return out


def c1_bench(sess, n=1024, reps=20):
    """BASELINE configs[0]: si64Matrix 3PC matmul 1024^3 (zero-share form, no truncation), device-timed."""
    rng = np.random.default_rng(19)
    a = rng.integers(-2**63, 2**63, (n, n), dtype=np.int64)
    b = rng.integers(-2**63, 2**63, (n, n), dtype=np.int64)
    A, B = sess.share_int(0, a), sess.share_int(1, b)
    C = sess.mul(A, B)
    ok = bool(np.array_equal(sess.reveal(C, 0)[:4], (a[:4].view(np.uint64) @ b.view(np.uint64)).view(np.int64)))
    for _ in range(3):
        sess.mul(A, B, out=C)
    sess.sync()
    l0 = sess.launches
    sess.timer_begin()
    for _ in range(reps):
        sess.mul(A, B, out=C)
    ms = sess.timer_end() / reps
    out = {"ring_mac_per_s": n ** 3 / (ms * 1e-3), "ms_per_product": ms, "kernel_launches_per_product": (sess.launches - l0) / reps,
           "reveal_matches_plain": ok, "timing": "CUDA events across the three party streams"}
    for h in (A, B, C):
        sess.free(h)
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 3
    replicas = max(1, min(cores // 3, 16))          # bounded: every instance holds its own 0.8 GB of B shares
    rows = 64
    cb, t = cpu_baseline(replicas, rows, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": "sf64<D16> 3PC matmul+truncation 4096x4096x4096 (CPU: bounded sample, %s)" % cb["sample"],
                       "host_cores": cores,
                       "note": "reference sources compiled from /root/reference with stand-in third-party headers (oracle/shim)"
                               if cb["kind"] == "reference" else "oracle port of the reference algorithm (oracle/_ref not available)"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def measured_int8_peak(device):
    """Dense INT8 tensor rate of THIS GPU (not in MEASURED_PEAKS.json, measure it on the box): the library
    int8 GEMM (torch._int_mm -> cuBLASLt, s8 x s8 -> s32) at 8192^3, best of 10, CUDA events.  TOP/s, or None."""
    try:
        import torch
        dev = torch.device("cuda", device)
        n = 8192
        a = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=dev)
        b = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=dev)
        for _ in range(3):
            torch._int_mm(a, b)
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def profile_counters(kernel="gemm_tc"):
    """dram traffic and tensor-pipe activity of the dominant kernel, read at run time from the newest tracked
    `profiles/r*_<kernel>_raw.csv` (one launch under `ncu --set full`, `ncu --page raw --csv`): row 0 = metric names,
    row 1 = units, row 2 = values."""
    import csv
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_%s_raw.csv" % kernel)),
                   key=lambda f: int(re.search(r"r(\d+)_", os.path.basename(f)).group(1)))
    if not files:
        return None
    path = files[-1]
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        return None
    names, units, vals = rows[0], rows[1], rows[-1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "%": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3}

    def get(name):
        if name not in names:
            return None
        i = names.index(name)
        try:
            return float(vals[i].replace(",", "")) * scale.get(units[i], 1.0)
        except ValueError:
            return None

    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    return {"file": os.path.relpath(path, ROOT), "kernel": vals[names.index("Kernel Name")] if "Kernel Name" in names else None,
            "traffic": (rd + wr) if rd is not None and wr is not None else None,
            "tensor_pipe_active_pct": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "tensor_pipe_elapsed_pct": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "ms_under_ncu": get("gpu__time_duration.sum")}


def gemm_roofline(device, ms_per_step=None):
    """Dominant kernel: k_gemm_tc (tcgen05 u8-limb GEMM).  Algorithmic work per launch =
    144*M*N*K int8 ops (36 limb pairs x 2 products x 2 ops); timed with CUDA events on the
    launching stream, inputs (4 x 128 MiB + 512 MiB limb planes) larger than L2."""
    from aby3_b200 import abi
    lib = abi.lib
    ctx = abi.Ctx(device)
    n = SIZE * SIZE
    bufs = [ctx.alloc(8 * n) for _ in range(5)]
    for i, b in enumerate(bufs):
        abi.check(lib.aby3cu_aes_ctr_fill(ctx.h, bytes([i + 1] * 16), 0, b.p, 8 * n))
    p = [b.p for b in bufs]
    ms_all = []
    for it in range(6):
        abi.check(lib.aby3cu_gemm_cross(ctx.h, abi.GEMM_TCGEN05, p[0], p[1], p[2], p[3], SIZE, SIZE, SIZE, p[4], 1))
        ctx.sync()
        ms = abi.C.c_float(0)
        abi.check(lib.aby3cu_gemm_last_main_kernel_ms(ctx.h, abi.C.byref(ms)))
        if it >= 2:
            ms_all.append(ms.value)
    ctx.close()
    ms = float(np.mean(ms_all))
    ops = 144.0 * SIZE ** 3
    achieved = ops / (ms * 1e-3) / 1e12
    pk = peaks()
    bf16 = pk.get("bf16_tflops")
    twice_bf16 = 2.0 * bf16 if bf16 else 2.0 * 1590.0
    int8 = measured_int8_peak(device)
    # denominator: the larger of the library int8 GEMM measured on this GPU and 2 x the measured dense bf16 rate
    # (kind::i8 issues at twice the bf16 rate); the nominal dense int8 figure is 4500 TOP/s
    peak = max(int8 or 0.0, twice_bf16)
    prof = profile_counters("gemm_tc") or {}
    out = {"bound": "tensor", "kernel": "k_gemm_tc", "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak,
           "traffic": prof.get("traffic"),
           "traffic_source": "%s: dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (read at run time)" % prof.get("file"),
           "algorithmic_bytes_per_launch": 2 * 8 * 2 * SIZE * SIZE + 2 * 8 * SIZE * SIZE,      # A_cat + B_cat limb planes (8 B per element of [A0|A1], [B0+B1;B0]) + C read-modify-write
           "ms_per_launch": ms,
           "peak_cublaslt_int8_8192": int8, "peak_twice_measured_bf16": twice_bf16, "frac_of_nominal_4500": achieved / 4500.0,
           "tensor_pipe_active_pct_ncu": prof.get("tensor_pipe_active_pct"), "tensor_pipe_elapsed_pct_ncu": prof.get("tensor_pipe_elapsed_pct"),
           "note": "int8 TOP/s (144*M*N*K ops per launch); frac is of %s: peak = max(cuBLASLt int8 GEMM measured live on this GPU, 2 x the %s dense "
                   "bf16 rate of MEASURED_PEAKS.json -- kind::i8 issues at twice the bf16 rate), i.e. %.2f of 2x-bf16-measured and %.2f of the 4500 TOP/s "
                   "nominal; a library GEMM as denominator lets a good kernel read a little above 1"
                   % (("measured", "measured", achieved / twice_bf16, achieved / 4500.0) if bf16 else ("fallback", "fallback", achieved / twice_bf16, achieved / 4500.0))}
    sus = pk.get("bf16_tflops_sustained")
    if sus:
        # MEASURED_PEAKS.json's figure for a kernel timed inside a long step: the board's 1000 W limit holds the SM clock near
        # 1.5 GHz under back-to-back dense tensor work (tools/power_probe.py, profiles/r2_power_probe.log: this kernel, 993 W,
        # 1507 MHz, sw_power_cap the only reason active)
        out["peak_twice_measured_bf16_sustained"] = 2.0 * sus
        out["frac_of_twice_bf16_sustained"] = achieved / (2.0 * sus)
    if ms_per_step:
        if sus:
            out["step_frac_of_twice_bf16_sustained"] = 3.0 * ops / (ms_per_step * 1e-3) / 1e12 / (2.0 * sus)
        # what the tensor pipe does over the WHOLE step: three parties' launches over the device-timed step
        out["step_frac"] = 3.0 * ops / (ms_per_step * 1e-3) / 1e12 / peak
        out["step_frac_of_nominal_4500"] = 3.0 * ops / (ms_per_step * 1e-3) / 1e12 / 4500.0
        out["gemm_share_of_step"] = 3.0 * ms / ms_per_step
    return out


def linreg_bench(sess, samples, features=1024, batch=128, iters=300, lr=2.0 ** -10):
    """BASELINE configs[2]: aby3-ML linear regression training (main-linear), batch 128 x 1024
    features, sf64<D16>, lr = 2^-10 (aby3-ML/main-linear.cpp:59, Regression.h:112-184).  One
    iteration = two 3-party truncating products (B x F)*(F x 1) and (F x B)*(B x 1) plus the
    local updates; inputs are shared once, before the timed region.  Wall clock around the
    whole loop (the path is latency-bound: 2 communication rounds per iteration)."""
    rng = np.random.default_rng(11)
    pid, xv = sess.plain(0, samples, features)
    step = 1 << 14
    for r0 in range(0, samples, step):
        xv[r0:r0 + step] = (rng.normal(1.0, 1.0, (min(step, samples - r0), features)) * (1 << SHIFT)).astype(np.int64)
    model = np.zeros((features, 1))
    model[:10, 0] = rng.integers(0, 10, 10)
    yv = ((xv[:, :10].astype(np.float64) / (1 << SHIFT)) @ model[:10] * (1 << SHIFT)).astype(np.int64)
    X = sess.share_plain(0, pid, samples, features)
    sess.free(pid)
    Y = sess.share_int(0, yv)
    W = sess.share_int(0, np.zeros((features, 1), dtype=np.int64))
    idx = rng.integers(0, samples, iters * batch).astype(np.uint64)
    # (1) the reference's loop over the sh3 facade: three party threads, one kernel launch at a time
    sess.linreg(X, Y, W, idx[:20 * batch], 20, batch, lr)          # warm-up
    sess.sync()
    l0 = sess.launches
    t0 = time.perf_counter()
    sess.linreg(X, Y, W, idx, iters, batch, lr)
    sess.sync()
    dt_loop = time.perf_counter() - t0
    launches_loop = (sess.launches - l0) / iters
    # (2) the same iteration for co-located parties replayed as ONE CUDA graph per iteration (ml/SgdGraph.h);
    # identical kernels, keystream offsets and results (tests/test_gpu_sh3.py::test_graph_sgd_matches_facade_and_oracle)
    giters = 10 * iters
    gidx = rng.integers(0, samples, giters * batch).astype(np.uint64)
    sess.linreg_graph(X, Y, W, gidx[:20 * batch], 20, batch, lr)   # warm-up
    sess.sync()
    l0 = sess.launches
    t0 = time.perf_counter()
    sess.linreg_graph(X, Y, W, gidx, giters, batch, lr)
    sess.sync()
    dt = time.perf_counter() - t0
    graph_kernels = (sess.launches - l0) / giters
    # (3) the whole run as ONE persistent kernel (csrc/sgd_fused.cu): the grid walks the iterations, two grid barriers per
    # iteration; identical results (tests/test_gpu_sh3.py::test_fused_sgd_matches_facade_and_oracle)
    fiters = 4 * giters
    fidx = rng.integers(0, samples, fiters * batch).astype(np.uint64)
    sess.linreg_fused(X, Y, W, fidx[:20 * batch], 20, batch, lr)   # warm-up
    sess.sync()
    l0 = sess.launches
    fsampler = ClockSampler(0)
    fsampler.start()
    fsampler.mark()
    # three training runs of fiters iterations each (the weights keep training), the median is reported: one run is 80 ms of a
    # latency-bound kernel and moves by +-5 % with the box
    runs = []
    for _ in range(3):
        t0 = time.perf_counter()
        sess.linreg_fused(X, Y, W, fidx, fiters, batch, lr)
        sess.sync()
        runs.append(time.perf_counter() - t0)
    dtf = sorted(runs)[1]
    fclocks = fsampler.stop()
    out = {"iters_per_s": fiters / dtf, "iters": fiters, "batch": batch, "features": features, "samples": samples,
           "decimal": "D16", "lr": lr, "kernel_launches": (sess.launches - l0) // 3, "ms_total": dtf * 1e3, "ms_runs": [round(r * 1e3, 3) for r in runs], "clocks": fclocks,
           "path": "SGD_Linear, three co-located parties, the whole run as one persistent cooperative kernel "
                   "(upload of the batch indices and the final sync included in the time)",
           "graph_replay_iters_per_s": giters / dt, "graph_replay_kernels_per_iter": graph_kernels,
           "facade_loop_iters_per_s": iters / dt_loop, "facade_loop_kernel_launches_per_iter": launches_loop,
           "timing": "host wall clock around the training call, device drained at the end"}
    for h in (X, Y, W):
        sess.free(h)
    return out


def logistic_bench(sess, rows, features=512, reps=3):
    """BASELINE configs[3]: aby3-ML logistic-regression inference with the piecewise sigmoid:
    y = logisticFunc(X * W), X rows x 512, sf64<D16> (aby3-ML/aby3ML.h:102-139).  One pass =
    the truncating 3-party product (skinny GEMV, HBM-bound on X), the region circuit
    (two 64-bit MSB-of-sum adders + one na_And, bit-sliced), two bit x arithmetic products."""
    rng = np.random.default_rng(13)
    pid, xv = sess.plain(0, rows, features)
    try:
        # synthetic rows: one random block of 2^16 rows, repeated (filling 2^22 x 512 values from the generator takes a minute)
        step = min(rows, 1 << 16)
        blk = (rng.uniform(-1.0, 1.0, (step, features)) * (1 << SHIFT)).astype(np.int64)
        for r0 in range(0, rows, step):
            xv[r0:r0 + step] = blk[:min(step, rows - r0)]
        X = sess.share_plain(0, pid, rows, features)
    finally:
        sess.free(pid)
    w = (rng.uniform(-0.1, 0.1, (features, 1)) * (1 << SHIFT)).astype(np.int64)
    W = sess.share_int(0, w)
    try:
        return _logistic_passes(sess, X, W, rows, features, reps)
    finally:
        for h in (X, W):
            sess.free(h)


def _logistic_passes(sess, X, W, rows, features, reps):
    th, coef = [-0.5, 0.5], [[], [0.5, 1], [1]]

    def once():
        z = sess.mul(X, W, shift=SHIFT)
        y = sess.piecewise(z, th, coef, SHIFT)
        return z, y

    z, y = once()
    # sanity on a sample of rows: exact piecewise function of the revealed linear part
    zr, yr = sess.reveal(z, 0), sess.reveal(y, 0)
    exp = np.where(zr < -(1 << (SHIFT - 1)), 0, np.where(zr < (1 << (SHIFT - 1)), zr + (1 << (SHIFT - 1)), 1 << SHIFT))
    ok = bool(np.array_equal(yr, exp))
    for h in (z, y):
        sess.free(h)
    sess.sync()
    l0 = sess.launches
    sess.timer_begin()
    t0 = time.perf_counter()
    for _ in range(reps):
        z, y = once()
        sess.free(z)
        sess.free(y)
    ms = sess.timer_end()
    wall = time.perf_counter() - t0
    out = {"rows_per_s": rows * reps / (ms * 1e-3), "rows": rows, "features": features, "ms_per_pass": ms / reps,
           "wall_ms_per_pass": wall * 1e3 / reps, "kernel_launches_per_pass": (sess.launches - l0) / reps,
           "output_matches_plain_piecewise": ok, "decimal": "D16",
           "timing": "CUDA events across the three party streams"}
    return out


def basic_bench(sess, n):
    """BASELINE configs[4]: aby3-Basic greater-than and odd-even merge over secret-shared int64
    elements (Sh3BinaryEvaluator AND-layer throughput).  (i) cipher_gt on two vectors of n elements
    (MSB-of-sum circuit, 64-bit prefix adder); (ii) odd_even_merge of two sorted runs of n/2
    elements: ceil(log2(n/2)) + 1 compare-exchange stages, each one int_int_lt(64) circuit plus two
    bitwiseAnd(64) circuits (BoolBasic.cpp:275-312, Sort.cpp:327-406)."""
    rng = np.random.default_rng(17)
    a = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
    b = rng.integers(-2**62, 2**62, (n, 1), dtype=np.int64)
    A, B = sess.share_int(0, a), sess.share_int(0, b)
    gt = sess.cipher_gt(A, B)
    ok_gt = bool(np.array_equal(sess.reveal(gt, 0, binary=True) & 1, (a > b).astype(np.int64)))
    sess.free(gt)
    sess.sync()
    l0 = sess.launches
    sess.timer_begin()
    gt = sess.cipher_gt(A, B)
    ms_gt = sess.timer_end()
    l_gt = sess.launches - l0
    for h in (gt, A, B):
        sess.free(h)
    half = n // 2
    d1 = np.sort(a[:half, 0]).reshape(-1, 1)
    d2 = np.sort(b[:half, 0]).reshape(-1, 1)
    D1, D2 = sess.share_bin(0, d1, 64), sess.share_bin(0, d2, 64)
    for _ in range(3):                                # warm-up: the stages' buffer sizes enter the pool (the second and third pass still allocate a few)
        sess.free(sess.odd_even_merge(D1, D2))
    sess.sync()
    # three timed merges: the median is reported, the spread says whether the buffer pool has settled
    runs = []
    for rep in range(3):
        l0 = sess.launches
        sess.timer_begin()
        t0 = time.perf_counter()
        m = sess.odd_even_merge(D1, D2)
        ms_rep = sess.timer_end()
        runs.append((ms_rep, (time.perf_counter() - t0) * 1e3, sess.launches - l0))
        if rep < 2:
            sess.free(m)
    ms_all = sorted(r[0] for r in runs)
    ms_merge, wall, l_merge = ms_all[1], sorted(r[1] for r in runs)[1] / 1e3, runs[-1][2]
    merged = sess.reveal(m, 0, binary=True).reshape(-1)
    ok_merge = bool(np.all(np.diff(merged) >= 0)) and merged.size == 2 * half
    for h in (m, D1, D2):
        sess.free(h)
    return {"elements": n, "gt_elements_per_s": n / (ms_gt * 1e-3), "gt_ms": ms_gt, "gt_kernel_launches": l_gt, "gt_correct": ok_gt,
            "merge_elements_per_s": 2 * half / (ms_merge * 1e-3), "merge_ms": ms_merge, "merge_wall_ms": wall * 1e3,
            "merge_kernel_launches": l_merge, "merge_sorted": ok_merge, "merge_ms_runs": [round(x, 3) for x in ms_all],
            "merge_spread": (ms_all[-1] - ms_all[0]) / ms_all[1],
            "timing": "CUDA events across the three party streams"}


def pcie_probe(device, nbytes=256 << 20):
    """Raw page-locked h2d / d2h rate of this rank's GPU (GB/s), both directions one after the other; at N > 1 every rank
    runs it at the same time (barrier before), so the number is the per-rank share of the host's PCIe / DRAM bandwidth."""
    from aby3_b200 import abi
    ctx = abi.Ctx(device)
    hp = abi.C.c_void_p()
    abi.check(abi.lib.aby3cu_host_alloc(abi.C.byref(hp), nbytes))
    d = ctx.alloc(nbytes)
    out = {}
    try:
        for name, fn, args in (("h2d", abi.lib.aby3cu_h2d, (d.p, hp)), ("d2h", abi.lib.aby3cu_d2h, (hp, d.p))):
            abi.check(fn(ctx.h, *args, nbytes))
            ctx.sync()
            t0 = time.perf_counter()
            for _ in range(3):
                abi.check(fn(ctx.h, *args, nbytes))
            ctx.sync()
            out[name + "_GBps"] = 3 * nbytes / (time.perf_counter() - t0) / 1e9
    finally:
        abi.lib.aby3cu_host_free(hp)
        ctx.close()
    return out


class _DevArray:
    """a raw device pointer as a torch-importable object (torch.as_tensor reads __cuda_array_interface__)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}


def strong_scaling_leg(sess, dist, rank, world, local, steps, warmup, rows_total=8192):
    """STRONG scaling: ONE fixed sf64<D16> (8192 x 4096) * (4096 x 4096) product with truncation, output rows sharded over
    the ranks.  B is produced (shared) on rank 0 only; inside every timed step its share planes are broadcast
    from rank 0 over NVLink by NCCL (torch.distributed.broadcast issued on party 0's own CUDA stream, so there is no host
    synchronisation): the three DISTINCT planes x_0, x_1, x_2 travel (3 x 128 MiB), each rank fills the replicated copies
    (plane 1 of party p+1 = plane 0 of party p) by a device copy.  Then every rank multiplies its row block."""
    from aby3_b200 import distutil
    K = N = SIZE
    r0, r1 = distutil.row_block(rows_total, rank, world)
    rows = r1 - r0
    a = (np.random.default_rng(5000 + rank).uniform(-4.0, 4.0, (rows, K)) * (1 << SHIFT)).astype(np.int64)
    A = sess.share_int(0, a)
    plane_bytes = 8 * K * N
    if rank == 0:
        _, b = synth_inputs(1, K, N, 0)
        B = sess.share_int(0, b)
    else:
        B = sess.alloc_shares(K, N)
    bcast = None
    Bs = [B, B]
    pipelined = dist is not None and os.environ.get("ABY3_BENCH_STRONG_PIPELINE", "1") != "0"
    if dist is not None:
        import torch
        dev = torch.device("cuda", local)
        if pipelined and rank != 0:
            Bs = [B, sess.alloc_shares(K, N)]            # the receiving ranks double-buffer B: step k+1's planes land while step k multiplies
        t = []
        for Bh in (Bs if Bs[1] != Bs[0] else Bs[:1]):
            ptrs = sess.device_ptrs(Bh)
            t.append([[torch.as_tensor(_DevArray(ptrs[p][pl], K * N), device="cuda") for pl in range(2)] for p in range(3)])
        if len(t) == 1:
            t.append(t[0])
        streams = [torch.cuda.ExternalStream(sess.party_stream(p), device=dev) for p in range(3)]
        comm = torch.cuda.Stream(device=dev)

        def bcast(slot=0, on=None, after=None):
            """B's three distinct planes from rank 0 into buffer `slot`, replicas filled by device copies, on stream `on`
            (default: party 0's) once the events `after` (the last readers of that buffer) have passed -> completion event"""
            st = on if on is not None else streams[0]
            if after is None:
                after = []
                for q in (1, 2):         # parties 1 and 2 may still be reading the previous step's planes
                    e = torch.cuda.Event()
                    e.record(streams[q])
                    after.append(e)
            for e in after:
                st.wait_event(e)
            tt = t[slot]
            with torch.cuda.stream(st):
                for p in range(3):
                    dist.broadcast(tt[p][0], src=0)
                if rank != 0:
                    for p in range(3):
                        tt[(p + 1) % 3][1].copy_(tt[p][0])
                ev = torch.cuda.Event()
                ev.record(st)
            return ev

    state = {"k": 0, "ready": [None, None], "done": [None, None]}

    def step(C):
        if bcast is None:
            return sess.mul(A, B, shift=SHIFT, out=C)
        if not pipelined:
            ev = bcast()
            streams[1].wait_event(ev)
            streams[2].wait_event(ev)
            return sess.mul(A, B, shift=SHIFT, out=C)
        # pipelined one step ahead: step k multiplies buffer k % 2 while the broadcast of step k + 1 fills the other one on a
        # communication stream (every step still broadcasts B once; the planes of step 0 are broadcast by the call before)
        k = state["k"]
        cur, nxt = k % 2, (k + 1) % 2
        if state["ready"][cur] is None:
            state["ready"][cur] = bcast(cur, on=comm, after=[])
        for q in range(3):
            streams[q].wait_event(state["ready"][cur])
        out = sess.mul(A, Bs[cur], shift=SHIFT, out=C)
        done = []
        for q in range(3):
            e = torch.cuda.Event()
            e.record(streams[q])
            done.append(e)
        state["done"][cur] = done
        state["ready"][nxt] = bcast(nxt, on=comm, after=state["done"][nxt] or [])
        state["k"] = k + 1
        return out

    def drain():
        # the broadcast issued by the last step belongs to the timed region too
        if bcast is not None and pipelined and state["ready"][state["k"] % 2] is not None:
            streams[0].wait_event(state["ready"][state["k"] % 2])

    C = step(0)
    for _ in range(max(warmup, 3) - 1):
        step(C)
    drain()
    sess.sync()
    if dist is not None:
        dist.barrier()
    sess.timer_begin()
    for _ in range(steps):
        step(C)
    drain()
    ms = sess.timer_end()
    sess.sync()
    if bcast is not None:
        torch.cuda.synchronize(dev)
    # where the time goes: the broadcast alone and the row-block product alone
    ms_b = None
    if bcast is not None:
        dist.barrier()
        sess.timer_begin()
        for _ in range(steps):
            bcast()
        ms_b = sess.timer_end() / steps
        sess.sync()
        dist.barrier()
    sess.timer_begin()
    for _ in range(steps):
        sess.mul(A, Bs[0], shift=SHIFT, out=C)
    ms_c = sess.timer_end() / steps
    sess.sync()
    # the product of the broadcast planes is the right one: reveal a few rows on every rank
    c = sess.reveal(C, 0)
    _, b = synth_inputs(1, K, N, 0)
    err = int(np.max(np.abs(c[:4] - ((a[:4] @ b) >> SHIFT))))
    for h in {A, Bs[0], Bs[1], C}:
        sess.free(h)
    ms_max, _, _ = distutil.combine(dist, "cuda", ms, 0, 0.0)
    ms_c_max, _, _ = distutil.combine(dist, "cuda", ms_c, 0, 0.0)
    ms_b_max = distutil.combine(dist, "cuda", ms_b, 0, 0.0)[0] if ms_b is not None else None
    err_max = int(distutil.combine(dist, "cuda", float(err), 0, 0.0)[0])
    if err_max > 4:
        raise SystemExit("bench: strong-scaling product is off by %d ulp" % err_max)
    per_step = ms_max / steps
    out = {"value": float(rows_total) * K * N / (per_step * 1e-3), "unit": UNIT, "scaling": "strong", "ms_per_step": per_step,
           "workload": "sf64<D16> (%d x %d) * (%d x %d) with truncation, output rows sharded over %d GPU(s), %d rows per GPU" % (rows_total, K, K, N, world, rows),
           "product_only_ms": ms_c_max, "max_abs_err_ulp_vs_plain": err_max,
           "timing": "CUDA events across the three party streams, broadcast inside the timed region, max over ranks"}
    if ms_b_max is not None:
        out.update({"bcast_ms": ms_b_max, "bcast_bytes_per_step": 3 * plane_bytes,
                    "bcast_GBps": 3 * plane_bytes / (ms_b_max * 1e-3) / 1e9,
                    "bcast": "torch.distributed.broadcast (NCCL) of B's three distinct share planes from rank 0; replicas filled by device copies; "
                             + ("issued one step ahead on a communication stream into the other of two B buffers (every step broadcasts once, all of them inside the timed region)"
                                if pipelined else "on party 0's stream before the product"),
                    "pipelined": bool(pipelined),
                    "limiter": "broadcast of B alone %.2f ms, row-block product alone %.2f ms, step %.2f ms: %.2f ms of the broadcast is not hidden (NCCL's kernel only gets SMs between the persistent GEMM launches)"
                               % (ms_b_max, ms_c_max, per_step, max(0.0, per_step - ms_c_max))})
    return out


def distributed_leg(steps, warmup, n=SIZE):
    """DISTRIBUTED placement (north_star: "the three-party ring reshare moves by NCCL send/recv over NVLink when
    parties sit on different GPUs"): party p on GPU p, one process (this rank) driving the three party threads, every
    message between parties an ncclSend / ncclRecv pair on the parties' streams (one ncclGroup per protocol step;
    reference sends: Sh3Evaluator.cpp:109-110, 681-684).  Same 4096^3 truncating product as the headline."""
    from aby3_b200 import harness
    s = harness.Session(devices=(0, 1, 2), transport="nccl")
    a, b = synth_inputs(n, n, n, 1234)
    A, B = s.share_int(0, a), s.share_int(0, b)
    C = s.mul(A, B, shift=SHIFT)
    for _ in range(max(warmup, 3) - 1):
        s.mul(A, B, shift=SHIFT, out=C)
    s.sync()
    sent0 = s.bytes_sent
    t0 = time.perf_counter()
    for _ in range(steps):
        s.mul(A, B, shift=SHIFT, out=C)
    s.sync()
    dt = (time.perf_counter() - t0) / steps
    sent = (s.bytes_sent - sent0) / steps
    c = s.reveal(C, 0)
    err = int(np.max(np.abs(c[:8] - ((a[:8] @ b) >> SHIFT))))
    s.close()
    if err > 4:
        raise RuntimeError("distributed placement: product off by %d ulp" % err)
    return {"value": float(n) ** 3 / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "devices": [0, 1, 2], "transport": "nccl send/recv (ring of 3 ranks, ncclCommInitAll)",
            "bytes_between_parties_per_step": sent, "nvlink_GBps_aggregate": sent / dt / 1e9, "max_abs_err_ulp_vs_plain": err,
            "messages_per_step": "4 x %d MiB: the opened xy - r from P0->P1, P1->P0, P2->P0, P2->P1 (Sh3Evaluator.cpp:681-684)" % (8 * n * n >> 20),
            "timing": "host wall clock over the steps with the three devices drained on both sides (the parties' streams live on three devices)"}


class E2EPipe:
    """The end-to-end step through the public calls (localIntMatrix / asyncMul(..., shift) / revealAll) of ONE Session with HOST
    buffers: two sets of page-locked input / output buffers, uploads prefetched on the owner's copy stream, downloads on a
    second copy stream, consecutive steps pipelined two deep."""

    def __init__(self, sess, a, b, nb):
        self.sess, self.nb = sess, nb
        self.M, self.K = a.shape
        self.N = b.shape[1]
        self.rb = self.M // nb
        self.sets = [self._make_set(a, b), self._make_set(a, b)]

    def _make_set(self, a, b):
        sess = self.sess
        st = {"pb": sess.plain(0, self.K, self.N), "pa": [], "pc": []}
        st["pb"][1][...] = b
        for i in range(self.nb):
            pid, view = sess.plain(0, self.rb, self.K)
            view[...] = a[i * self.rb:(i + 1) * self.rb]
            st["pa"].append(pid)
            st["pc"].append(sess.plain(0, self.rb, self.N))
        return st

    def upload(self, st):
        sess = self.sess
        sess.plain_touch(0, st["pb"][0])          # the host buffers count as freshly written: device copies are stale
        for pid in st["pa"]:
            sess.plain_touch(0, pid)
        sess.plain_prefetch(0, st["pb"][0])
        for pid in st["pa"]:
            sess.plain_prefetch(0, pid)

    def compute(self, st):
        sess = self.sess
        hb = sess.share_plain(0, st["pb"][0], self.K, self.N)
        live = [hb]
        for i in range(self.nb):
            ha = sess.share_plain(0, st["pa"][i], self.rb, self.K)
            hc = sess.mul(ha, hb, shift=SHIFT)
            sess.reveal_plain_async(hc, 0, st["pc"][i][0])
            live += [ha, hc]
        return live

    def finish(self, st, live):
        for i in range(self.nb):
            self.sess.plain_wait(0, st["pc"][i][0])
        for h in live:
            self.sess.free(h)

    def run(self, nsteps):
        self.upload(self.sets[0])
        prev = None
        for k in range(nsteps):
            live = self.compute(self.sets[k % 2])
            if prev is not None:
                self.finish(*prev)
            prev = (self.sets[k % 2], live)
            if k + 1 < nsteps:
                self.upload(self.sets[(k + 1) % 2])
        self.finish(*prev)

    def close(self):
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=SIZE, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--linreg-samples", type=int, default=1 << 20)
    ap.add_argument("--no-linreg", action="store_true")
    ap.add_argument("--logistic-rows", type=int, default=1 << 22)
    ap.add_argument("--no-logistic", action="store_true")
    ap.add_argument("--basic-elements", type=int, default=1 << 24)
    ap.add_argument("--no-basic", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-distributed", action="store_true")
    ap.add_argument("--no-c1", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # Page-locked buffers and the party threads should live on the NUMA node this rank's GPU hangs off: with 8 ranks
    # the h2d / d2h legs of the end-to-end number otherwise cross the socket interconnect.
    numa = "unset"
    if world > 1:
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = int(vis.split(",")[local]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else local
            nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(idx))
            numa = "gpu-local cpus (%d)" % len(os.sched_getaffinity(0))
        except Exception as e:
            numa = "unchanged (%s)" % type(e).__name__

    from aby3_b200 import harness
    M = K = N = args.size
    sess = harness.Session(devices=(local, local, local))
    # inputs resident in HBM before the timed region: party 0 shares a and b
    a, b = synth_inputs(M, K, N, 1000 + rank)
    A, B = sess.share_int(0, a), sess.share_int(0, b)
    sampler = ClockSampler(local)
    sampler.start()                      # samples cover warm-up + timed region (same load)
    Cid = sess.mul(A, B, shift=SHIFT)
    for _ in range(max(args.warmup, 3) - 1):
        sess.mul(A, B, shift=SHIFT, out=Cid)
    sess.sync()

    def barrier():
        if dist is not None:
            dist.barrier()

    barrier()
    sess.sync()
    launches0 = sess.launches
    sampler.mark()
    sess.timer_begin()
    for _ in range(args.steps):
        sess.mul(A, B, shift=SHIFT, out=Cid)
    ms = sess.timer_end()
    sess.sync()
    clocks = sampler.stop()
    launches = sess.launches - launches0
    barrier()

    from aby3_b200 import distutil
    step_macs = float(M) * K * N
    ms_max, launches, units_total = distutil.combine(dist, "cuda", ms, launches, step_macs * args.steps)

    # ---- end to end through the public API with HOST buffers (every step: h2d of the
    # plaintext inputs from page-locked memory, share, multiply, reveal, d2h of the result)
    e2e_steps = max(1, min(args.steps, 3))
    pa, va = sess.plain(0, M, K)
    pb, vb = sess.plain(0, K, N)
    pc, out = sess.plain(0, M, N)          # page-locked: the reveal's d2h copy lands here
    va[...] = a
    vb[...] = b

    def e2e_once():
        sess.plain_touch(0, pa)
        sess.plain_touch(0, pb)
        ha = sess.share_plain(0, pa, M, K)
        hb = sess.share_plain(0, pb, K, N)
        hc = sess.mul(ha, hb, shift=SHIFT)
        sess.reveal_plain(hc, 0, pc)
        for h in (ha, hb, hc):
            sess.free(h)

    e2e_once()
    e2e_once()          # the second pass settles the buffer pools (the first one grows them)
    sess.sync()
    barrier()
    pool0 = sess.pool_stats

    def timed(fn):
        # the better of two repetitions of e2e_steps steps: one in a few runs pays a page-locking / allocator hiccup
        best = None
        for _ in range(2):
            sess.sync()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                fn()
            sess.sync()
            dt = (time.perf_counter() - t0) / e2e_steps
            best = dt if best is None else min(best, dt)
        return best

    e2e_single = timed(e2e_once)
    pool1 = sess.pool_stats

    # The same end-to-end step with the transfers overlapped: B is uploaded first, A travels as NB row blocks on the
    # owner's copy stream while the previous block is being multiplied, and each block of the result is downloaded on
    # the copy stream while the next block computes.  Consecutive steps are pipelined two deep: step k+1's uploads are
    # issued (into the other set of page-locked buffers) while step k computes, and step k's download is only awaited
    # after step k+1's kernels have been queued.  Same public calls (localIntMatrix / asyncMul / revealAll), the same
    # bytes over PCIe in EVERY step, all of it inside the timed region.
    NB = int(os.environ.get("ABY3_BENCH_ROW_BLOCKS", "1"))      # row blocks per step (1: one product; 2: 1024 tiles = 6.9 waves of 148 each)
    if NB < 1 or M % NB:
        NB = 1
    pipe = E2EPipe(sess, a, b, NB)
    sets = pipe.sets

    pipe_steps = max(4, min(args.steps, 10))
    pipe.run(3)
    pipe.run(3)
    barrier()
    best = None
    for _ in range(2):
        sess.sync()
        barrier()
        t0 = time.perf_counter()
        pipe.run(pipe_steps)
        sess.sync()
        dt = (time.perf_counter() - t0) / pipe_steps
        best = dt if best is None else min(best, dt)
    e2e_t = best
    streamed_err = max(int(np.max(np.abs(st["pc"][0][1][:8] - ((a[:8] @ b) >> SHIFT)))) for st in sets)
    if streamed_err > 4:
        raise SystemExit("bench: streamed end-to-end product is off by %d ulp" % streamed_err)
    if e2e_single < e2e_t:          # report the better public-API path
        e2e_t, e2e_path, e2e_steps_used = e2e_single, "single call", e2e_steps
    else:
        e2e_path, e2e_steps_used = "transfers on copy streams, steps pipelined two deep (%d row block(s) per step)" % NB, pipe_steps
    barrier()
    pcie = None
    try:
        pcie = pcie_probe(local)
    except Exception as e:
        pcie = {"error": str(e)}
    if dist is not None:
        import torch
        t = torch.tensor([e2e_t], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t.item())
    # sanity: the revealed product equals the plaintext fixed-point product within the protocol's ulp bound
    chk_rows = 8
    ref = (a[:chk_rows] @ b) >> SHIFT               # int64 matmul wraps mod 2^64, like the ring
    max_err = int(np.max(np.abs(out[:chk_rows] - ref)))
    if max_err > 4:
        raise SystemExit("bench: revealed product is off by %d ulp (> 4) from the plaintext product" % max_err)
    # ---- strong scaling: one fixed product, rows sharded over the ranks, B broadcast by NCCL inside the timed region
    strong = None
    sess.trim()
    if not args.no_strong:
        try:
            strong = strong_scaling_leg(sess, dist, rank, world, local, args.steps, args.warmup)
        except SystemExit:
            raise
        except Exception as e:
            strong = {"error": "%s: %s" % (type(e).__name__, e)}
    sess.trim()
    # ---- distributed placement: party p on GPU p, NCCL send/recv ring, driven by rank 0 while the other ranks are
    # parked on a CPU (gloo) barrier so that their GPUs are free
    distributed = None
    if world >= 3 and not args.no_distributed:
        import torch.distributed as tdist
        park = tdist.new_group(backend="gloo")
        sess.sync()
        tdist.barrier(group=park)
        if rank == 0:
            try:
                distributed = distributed_leg(args.steps, args.warmup)
            except Exception as e:
                distributed = {"error": "%s: %s" % (type(e).__name__, e)}
        tdist.barrier(group=park)
    side = [] if rank != 0 else [w for w, off in (("linreg", args.no_linreg), ("c1", args.no_c1), ("logistic", args.no_logistic), ("basic", args.no_basic)) if not off]
    cpu_side = cpu_reference_side_figures(side) if (side and world == 1 and not args.no_cpu_baseline) else {}
    c1 = None
    if rank == 0 and not args.no_c1:
        try:
            c1 = c1_bench(sess)
            c1["cpu_reference"] = cpu_side.get("c1")
        except Exception as e:
            c1 = {"error": str(e)}
    linreg = None
    sess.trim()          # each side workload starts with empty buffer pools
    if rank == 0 and not args.no_linreg:
        try:
            linreg = linreg_bench(sess, args.linreg_samples)
            linreg["cpu_reference"] = cpu_side.get("linreg")
        except Exception as e:
            linreg = {"error": str(e)}
    logistic = None
    sess.trim()
    if rank == 0 and not args.no_logistic:
        # BASELINE configs[3] is 2^22 rows x 512 features (96 GiB of share planes + the 16 GiB plaintext); a smaller GPU
        # memory budget falls back to half of it and says so
        for rows in (args.logistic_rows, args.logistic_rows // 2):
            try:
                logistic = logistic_bench(sess, rows)
                break
            except Exception as e:
                logistic = {"error": "%d rows: %s" % (rows, e)}
                sess.trim()
        logistic["cpu_reference"] = cpu_side.get("logistic")
    basic = None
    sess.trim()
    if rank == 0 and not args.no_basic:
        try:
            basic = basic_bench(sess, args.basic_elements)
            basic["cpu_reference"] = {"gt": cpu_side.get("gt"), "merge": cpu_side.get("merge")}
        except Exception as e:
            basic = {"error": str(e)}
    sess.close()

    if rank == 0:
        value = distutil.throughput(units_total, ms_max)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": "sf64<D16> 3PC matmul+truncation %dx%dx%d per GPU (BASELINE configs[1]); three parties co-located on each GPU" % (M, K, N),
                       "global_rows": M * world, "gemm": "tcgen05 u8-limb (kind::i8), 128x64 tiles, 12 MMA per 32-deep k-step (wide limbs split in equal halves)",
                       "l2": "inputs larger than L2 (per party 4 x %d MiB share planes + %d MiB limb planes)" % (8 * M * K >> 20, 2 * 16 * M * K >> 20),
                       "timing": "CUDA events across the three party streams (fork/join on one start and one end event), max over ranks",
                       "executed_u64_mac_per_s": 6.0 * value, "max_abs_err_ulp_vs_plain": max_err, "cpu_affinity": numa},
            "clocks": clocks,
            "e2e": {"value": world * step_macs / e2e_t, "unit": UNIT, "h2d_bytes_per_step": 8 * (M * K + K * N), "d2h_bytes_per_step": 8 * M * N,
                    "ms_per_step": e2e_t * 1e3, "steps": e2e_steps_used, "repetitions": "better of 2 x %d steps" % e2e_steps_used, "single_call_ms_per_step": e2e_single * 1e3, "variant": e2e_path,
                    "tried": "two protocol instances in flight (a second Session taking every other step from a second host thread): 13.98 vs 13.66 ms/step -- the GPU is bound by the sum of the kernels' work, one instance's sharing phase under the other's GEMMs buys nothing (profiles/r2_e2e_notes.md)",
                    "pcie_GBps_per_step_effective": {"h2d": 8 * (M * K + K * N) / e2e_t / 1e9, "d2h": 8 * M * N / e2e_t / 1e9},
                    "pcie_probe_rank0": pcie,
                    "driver_allocations_during_single_call_steps": int(pool1[0] - pool0[0]),
                    "path": "enc.localIntMatrix(page-locked host a,b) -> eval.asyncMul(..., shift) -> enc.revealAll -> page-locked host c; "
                            "streamed variant: b first, a as %d row blocks prefetched on a copy stream, result blocks downloaded on the copy stream, consecutive steps pipelined two deep over two sets of host buffers" % NB},
            "gpu_launches": launches,
            "strong": strong,
            "distributed": distributed,
            "si64_matmul_1024": c1,
            "linreg": linreg,
            "logistic_inference": logistic,
            "gt_and_merge": basic,
        }
        try:
            line["roofline"] = gemm_roofline(local, ms_max / args.steps)
        except Exception as e:  # keep the headline even if the side measurement fails
            line["roofline"] = {"error": str(e)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb, _ = cpu_baseline(1, 64, steps=1)
                line["cpu_baseline"] = cb
            except Exception as e:
                line["cpu_baseline"] = {"error": str(e)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
