#!/bin/bash
# tests of the kernels / facade + the logistic-inference leg of bench.py (ring GEMV on and off)
ABY3_RING_GEMV=1 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_sh3.py -m gpu -x -q 2>&1 | tail -4
for R in 1 0; do export ABY3_RING_GEMV=$R;
  ABY3_RING_GEMV=$R timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-linreg --no-basic --no-c1 --no-strong > gpurun_out/ring_$R.log 2> gpurun_out/ring_$R.err || tail -3 gpurun_out/ring_$R.err
  python - <<PY
import json
for line in open("gpurun_out/ring_$R.log"):
    if line.startswith("{"):
        d = json.loads(line)
        l = d["logistic_inference"]
        print("ABY3_RING_GEMV=$R logistic %.3f ms/pass, %s launches, correct %s; headline step %.3f ms" % (l["ms_per_pass"], l["kernel_launches_per_pass"], l["output_matches_plain_piecewise"], d["ms_per_step"]))
PY
done
