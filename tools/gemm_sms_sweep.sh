#!/bin/bash
# headline step with the persistent GEMM restricted to n SMs (ABY3CU_GEMM_SMS): does freeing SMs for the other parties'
# keystream / pre-pass kernels pay under the power limit?
set -u
LIST=${1:-148 140 132 124 116 108}
for S in $LIST; do
  ABY3CU_GEMM_SMS=$S timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-linreg --no-logistic --no-basic --no-c1 --no-strong \
      > gpurun_out/sms_$S.log 2> gpurun_out/sms_$S.err || tail -3 gpurun_out/sms_$S.err
  python - <<PY
import json
for line in open("gpurun_out/sms_$S.log"):
    if line.startswith("{"):
        d = json.loads(line)
        print("GEMM_SMS=$S step %.3f ms  e2e %.3f ms  gemm launch %.3f ms" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_launch"]))
PY
done
