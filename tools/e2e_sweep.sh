#!/bin/bash
# end-to-end leg of bench.py with 1 / 2 / 4 row blocks per step (ABY3_BENCH_ROW_BLOCKS)
for NB in ${1:-1 2 4}; do
  ABY3_BENCH_ROW_BLOCKS=$NB timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-linreg --no-logistic --no-basic --no-c1 --no-strong \
      > gpurun_out/e2e_nb$NB.log 2> gpurun_out/e2e_nb$NB.err || tail -3 gpurun_out/e2e_nb$NB.err
  python - <<PY
import json
for line in open("gpurun_out/e2e_nb$NB.log"):
    if line.startswith("{"):
        d = json.loads(line)
        e = d["e2e"]
        print("row blocks $NB: step %.3f ms  e2e %.3f ms (%s; single call %.3f ms)" % (d["ms_per_step"], e["ms_per_step"], e["variant"], e["single_call_ms_per_step"]))
PY
done
