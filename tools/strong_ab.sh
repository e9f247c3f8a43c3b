#!/bin/bash
# strong-scaling leg of bench.py with the broadcast pipelined one step ahead (default) and in line (ABY3_BENCH_STRONG_PIPELINE=0)
# usage (through gpurun --gpus N): bash tools/strong_ab.sh "2 4" r2
set -u
R=${2:-r2}
for N in $1; do
  for P in 1 0; do
    ABY3_BENCH_STRONG_PIPELINE=$P timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-linreg --no-logistic --no-basic --no-c1 --no-distributed \
        > gpurun_out/${R}_strong_g${N}_p${P}.log 2> gpurun_out/${R}_strong_g${N}_p${P}.err || tail -5 gpurun_out/${R}_strong_g${N}_p${P}.err
    python - <<PY
import json
for line in open("gpurun_out/${R}_strong_g${N}_p${P}.log"):
    if line.startswith("{"):
        d = json.loads(line)
        s = d.get("strong") or {}
        print("N=$N pipelined=$P value %.3e ms/step %.3f | strong ms %.3f product %.3f bcast %s err %s | e2e ms %.2f" % (d["value"], d["ms_per_step"], s.get("ms_per_step", -1), s.get("product_only_ms", -1), s.get("bcast_ms"), s.get("max_abs_err_ulp_vs_plain"), d["e2e"]["ms_per_step"]))
PY
  done
done
