#!/bin/bash
# Round-end evidence on one B200 (run through gpurun): GPU tests, the bench line, the ncu launch list of the SAME bench
# command and one full-section capture of the dominant kernel.  Everything lands under gpurun_out/ (copied into profiles/).
# A number printed by a run under ncu is never a bench value.
set -u
R=${1:-r1}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${R}_tests_final.log
python bench.py > gpurun_out/${R}_bench_final.log 2> gpurun_out/${R}_bench_final.err || tail -5 gpurun_out/${R}_bench_final.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-linreg --no-logistic --no-basic > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-linreg --no-logistic --no-basic > gpurun_out/${R}_ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:k_gemm_tc --launch-skip 3 --launch-count 1 -o /tmp/${R}_gemm_tc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-linreg --no-logistic --no-basic > gpurun_out/${R}_ncu_gemm.log 2>&1
ncu -i /tmp/${R}_gemm_tc.ncu-rep --page raw --csv > gpurun_out/${R}_gemm_tc_raw.csv 2>/dev/null
S=$(stat -c %s /tmp/${R}_gemm_tc.ncu-rep 2>/dev/null || echo 999999999)
if [ "$S" -lt 20000000 ]; then cp /tmp/${R}_gemm_tc.ncu-rep gpurun_out/; fi
tail -c 300 gpurun_out/${R}_bench_final.log
