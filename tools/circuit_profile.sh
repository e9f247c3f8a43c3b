#!/bin/bash
# Config 5 (compare-exchange stage): tests of the binary engine, the stage profile, the launch list of one stage and (with a
# second argument) one full-section ncu capture each of the keystream-carrying circuit kernels.  Run through gpurun; outputs
# under gpurun_out/.
set -u
R=${1:-r2}
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_sh3.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${R}_circuit_tests.log
timeout 600 python tools/basic_profile.py 8388605 > gpurun_out/${R}_basic_profile3.log 2>&1; cat gpurun_out/${R}_basic_profile3.log
ABY3_BASIC_ONLY=split timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_split_launches.csv \
    python tools/basic_profile.py 8388605 > gpurun_out/${R}_split_ncu.log 2>&1
if [ -n "${2:-}" ]; then
for K in $2; do
  ABY3_BASIC_ONLY=split timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 4 --launch-count 1 -o /tmp/${R}_$K \
      python tools/basic_profile.py 8388605 > gpurun_out/${R}_ncu_$K.log 2>&1
  ncu -i /tmp/${R}_$K.ncu-rep --page raw --csv > gpurun_out/${R}_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/${R}_$K.ncu-rep --page source --csv > gpurun_out/${R}_${K}_source.csv 2>/dev/null
  ncu -i /tmp/${R}_$K.ncu-rep --page details > gpurun_out/${R}_${K}_details.txt 2>/dev/null
done
fi
