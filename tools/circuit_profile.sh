#!/bin/bash
# Config 5 (compare-exchange stage): tests of the binary engine, the stage profile, and one full-section ncu capture each of the
# two keystream-carrying circuit kernels (run through gpurun; outputs under gpurun_out/).
set -u
R=${1:-r2}
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_sh3.py tests/test_gpu_kernels.py -m gpu -x -q -k "bin or circuit or basic or bitwise or and or merge or comparison" 2>&1 | tail -3 | tee gpurun_out/${R}_circuit_tests.log
timeout 600 python tools/basic_profile.py 8388605 > gpurun_out/${R}_basic_profile3.log 2>&1; cat gpurun_out/${R}_basic_profile3.log
for K in k_bitwise_rowmajor k_bin_and_layer; do
  ABY3_BASIC_ONLY=split timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 9 --launch-count 1 -o /tmp/${R}_$K \
      python tools/basic_profile.py 8388605 > gpurun_out/${R}_ncu_$K.log 2>&1
  ncu -i /tmp/${R}_$K.ncu-rep --page raw --csv > gpurun_out/${R}_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/${R}_$K.ncu-rep --page source --csv > gpurun_out/${R}_${K}_source.csv 2>/dev/null
  ncu -i /tmp/${R}_$K.ncu-rep --page details > gpurun_out/${R}_${K}_details.txt 2>/dev/null
done
ls -la gpurun_out | tail -8
