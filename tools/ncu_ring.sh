#!/bin/bash
# one full-section ncu capture of k_gemv_ring at the logistic-inference shape (2^22 x 512)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemv_ring --launch-skip 1 --launch-count 1 -o /tmp/r2_gemv_ring \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-linreg --no-basic --no-c1 --no-strong > gpurun_out/r2_ncu_ring.log 2>&1
ncu -i /tmp/r2_gemv_ring.ncu-rep --page raw --csv > gpurun_out/r2_k_gemv_ring_raw.csv 2>/dev/null
ncu -i /tmp/r2_gemv_ring.ncu-rep --page details > gpurun_out/r2_k_gemv_ring_details.txt 2>/dev/null
grep -E "Duration|DRAM Throughput|Memory Throughput|Registers Per|Achieved Occ|Grid Size" gpurun_out/r2_k_gemv_ring_details.txt | head
