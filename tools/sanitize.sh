#!/bin/bash
# compute-sanitizer over the kernels where a race would be silent (the reference's own equivalent is the
# poison/shadow check of Sh3BinaryEvaluator.cpp:578-621).  Run through gpurun on ONE B200; summaries -> gpurun_out/<R>_san_*.log
# (copied into profiles/).  Each target also checks its result, so a sanitizer run is a parity run.
set -u
R=${1:-r2}
CS=/usr/local/cuda/bin/compute-sanitizer
mkdir -p gpurun_out
run() {  # tool target [extra env]
  local tool=$1 tgt=$2
  local out=gpurun_out/${R}_san_${tool}_${tgt}.log
  timeout 300 $CS --tool $tool --print-limit 20 --error-exitcode 77 python tools/sanitize_targets.py $tgt > $out 2>&1
  echo "rc=$?" >> $out
  echo "== $tool $tgt: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|rc=' $out | tr '\n' ' ')"
}
for tgt in gemm sgd early binary smoke; do run memcheck $tgt; done
for tgt in gemm sgd binary; do run racecheck $tgt; done
for tgt in gemm sgd early; do run synccheck $tgt; done
run initcheck gemm
