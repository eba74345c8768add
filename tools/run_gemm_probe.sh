#!/bin/bash
# Runs each GEMM probe group in its own process (a trap in one cannot poison the next).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gemm_probe.log 2>&1
for g in basic major epi perf; do
  echo "=== group $g ===" >> gpurun_out/gemm_probe.log
  timeout 180 python tools/gemm_probe.py $g >> gpurun_out/gemm_probe.log 2>&1
  echo "exit=$?" >> gpurun_out/gemm_probe.log
done
tail -150 gpurun_out/gemm_probe.log
