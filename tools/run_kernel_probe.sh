#!/bin/bash
# Each group in its own process with a timeout (a trapped kernel cannot poison the next group).
mkdir -p gpurun_out
: > gpurun_out/kernel_probe.log
for g in ${@:-rowwise elem attn attn_perf}; do
  echo "=== group $g ===" >> gpurun_out/kernel_probe.log
  timeout 240 python tools/kernel_probe.py $g >> gpurun_out/kernel_probe.log 2>&1
  echo "exit=$?" >> gpurun_out/kernel_probe.log
done
tail -200 gpurun_out/kernel_probe.log
