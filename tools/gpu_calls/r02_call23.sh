#!/bin/bash
mkdir -p gpurun_out/c23
timeout 300 python tools/kernel_probe.py rowwise > gpurun_out/c23/rowwise.log 2>&1; tail -1 gpurun_out/c23/rowwise.log
timeout 300 python tools/row_probe.py perf > gpurun_out/c23/auto.log 2>&1
echo "== auto"; grep -E "perf|bwd" gpurun_out/c23/auto.log | grep -v qknorm | cut -c1-120
MMDIT_ROW_RPB=16 timeout 300 python tools/row_probe.py perf > gpurun_out/c23/r16.log 2>&1
echo "== 16"; grep -E "perf|bwd" gpurun_out/c23/r16.log | grep -v qknorm | cut -c1-120
