#!/bin/bash
mkdir -p gpurun_out/c24
timeout 300 python tools/kernel_probe.py rowwise > gpurun_out/c24/rowwise.log 2>&1; tail -1 gpurun_out/c24/rowwise.log
timeout 300 python tools/row_probe.py all > gpurun_out/c24/auto.log 2>&1
echo "== one wave"; grep -E "CHECK|perf|fwd" gpurun_out/c24/auto.log | grep -v "qknorm\|differing\|vs fp32" | cut -c1-120
MMDIT_ROW_ONE_WAVE=0 timeout 300 python tools/row_probe.py perf > gpurun_out/c24/all.log 2>&1
echo "== all blocks"; grep -E "perf|fwd" gpurun_out/c24/all.log | grep -v qknorm | cut -c1-120
