#!/bin/bash
mkdir -p gpurun_out/c19
for r in 16 32 64; do
  MMDIT_ROW_RPB=$r timeout 300 python tools/row_probe.py perf > gpurun_out/c19/rpb$r.log 2>&1
  echo "== rpb $r"; grep -E "perf|bwd" gpurun_out/c19/rpb$r.log | cut -c1-120
done
