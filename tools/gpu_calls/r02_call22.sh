#!/bin/bash
mkdir -p gpurun_out/c22
for c in 4 8 16; do
  MMDIT_GRID_CAP=$c timeout 300 python tools/row_probe.py perf > gpurun_out/c22/g$c.log 2>&1
  echo "== grid cap $c"; grep -E "perf|qknorm_rope_fwd|gate_residual_fwd|qknorm_rope_bwd .fp32" gpurun_out/c22/g$c.log | cut -c1-120
done
