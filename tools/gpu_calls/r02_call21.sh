#!/bin/bash
mkdir -p gpurun_out/c21
for c in 3 6 12; do
  MMDIT_QKN_CAP=$c timeout 300 python tools/row_probe.py perf > gpurun_out/c21/cap$c.log 2>&1
  echo "== cap $c"; grep -E "perf|qknorm|swiglu" gpurun_out/c21/cap$c.log | cut -c1-120
done
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c21/pytest.log 2>&1; tail -2 gpurun_out/c21/pytest.log
