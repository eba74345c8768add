#!/bin/bash
mkdir -p gpurun_out/c30
O=gpurun_out/c30
MMDIT_FUSED_GATE=1 timeout 600 python tools/sample_bench.py > $O/sample_fusedgate.log 2>&1; echo "exit=$?"; tail -1 $O/sample_fusedgate.log
timeout 600 python tools/sample_bench.py > $O/sample_base.log 2>&1; echo "exit=$?"; tail -1 $O/sample_base.log
MMDIT_DUAL_STREAM_INFER=1 timeout 600 python tools/sample_bench.py > $O/sample_ds.log 2>&1; echo "exit=$?"; tail -1 $O/sample_ds.log
