#!/bin/bash
export MMDIT_FUSED_GATE_LN=0
bash tools/r02_multi.sh 1 cfg3 64 _base
bash tools/r02_multi.sh 1 cfg4 16 _base
bash tools/r02_multi.sh 1 cfg2 0 _base
export MMDIT_FUSED_GATE_LN=1
bash tools/r02_multi.sh 1 cfg2 0 _fusedln
mkdir -p gpurun_out/c9
timeout 600 python tools/kernel_probe.py rowwise > gpurun_out/c9/rowwise.log 2>&1; tail -3 gpurun_out/c9/rowwise.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/c9/pytest_parity.log 2>&1; tail -3 gpurun_out/c9/pytest_parity.log
timeout 300 python tools/attn_bwd_timeline.py > gpurun_out/c9/bwd_timeline.log 2>&1; tail -2 gpurun_out/c9/bwd_timeline.log | cut -c1-300
