#!/bin/bash
mkdir -p gpurun_out/multi
timeout 600 python -m pytest tests/test_gpu_ddp.py -x -q -m gpu > gpurun_out/multi/pytest_ddp_2gpu_b.log 2>&1; echo "exit=$?" >> gpurun_out/multi/pytest_ddp_2gpu_b.log; tail -4 gpurun_out/multi/pytest_ddp_2gpu_b.log
bash tools/r02_multi.sh 2 cfg2 0 _b
bash tools/r02_multi.sh 1 cfg2 0 _b
