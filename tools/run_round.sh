#!/bin/bash
# Full GPU check of one round on ONE B200: pytest -m gpu, smoke, bench (graph + eager). Logs under gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/run_round.sh'
# Multi-GPU (N = 2, 4, 8; charged N x):
#   gpurun --gpus N -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
#       --master-port 29511 tools/ddp_check.py'      # exchange kernel vs NCCL, trainer modes
#   ... bench.py --gpus N [--exchange nccl]           # scaling
# Per-shape GEMM table vs cuBLAS: python tools/gemm_shapes.py cfg2|cfg3
# Launch list / per-kernel metrics of one step: see tools/profile_step.py, tools/tensor_metrics.py
# (do NOT run ncu over bench.py itself: ~8000 launches take > 5 minutes of box time)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_graph.log 2>&1; echo "exit=$?" >> gpurun_out/bench_graph.log; tail -6 gpurun_out/bench_graph.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_eager.log 2>&1; echo "exit=$?" >> gpurun_out/bench_eager.log; tail -6 gpurun_out/bench_eager.log
