#!/bin/bash
# Full GPU check of one round: pytest -m gpu, smoke, bench (graph + eager). Logs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_graph.log 2>&1; echo "exit=$?" >> gpurun_out/bench_graph.log; tail -6 gpurun_out/bench_graph.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_eager.log 2>&1; echo "exit=$?" >> gpurun_out/bench_eager.log; tail -6 gpurun_out/bench_eager.log
