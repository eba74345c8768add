#!/bin/bash
# Cluster (DSMEM) fold of the backward row kernels column sums: checks, GB/s table, GPU suite, bench.
mkdir -p gpurun_out/c34
O=gpurun_out/c34
timeout 300 python tools/row_probe.py check > $O/row_check.log 2>&1; echo "row check exit=$?"; grep -c FAIL $O/row_check.log; tail -1 $O/row_check.log
timeout 200 python tools/kernel_probe.py rowwise > $O/kp_rowwise.log 2>&1; echo "exit=$?"; tail -1 $O/kp_rowwise.log
timeout 200 python tools/kernel_probe.py elem > $O/kp_elem.log 2>&1; echo "exit=$?"; tail -1 $O/kp_elem.log
timeout 300 python tools/row_probe.py perf > $O/row_perf.log 2>&1; echo "row perf exit=$?"; grep -A24 "perf cfg2 image" $O/row_perf.log
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest exit=$?"; tail -2 $O/pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline > $O/bench_g2.log 2>&1; echo "exit=$?"; tail -1 $O/bench_g2.log | cut -c1-220
