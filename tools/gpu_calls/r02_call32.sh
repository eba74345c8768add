#!/bin/bash
# Second-generation row kernels: every kernel against its first-generation counterpart and against fp32
# torch, then the GB/s table of both generations, then one bench line per generation.
mkdir -p gpurun_out/c32
O=gpurun_out/c32
timeout 300 python tools/row_probe.py check > $O/row_check.log 2>&1; echo "row check exit=$?"; grep -c FAIL $O/row_check.log; tail -2 $O/row_check.log
MMDIT_ROW_KERNELS=2 timeout 200 python tools/kernel_probe.py rowwise > $O/kp_rowwise_g2.log 2>&1; echo "exit=$?"; tail -1 $O/kp_rowwise_g2.log
MMDIT_ROW_KERNELS=2 timeout 200 python tools/kernel_probe.py elem > $O/kp_elem_g2.log 2>&1; echo "exit=$?"; tail -1 $O/kp_elem_g2.log
timeout 300 python tools/row_probe.py perf > $O/row_perf.log 2>&1; echo "row perf exit=$?"; grep -A12 "perf cfg2 image" $O/row_perf.log
MMDIT_ROW_KERNELS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline > $O/bench_g1.log 2>&1; echo "exit=$?"; tail -1 $O/bench_g1.log | cut -c1-220
MMDIT_ROW_KERNELS=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline > $O/bench_g2.log 2>&1; echo "exit=$?"; tail -1 $O/bench_g2.log | cut -c1-220
