#!/bin/bash
mkdir -p gpurun_out/c27
O=gpurun_out/c27
run() { name=$1; shift; timeout 900 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -${TAILN:-3} $O/$name.log | cut -c1-200; }
MMDIT_FUSED_SWIGLU_BWD=1 run pytest python -m pytest tests -x -q -m gpu
MMDIT_FUSED_SWIGLU_BWD=0 run bench_f0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
MMDIT_FUSED_SWIGLU_BWD=1 run bench_f1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
MMDIT_FUSED_SWIGLU_BWD=0 run bench_f0b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
MMDIT_FUSED_SWIGLU_BWD=1 run bench_f1b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
