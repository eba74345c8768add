#!/bin/bash
mkdir -p gpurun_out/c20
O=gpurun_out/c20
run() { name=$1; shift; timeout 900 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -${TAILN:-3} $O/$name.log | cut -c1-200; }
TAILN=2 run attn_probe python tools/kernel_probe.py attn
TAILN=2 run rowwise python tools/kernel_probe.py rowwise
run pytest python -m pytest tests -x -q -m gpu
MMDIT_ROW_FOLD=0 run bench_fold0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
run bench_fold1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
MMDIT_ROW_FOLD=0 run bench_fold0b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
run bench_fold1b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
