#!/bin/bash
mkdir -p gpurun_out/c14
O=gpurun_out/c14
run() { name=$1; shift; timeout 900 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -${TAILN:-3} $O/$name.log | cut -c1-300; }
TAILN=12 run elem python tools/kernel_probe.py elem
run pytest python -m pytest tests -x -q -m gpu
run bench_cfg2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
run bench_cfg2b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
