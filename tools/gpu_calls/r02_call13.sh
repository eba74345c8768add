#!/bin/bash
mkdir -p gpurun_out/c13
O=gpurun_out/c13
run() { name=$1; shift; timeout 900 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -${TAILN:-3} $O/$name.log | cut -c1-300; }
MMDIT_PDL=0 run bench_pdl0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
MMDIT_PDL=1 run bench_late1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
MMDIT_PDL=0 run bench_pdl0b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
MMDIT_PDL=1 run bench_late1b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
