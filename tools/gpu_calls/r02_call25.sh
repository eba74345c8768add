#!/bin/bash
mkdir -p gpurun_out/c25
O=gpurun_out/c25
run() { name=$1; shift; timeout 900 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -${TAILN:-3} $O/$name.log | cut -c1-200; }
run pytest python -m pytest tests -x -q -m gpu
run smoke python __graft_entry__.py smoke
run bench_cfg2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
run bench_cfg2b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
run bench_cfg3 python bench.py --config cfg3 --batch 64 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
run bench_cfg4 python bench.py --config cfg4 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
