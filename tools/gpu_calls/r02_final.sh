#!/bin/bash
# Final check of the round: what the driver runs (GPU tests, smoke, both bench arms).
mkdir -p gpurun_out/final
O=gpurun_out/final
run() { name=$1; shift; timeout 1200 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -${TAILN:-3} $O/$name.log | cut -c1-300; }
run pytest python -m pytest tests -x -q -m gpu
run smoke python __graft_entry__.py smoke
run bench python bench.py --gpus 1 --steps 20 --warmup 5
run bench_ref python bench.py --impl reference --gpus 1 --steps 20 --warmup 5
