#!/bin/bash
mkdir -p gpurun_out/c3
O=gpurun_out/c3
run() { name=$1; shift; timeout 900 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -4 $O/$name.log; }
run attn_probe python tools/kernel_probe.py attn
run attn_lib python tools/attn_lib_compare.py $O/attn_lib.json
run pytest python -m pytest tests -q -m gpu -s --timeout 850
run bench_cfg2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
