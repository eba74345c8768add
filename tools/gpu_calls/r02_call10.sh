#!/bin/bash
mkdir -p gpurun_out/c10
O=gpurun_out/c10
run() { name=$1; shift; timeout 900 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -5 $O/$name.log | cut -c1-400; }
run attn_probe python tools/kernel_probe.py attn
run attn_lib python tools/attn_lib_compare.py $O/attn_lib.json
run bwd_timeline python tools/attn_bwd_timeline.py
run bench_cfg2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
