#!/bin/bash
mkdir -p gpurun_out/c6
O=gpurun_out/c6
run() { name=$1; shift; timeout 600 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -12 $O/$name.log; }
run gd_rand32 python tools/graph_debug.py rand 32
run gd_synth32 python tools/graph_debug.py synth 32
run gd_synth16 python tools/graph_debug.py synth 16
run attn_lib python tools/attn_lib_compare.py $O/attn_lib.json
run gemm_shapes python tools/gemm_shapes.py cfg2
run bench_cfg2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
