#!/bin/bash
# Round-2 GPU call 1: baseline of HEAD, validation of the never-run variants, library baselines, cfg3/cfg4.
mkdir -p gpurun_out/c1
O=gpurun_out/c1
run() { name=$1; shift; timeout 600 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -4 $O/$name.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.log
run pytest python -m pytest tests -x -q -m gpu
run qknorm python tools/gemm_probe.py qknorm
MMDIT_ATTN_BWD_PT_TMEM=1 run attn_pt_tmem python tools/kernel_probe.py attn
run attn_lib python tools/attn_lib_compare.py $O/attn_lib.json
run bench_cfg2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run bench_cfg3_b32 python bench.py --config cfg3 --steps 5 --warmup 3 --no-cpu-baseline
run bench_cfg3_b64 python bench.py --config cfg3 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline
run bench_cfg4_b16 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline
MMDIT_FUSED_QKNORM=1 run bench_cfg2_fusedqk python bench.py --steps 10 --warmup 3 --no-cpu-baseline
MMDIT_ATTN_BWD_PT_TMEM=1 run bench_cfg2_pt_tmem python bench.py --steps 10 --warmup 3 --no-cpu-baseline
MMDIT_DUAL_STREAM_INFER=1 run sample_ds python tools/sample_bench.py
run sample_base python tools/sample_bench.py
