#!/bin/bash
mkdir -p gpurun_out/c5
O=gpurun_out/c5
run() { name=$1; shift; timeout 600 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -5 $O/$name.log; }
run t1 python -m pytest tests/test_gpu_next_rows.py -q -m gpu -k captured
run t2 python -m pytest tests/test_gpu_real_shapes.py -q -m gpu -k 50_step -s
MMDIT_ATTN_FWD_V2=0 run t3 python -m pytest tests/test_gpu_next_rows.py tests/test_gpu_real_shapes.py -q -m gpu -k "captured or 50_step" -s
run t4 python -m pytest tests/test_gpu_next_rows.py -q -m gpu
run attn_lib python tools/attn_lib_compare.py $O/attn_lib.json
run attn_probe python tools/kernel_probe.py attn
