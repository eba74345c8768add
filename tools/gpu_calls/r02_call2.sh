#!/bin/bash
# Round-2 GPU call 2: new real-shape tests (no -x: see every failure), new bench probe + library baseline,
# attention library table, ncu source-level captures of the attention kernels.
mkdir -p gpurun_out/c2
O=gpurun_out/c2
run() { name=$1; shift; timeout 900 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -4 $O/$name.log; }
run pytest python -m pytest tests -q -m gpu -s --timeout 850
run attn_lib python tools/attn_lib_compare.py $O/attn_lib.json
run bench_cfg2 python bench.py --steps 10 --warmup 3
run ncu_attn_fwd ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 1 -c 1 -f -o $O/attn_fwd_r02a python tools/attn_one.py cfg2 2
run ncu_attn_bwd ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 1 -c 1 -f -o $O/attn_bwd_r02a python tools/attn_one.py cfg2 2
ls -la $O
