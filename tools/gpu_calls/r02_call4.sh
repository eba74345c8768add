#!/bin/bash
mkdir -p gpurun_out/c4
O=gpurun_out/c4
run() { name=$1; shift; timeout 900 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -4 $O/$name.log; }
run timeline python tools/attn2_timeline.py cfg2 0 101
run pytest python -m pytest tests -q -m gpu -s --timeout 850
run sample python tools/sample_bench.py
