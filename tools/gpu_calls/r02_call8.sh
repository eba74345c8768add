#!/bin/bash
mkdir -p gpurun_out/c8
O=gpurun_out/c8
run() { name=$1; shift; timeout 900 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -6 $O/$name.log; }
run pytest python -m pytest tests -q -m gpu --timeout 850
run gd2_256 python tools/graph_debug2.py 256 4 2 32 6
run sample python tools/sample_bench.py
