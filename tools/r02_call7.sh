#!/bin/bash
mkdir -p gpurun_out/c7
O=gpurun_out/c7
run() { name=$1; shift; timeout 600 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -14 $O/$name.log; }
run gd2_512 python tools/graph_debug2.py 512 8 4 32 6
run gd2_256 python tools/graph_debug2.py 256 4 2 32 6
