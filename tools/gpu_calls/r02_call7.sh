#!/bin/bash
mkdir -p gpurun_out/c7
O=gpurun_out/c7
run() { name=$1; shift; timeout 800 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -14 $O/$name.log; }
run gd7 python tools/graph_debug7.py
run gd7_512 python tools/graph_debug7.py 512 8 4 32
