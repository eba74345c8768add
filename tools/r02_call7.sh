#!/bin/bash
mkdir -p gpurun_out/c7
O=gpurun_out/c7
run() { name=$1; shift; timeout 600 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -16 $O/$name.log; }
run gd4 python tools/graph_debug4.py
run gd4_pre python tools/graph_debug4.py preloop
