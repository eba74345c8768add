#!/bin/bash
mkdir -p gpurun_out/c18
O=gpurun_out/c18
run() { name=$1; shift; timeout 600 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -${TAILN:-4} $O/$name.log | cut -c1-250; }
TAILN=16 run attn_probe python tools/kernel_probe.py attn
run real_shapes python -m pytest tests/test_gpu_real_shapes.py -x -q -m gpu -k "attention"

