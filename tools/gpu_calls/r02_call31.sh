#!/bin/bash
mkdir -p gpurun_out/c31
MMDIT_ATTN_BWD_V2=0 timeout 300 python tools/kernel_probe.py attn > gpurun_out/c31/attn_v1.log 2>&1; echo "exit=$?"; tail -1 gpurun_out/c31/attn_v1.log
timeout 300 python tools/kernel_probe.py attn > gpurun_out/c31/attn_v2.log 2>&1; echo "exit=$?"; tail -1 gpurun_out/c31/attn_v2.log
