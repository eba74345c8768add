#!/bin/bash
mkdir -p gpurun_out/c26
SWIGLU_BWD_DEBUG=1 timeout 300 python tools/gemm_probe.py swiglu_bwd > gpurun_out/c26/swiglu_bwd.log 2>&1; echo "exit=$?"; tail -24 gpurun_out/c26/swiglu_bwd.log | cut -c1-220
