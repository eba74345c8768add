#!/bin/bash
mkdir -p gpurun_out/c28
timeout 300 python tools/gemm_probe.py swiglu > gpurun_out/c28/swiglu.log 2>&1; echo "exit=$?"; tail -9 gpurun_out/c28/swiglu.log | cut -c1-220
