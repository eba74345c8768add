#!/bin/bash
mkdir -p gpurun_out/c29
O=gpurun_out/c29
timeout 600 python -m pytest tests -x -q -m gpu -k "sampler or euler or gemm" > $O/pytest_sampler.log 2>&1; echo "exit=$?"; tail -2 $O/pytest_sampler.log
timeout 600 python tools/sample_bench.py > $O/sample.log 2>&1; echo "exit=$?"; tail -3 $O/sample.log
