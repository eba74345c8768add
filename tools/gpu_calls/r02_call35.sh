#!/bin/bash
# Same-box A/B of the step: row-kernel generation, cluster fold, zero pool (one bench line each), then the GPU suite.
mkdir -p gpurun_out/c35
O=gpurun_out/c35
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-library-baseline > $O/$name.log 2>&1; echo "$name exit=$? $(grep -o '"ms_per_step": [0-9.]*' $O/$name.log | head -2 | tr '\n' ' ')"; }
run a_gen1           MMDIT_ROW_KERNELS=1 MMDIT_ZERO_POOL=0
run b_gen2_nocl      MMDIT_ROW_KERNELS=2 MMDIT_ROW_CLUSTER=0 MMDIT_ZERO_POOL=0
run c_gen2_cl        MMDIT_ROW_KERNELS=2 MMDIT_ROW_CLUSTER=1 MMDIT_ZERO_POOL=0
run d_gen2_cl_pool   MMDIT_ROW_KERNELS=2 MMDIT_ROW_CLUSTER=1 MMDIT_ZERO_POOL=1
run e_gen2_nocl_pool MMDIT_ROW_KERNELS=2 MMDIT_ROW_CLUSTER=0 MMDIT_ZERO_POOL=1
run a2_gen1          MMDIT_ROW_KERNELS=1 MMDIT_ZERO_POOL=0
run d2_gen2_cl_pool  MMDIT_ROW_KERNELS=2 MMDIT_ROW_CLUSTER=1 MMDIT_ZERO_POOL=1
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest exit=$?"; tail -2 $O/pytest.log
