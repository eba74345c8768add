#!/bin/bash
# Round-2 profile evidence: ncu launch list of bench.py itself, per-kernel metrics of one step,
# --set full captures of the attention kernels and of in-step GEMM launches.
mkdir -p gpurun_out/c17
O=gpurun_out/c17
run() { name=$1; shift; timeout 1500 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -3 $O/$name.log | cut -c1-200; }
run ncu_bwd2 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2_kernel -s 1 -c 1 -f -o $O/attn_bwd2_r02 python tools/attn_one.py cfg2 2
run ncu_fwd2 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2_kernel -s 1 -c 1 -f -o $O/attn_fwd2_r02 python tools/attn_one.py cfg2 2
run ncu_gemm ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05 -s 60 -c 4 -f -o $O/gemm_r02 python tools/profile_step.py
run metrics ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file $O/metrics_r02.csv python tools/profile_step.py
python tools/tensor_metrics.py $O/metrics_r02.csv > $O/kernel_metrics_r02.json 2> $O/tensor_metrics.err; head -c 600 $O/kernel_metrics_r02.json
run launches ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-library-baseline
python tools/summarize_launches.py $O/launches_bench_r02.csv > $O/launches_bench_r02_summary.txt 2>&1; head -20 $O/launches_bench_r02_summary.txt
ls -la $O
