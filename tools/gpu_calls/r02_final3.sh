#!/bin/bash
# Round-2 closing profile evidence with the second-generation row kernels: per-kernel time / DRAM bytes / tensor-pipe
# activity of ONE eager cfg2 step (profiler start/stop window), and one --set full capture of the LN-modulate backward.
mkdir -p gpurun_out/final3
O=gpurun_out/final3
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file $O/metrics_step.csv python tools/profile_step.py > $O/metrics.log 2>&1; echo "metrics exit=$?"
python tools/tensor_metrics.py $O/metrics_step.csv > $O/kernel_metrics.json 2>&1; head -c 1500 $O/kernel_metrics.json
python tools/summarize_launches.py $O/metrics_step.csv > $O/step_launches_summary.txt 2>&1; head -5 $O/step_launches_summary.txt
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ln_mod_bwd2 -s 2 -c 1 -f -o $O/ln_mod_bwd2_r02 python tools/profile_step.py > $O/ncu_full.log 2>&1; echo "full exit=$?"; ls -la $O
