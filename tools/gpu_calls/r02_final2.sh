#!/bin/bash
# Final check of the round after the row-kernel pass: what the driver runs (GPU tests, smoke, both bench arms), the
# cfg3 line, the cfg5 sampler, and the ncu launch list of bench.py itself.
mkdir -p gpurun_out/final2
O=gpurun_out/final2
run() { name=$1; shift; timeout 900 "$@" > $O/$name.log 2>&1; echo "$name exit=$?" | tee -a $O/$name.log; tail -${TAILN:-2} $O/$name.log | cut -c1-260; }
run pytest python -m pytest tests -x -q -m gpu
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench python bench.py --gpus 1 --steps 20 --warmup 5
run bench_ref python bench.py --impl reference --gpus 1 --steps 20 --warmup 5
run bench_cfg3 python bench.py --config cfg3 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline
run sample python tools/sample_bench.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline > $O/ncu_bench.log 2>&1; echo "ncu exit=$?"
python tools/summarize_launches.py $O/launches_bench.csv > $O/launches_bench_summary.txt 2>&1; head -12 $O/launches_bench_summary.txt
