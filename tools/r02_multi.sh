#!/bin/bash
# usage: bash tools/r02_multi.sh N config [batch] [tag]   -- one bench run on N GPUs of this box
N=$1; CFG=$2; BATCH=${3:-0}; TAG=${4:-}
mkdir -p gpurun_out/multi
OUT=gpurun_out/multi/bench_${CFG}_n${N}${TAG}.log
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 8 --warmup 3 --config $CFG --batch $BATCH --no-cpu-baseline --no-gpu-library-baseline > $OUT 2>&1
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 8 --warmup 3 --config $CFG --batch $BATCH --no-cpu-baseline --no-gpu-library-baseline > $OUT 2>&1
fi
echo "exit=$?" >> $OUT
grep -E '^\{' $OUT | tail -1 | cut -c1-400
tail -3 $OUT | cut -c1-300
