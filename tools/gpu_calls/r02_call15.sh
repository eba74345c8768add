#!/bin/bash
mkdir -p gpurun_out/c15
O=gpurun_out/c15
run() { name=$1; shift; timeout 600 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log; tail -${TAILN:-4} $O/$name.log | cut -c1-300; }
TAILN=3 run attn_probe python tools/kernel_probe.py attn
TAILN=8 run real_shapes python -m pytest tests/test_gpu_real_shapes.py -x -q -m gpu -k "attention"
TAILN=6 run attn_lib python tools/attn_lib_compare.py $O/attn_lib.json
python - <<'PY'
import json
j=json.load(open("gpurun_out/c15/attn_lib.json"))
j = j if isinstance(j,list) else [j]
for x in j: print(x['shape'], {k:round(v,1) for k,v in x.items() if k.endswith('_us')})
PY
