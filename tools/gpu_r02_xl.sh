#!/bin/bash
# XL / wide-head iteration: the wide-head tests, then the XL/1 @512 sampling + training objects of bench.py (short B/1 run in
# front).  ENVS: space-separated VAR=value settings to compare (default: the shipped configuration only).
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu tests -x -k "wide or xl or head_dim_72 or wider" 2>&1 | tail -3
IFS=';' read -ra SETS <<< "${ENVS:-X=1}"
for kv in "${SETS[@]}"; do
  env $kv timeout 600 python bench.py --num-steps 3 --steps 1 --warmup 3 --no-decode-extra --no-cond-only-extra --no-cpu-baseline --no-train --xl-train > gpurun_out/bench_xl.log 2> gpurun_out/bench_xl.err
  python - "$kv" <<'P'
import json, sys
d=json.loads(open('gpurun_out/bench_xl.log').read().strip().splitlines()[-1])
x=d['xl_512']; t=x.get('train', {})
print(sys.argv[1], 'xl sampling TF/s', round(x['tflops_per_gpu'],1), 'ms', round(x['ms'],1), {k: round(v,1) for k,v in x['class_ms'].items()}, '| train', t.get('value'), t.get('class_ms_per_step'))
P
done
