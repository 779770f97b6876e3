#!/bin/bash
# XL / wide-head iteration: the wide-head tests, then the XL/1 @512 sampling + training objects of bench.py (short B/1 run in front)
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu tests -x -k "wide or xl or head_dim_72 or wider" 2>&1 | tail -3
for tab in ${TABS:-1 0}; do
  LDMAE_ROPE_WIDE_TABLE=$tab timeout 600 python bench.py --num-steps 3 --steps 1 --warmup 3 --no-decode-extra --no-cond-only-extra --no-cpu-baseline --no-train --xl-train > gpurun_out/bench_xl_tab$tab.log 2> gpurun_out/bench_xl_tab$tab.err
  python - $tab <<'P'
import json, sys
d=json.loads(open(f'gpurun_out/bench_xl_tab{sys.argv[1]}.log').read().strip().splitlines()[-1])
x=d['xl_512']; t=x.get('train', {})
print('table', sys.argv[1], 'xl sampling TF/s', round(x['tflops_per_gpu'],1), 'ms', round(x['ms'],1), '| train', t.get('value'), t.get('class_ms_per_step'))
P
done
