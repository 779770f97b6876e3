#!/bin/bash
mkdir -p gpurun_out
for ps in 1 2; do
  LDMAE_ATTN_PERSIST=$ps timeout 300 python bench.py --num-steps 4 --steps 1 --warmup 3 --no-xl-extra --no-cond-only-extra --no-cpu-baseline --no-train > gpurun_out/v.log 2> gpurun_out/v.err
  python - $ps <<'P'
import json, sys
d=json.loads(open('gpurun_out/v.log').read().strip().splitlines()[-1])
print('persist', sys.argv[1], d['vmae_decode'])
P
done
LDMAE_ATTN_PERSIST=2 timeout 300 python -m pytest -q -m gpu tests -x -k "vmae or decode or config1" 2>&1 | tail -2
