#!/bin/bash
# Same-box A/B of two builds of the library: tools/bin/libldmae_base.so (baseline) against the in-tree one, alternating,
# short B/1 sampling bench; prints the job rate and the class rates.
mkdir -p gpurun_out
for rep in 1 2 3; do
  for which in base new; do
    if [ $which = base ]; then export LDMAE_B200_LIB=$PWD/tools/bin/libldmae_base.so; else unset LDMAE_B200_LIB; fi
    if [ "${MODE:-sample}" = train ]; then
      timeout 300 python bench.py --train-only > gpurun_out/ab.log 2> gpurun_out/ab.err
      python - $which <<'P'
import json, sys
d=json.loads(open('gpurun_out/ab.log').read().strip().splitlines()[-1])
print(sys.argv[1], 'train', round(d['value'],1), d['class_ms_per_step'])
P
    else
    timeout 300 python bench.py --num-steps ${PTS:-30} --steps 2 --warmup 3 --no-xl-extra --no-cond-only-extra --no-cpu-baseline --no-train --no-decode-extra > gpurun_out/ab.log 2> gpurun_out/ab.err
    python - $which <<'P'
import json, sys
d=json.loads(open('gpurun_out/ab.log').read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(d['value'],2), {k: round(v) for k,v in d['class_tflops'].items()}, d['clocks']['sm_mhz'])
P
    fi
  done
done
