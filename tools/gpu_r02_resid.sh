#!/bin/bash
# GEMM-epilogue iteration: kernel tests, standalone timings (residual epilogue at the proj / w3 shapes, fused-vs-cuBLAS table,
# forward launch groups), then the whole parity suite and a short bench.
mkdir -p gpurun_out
timeout 300 python -m pytest -q -m gpu tests/test_gpu_kernels.py -x 2>&1 | tail -3
for w in proj w3; do timeout 60 python tools/prof_resid.py 524288 $w; done
LDMAE_RESID_DEEP=2 timeout 60 python tools/prof_resid.py 524288 w3 | sed 's/^/deep=2 /'
timeout 120 python tools/bench_stages.py 512 2>&1 | tail -12
if [ "${TRACE:-0}" = "1" ]; then
  for w in proj w3; do LDMAE_B200_LIB=tools/bin/libldmae_trace.so timeout 100 python tools/trace_resid.py $w > gpurun_out/trace_resid_$w.txt 2>&1; head -14 gpurun_out/trace_resid_$w.txt; done
fi
if [ "${FULL:-1}" = "1" ]; then
  timeout 600 python -m pytest -q -m gpu tests -x 2>&1 | tail -4
  timeout 400 python bench.py --num-steps 30 --steps 2 --warmup 2 --no-xl-extra --no-cond-only-extra --no-cpu-baseline > gpurun_out/bench_short.log 2> gpurun_out/bench_short.err
  python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_short.log').read().strip().splitlines()[-1])
print('value', d['value'], 'class_ms', d['class_ms_per_step'], 'tflops', d['class_tflops'], 'clk', d['clocks'], 'train', d.get('train',{}).get('value'), d.get('train',{}).get('class_ms_per_step'))
P
fi
