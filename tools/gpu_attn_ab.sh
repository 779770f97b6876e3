#!/bin/bash
# A/B of the forward attention variants (tools/bench_attn.py) + the GPU test suite.
mkdir -p gpurun_out
{
timeout 900 python -m pytest -q -m gpu tests -x 2>&1 | tail -15
echo "== default lib"; timeout 120 python tools/bench_attn.py 2>&1 | tail -2
for v in tools/bin/libvar_*.so; do [ -f $v ] || continue; echo "== $v"; LDMAE_B200_LIB=$PWD/$v timeout 120 python tools/bench_attn.py 2>&1 | tail -2; done
} > gpurun_out/attn_ab.log 2>&1
cat gpurun_out/attn_ab.log
