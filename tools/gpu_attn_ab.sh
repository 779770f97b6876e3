#!/bin/bash
# A/B of forward attention variant libraries (tools/bin/libvar_*.so) with tools/bench_attn.py; TESTS=1 also runs the GPU suite.
mkdir -p gpurun_out
{
[ -n "$TESTS" ] && timeout 900 python -m pytest -q -m gpu tests -x 2>&1 | tail -5
echo "== default lib"; timeout 120 python tools/bench_attn.py 2>&1 | tail -2
for v in tools/bin/libvar_*.so; do [ -f $v ] || continue; echo "== $v"; LDMAE_B200_LIB=$PWD/$v timeout 120 python tools/bench_attn.py 2>&1 | tail -2; done
} > gpurun_out/attn_ab.log 2>&1
cat gpurun_out/attn_ab.log
