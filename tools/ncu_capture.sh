#!/bin/bash
# ncu --set full captures (one launch each) of the dominant kernels of the sampling step, taken from a short bench run
# after the plain command exited 0.  Reports land in gpurun_out/ (summaries are copied to profiles/ by tools/ncu_summary.py).
mkdir -p gpurun_out
set -x
CMD="python bench.py --batch 64 --num-steps 3 --steps 1 --warmup 1 --no-cpu-baseline --no-train"
timeout 300 $CMD > gpurun_out/cap_plain.log 2>&1 || exit 1
for k in ${KERNELS:-attn_fwd_persist EpiSwiGLU EpiQKV}; do
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -s 30 -c 1 -o gpurun_out/r01f_$k -f $CMD > gpurun_out/cap_$k.log 2>&1
done
ls -la gpurun_out/r01f_*.ncu-rep
