#!/bin/bash
# ncu --set full captures of the top kernels (one launch each), after the plain commands exited 0.  Reports land in gpurun_out/.
mkdir -p gpurun_out
set -x
timeout 200 python tools/bench_attn.py > gpurun_out/cap_attn_plain.log 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_persist -s 4 -c 1 -o gpurun_out/r01_attn_fwd_persist -f python tools/bench_attn.py > gpurun_out/cap_attn.log 2>&1
timeout 200 python tools/bench_bwd.py 32768 > gpurun_out/cap_bwd_plain.log 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 4 -c 2 -o gpurun_out/r01_attn_bwd -f python tools/bench_bwd.py 32768 > gpurun_out/cap_bwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_wgrad -s 12 -c 4 -o gpurun_out/r01_wgrad -f python tools/bench_bwd.py 32768 > gpurun_out/cap_wgrad.log 2>&1
ls -la gpurun_out/*.ncu-rep
