#!/bin/bash
# ncu --set full captures of the top kernels (one launch each), after the plain commands exited 0.  Reports land in gpurun_out/.
mkdir -p gpurun_out
set -x
timeout 200 python tools/bench_attn.py > gpurun_out/cap_attn_plain.log 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 6 -c 1 -o gpurun_out/r01_attn_fwd -f python tools/bench_attn.py > gpurun_out/cap_attn.log 2>&1
timeout 200 python tools/bench_bwd.py 32768 > gpurun_out/cap_bwd_plain.log 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd|gemm_wgrad" -s 8 -c 6 -o gpurun_out/r01_bwd_blocks -f python tools/bench_bwd.py 32768 > gpurun_out/cap_bwd.log 2>&1
timeout 300 python bench.py --train-only --train-batch 32 --train-steps 1 > gpurun_out/cap_train_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"resid_bwd|qkv_bwd|swiglu_bwd|EpiResidualT|EpiStoreI13" -s 300 -c 8 -o gpurun_out/r01_train_kernels -f python bench.py --train-only --train-batch 32 --train-steps 1 > gpurun_out/cap_train.log 2>&1
ls -la gpurun_out/*.ncu-rep
