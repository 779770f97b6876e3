#!/bin/bash
# Runs each GPU test group in its own process (a kernel trap poisons the CUDA context of that
# process only) under a timeout, logs to gpurun_out/diag_*.log and prints a one-line verdict each.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/diag_gpu.txt 2>&1
run() {
  name=$1; shift
  timeout 600 "$@" > gpurun_out/diag_$name.log 2>&1
  rc=$?
  echo "== $name rc=$rc :: $(tail -n 1 gpurun_out/diag_$name.log | cut -c1-200)"
}
run gemm_cg1_128 python -m pytest -q -m gpu tests/test_gpu_kernels.py -k "gemm and cfg0" -x
run gemm_cg1_256 python -m pytest -q -m gpu tests/test_gpu_kernels.py -k "gemm and cfg1" -x
run gemm_cg2_256 python -m pytest -q -m gpu tests/test_gpu_kernels.py -k "gemm and cfg2" -x
run gemm_bf16 python -m pytest -q -m gpu tests/test_gpu_kernels.py -k "bf16_out"
run attention python -m pytest -q -m gpu tests/test_gpu_kernels.py -k "attention"
run dit_tiny python -m pytest -q -m gpu tests/test_gpu_parity.py -k "dit_tiny"
run sampler python -m pytest -q -m gpu tests/test_gpu_parity.py -k "sampler"
run dit_b1 python -m pytest -q -m gpu tests/test_gpu_parity.py -k "dit_b1"
run vmae python -m pytest -q -m gpu tests/test_gpu_parity.py -k "vmae"
run job python -m pytest -q -m gpu tests/test_gpu_parity.py -k "sampling_job"
for f in gpurun_out/diag_*.log; do echo "---- $f"; grep -E "rel err|Error|error|assert|FAILED|passed|failed|timeout|mbarrier" $f | head -12; done
