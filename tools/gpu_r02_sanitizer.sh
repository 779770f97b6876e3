#!/bin/bash
# compute-sanitizer over the kernel-level tests (small shapes incl. ragged T = 36 / 196 / 320 and single-key-block cases).
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  t0=$(date +%s)
  timeout ${TMO:-420} $S --tool $tool --print-limit 40 --log-file gpurun_out/sanitizer_$tool.log \
    python -m pytest -q -m gpu tests/test_gpu_kernels.py tests/test_gpu_backward_kernels.py -x -q > gpurun_out/sanitizer_${tool}_pytest.log 2>&1
  echo "== $tool rc=$? ($(( $(date +%s) - t0 )) s) :: $(tail -n 1 gpurun_out/sanitizer_${tool}_pytest.log | cut -c1-200)"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error" gpurun_out/sanitizer_$tool.log | sort | uniq -c | sort -rn | head -12
done
