#!/bin/bash
# Short GPU visit: kernel + parity tests, GEMM micro-benchmark, a short bench.
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1
  rc=$?
  echo "== $name rc=$rc :: $(tail -n 1 gpurun_out/$name.log | cut -c1-400)"
}
run tests python -m pytest -q -m gpu tests -x -s
run bench_gemm python tools/bench_gemm.py
run bench_small python bench.py --batch 64 --num-steps 6 --steps 2 --warmup 1 --no-cpu-baseline --train-batch 64 --train-steps 5
grep -hE "rel err|Error|error|assert|FAILED|passed|failed|timeout|mbarrier|vmae " gpurun_out/tests.log | head -30
cat gpurun_out/bench_gemm.log
python - <<'PY'
import json
for l in open("gpurun_out/bench_small.log"):
    if l.startswith("{"):
        d = json.loads(l); print("value", d["value"], "dit_frac", d["dit_frac_of_bf16_peak"], "class_ms", d["class_ms_per_step"], "roof", d["roofline"])
PY
