#!/bin/bash
# One GPU visit: parity tests, smoke, a short and the default bench, and the ncu launch list of a short bench.
# Every step runs in its own process under a timeout and logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
run() {
  name=$1; shift
  timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1
  rc=$?
  echo "== $name rc=$rc :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"
}
run tests python -m pytest -q -m gpu tests -x -s
run smoke python __graft_entry__.py smoke
run bench_small python bench.py --batch 32 --num-steps 6 --steps 2 --warmup 1 --no-cpu-baseline
if [ "${FULL:-1}" = "1" ]; then
  TMO=1500 run bench_full python bench.py
fi
if [ "${NCU:-1}" = "1" ]; then
  CMD="python bench.py --batch 64 --num-steps 3 --steps 1 --warmup 1 --no-cpu-baseline"
  timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_run.log 2>&1
  echo "== ncu launches rc=$? :: $(wc -l < gpurun_out/launches.csv 2>/dev/null) lines"
fi
grep -hE "rel err|Error|error|assert|FAILED|passed|failed|timeout|mbarrier|vmae " gpurun_out/tests.log | head -30
