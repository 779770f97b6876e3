#!/bin/bash
# One GPU visit: parity tests, smoke, the default bench (sampling + train object), the reference arm, and the ncu launch
# lists of a short sampling bench and a short training bench.  Every step runs in its own process under a timeout and logs
# to gpurun_out/.
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
if [ "${FULL:-1}" = "1" ]; then
  TMO=1500 run bench_full python bench.py
  TMO=600 run bench_ref python bench.py --impl reference --steps 1 --warmup 1
fi
if [ "${NCU:-1}" = "1" ]; then
  CMD="python bench.py --batch 64 --num-steps 3 --steps 1 --warmup 1 --no-cpu-baseline --no-train"
  timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_run.log 2>&1
  echo "== ncu sampling launches rc=$? :: $(wc -l < gpurun_out/launches.csv 2>/dev/null) lines"
  CMD="python bench.py --train-only --train-batch 32 --train-steps 1"
  timeout 600 $CMD > gpurun_out/ncu_train_plain.log 2>&1 && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/ncu_train_run.log 2>&1
  echo "== ncu training launches rc=$? :: $(wc -l < gpurun_out/train_launches.csv 2>/dev/null) lines"
fi
grep -hE "rel err|Error|error|assert|FAILED|passed|failed|timeout|mbarrier|vmae " gpurun_out/tests.log | head -30
