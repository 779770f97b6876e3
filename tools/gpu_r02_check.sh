#!/bin/bash
# Round-2 routine visit: all parity tests (no -x), smoke, a short run of bench.py with every extra object.
mkdir -p gpurun_out
run() {
  name=$1; shift
  t0=$(date +%s)
  timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err
  rc=$?
  echo "== $name rc=$rc ($(( $(date +%s) - t0 )) s) :: $(tail -n 1 gpurun_out/$name.log | cut -c1-600)"
}
run tests python -m pytest -q -m gpu tests -s
grep -hE "rel err|within|grid index|FAILED|passed|failed|Error" gpurun_out/tests.log | head -60
run smoke python __graft_entry__.py smoke
TMO=500 run bench_short python bench.py --num-steps 30 --steps 2 --warmup 2 ${BENCH_ARGS}
tail -n 5 gpurun_out/bench_short.err
