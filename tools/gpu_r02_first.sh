#!/bin/bash
# Round-2 first GPU visit: parity tests, smoke, a short run of the restructured bench (all extras), the GPU comparator,
# and ncu --set full captures of the residual-epilogue GEMMs (proj / w3).  Every step logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
run() {
  name=$1; shift
  t0=$(date +%s)
  timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err
  rc=$?
  echo "== $name rc=$rc ($(( $(date +%s) - t0 )) s) :: $(tail -n 1 gpurun_out/$name.log | cut -c1-400)"
}
run tests python -m pytest -q -m gpu tests -x -s
grep -hE "rel err|within|growth|grid index|FAILED|passed|failed|Error" gpurun_out/tests.log | head -40
run smoke python __graft_entry__.py smoke
TMO=500 run bench_short python bench.py --num-steps 30 --steps 2 --warmup 2
tail -n 5 gpurun_out/bench_short.err
TMO=200 run bench_ref python bench.py --impl reference --steps 1 --warmup 1
TMO=700 run bench_cmp python bench.py --impl torch-gpu --cmp-evals 6 --train-steps 5
tail -n 12 gpurun_out/bench_cmp.err | cut -c1-300
if [ "${NCU:-1}" = "1" ]; then
  CMD="python bench.py --batch 64 --num-steps 3 --steps 1 --warmup 1 --no-cpu-baseline --no-train --no-xl-extra --no-cond-only-extra --no-decode-extra"
  timeout 300 $CMD > gpurun_out/cap_plain.log 2>&1 || { echo "plain capture command failed"; tail -5 gpurun_out/cap_plain.log; exit 0; }
  for k in EpiResidualT; do
    timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -s 24 -c 2 -o gpurun_out/r02a_$k -f $CMD > gpurun_out/cap_$k.log 2>&1
    echo "== ncu $k rc=$?"
  done
  ls -la gpurun_out/r02a_*.ncu-rep
fi
