#!/bin/bash
# ncu launch list (gpu__time_duration) of a short sampling run at the benchmark batch (Bf = 512) and of a short training run,
# each after the plain command exited 0; optional --set full captures of named kernels.
mkdir -p gpurun_out
CMD="python bench.py --batch 256 --num-steps 3 --steps 1 --warmup 1 --no-cpu-baseline --no-train --no-xl-extra --no-cond-only-extra --no-decode-extra"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain failed"; tail -3 gpurun_out/ncu_plain.log; exit 0; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_run.log 2>&1
echo "== ncu sampling launches rc=$? :: $(wc -l < gpurun_out/launches.csv) lines"
python tools/summarize_launches.py gpurun_out/launches.csv | head -40
if [ "${TRAIN:-1}" = "1" ]; then
  CMD="python bench.py --train-only --train-batch 128 --train-steps 1"
  timeout 300 $CMD > gpurun_out/ncu_train_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/train_launches.csv $CMD > gpurun_out/ncu_train_run.log 2>&1
  echo "== ncu training launches rc=$? :: $(wc -l < gpurun_out/train_launches.csv) lines"
  python tools/summarize_launches.py gpurun_out/train_launches.csv | head -50
fi
for k in ${KERNELS}; do
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -s ${SKIP:-30} -c 1 -o gpurun_out/r02_$k -f $CMD > gpurun_out/cap_$k.log 2>&1
  echo "== ncu full $k rc=$?"
done
