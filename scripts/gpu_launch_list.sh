#!/bin/bash
# ncu launch list (device time of every kernel) of a short bench run.
mkdir -p gpurun_out
ARGS="--no-extra --no-cpu --no-e2e --steps 1 --warmup 3 --melt 0 ${BENCH_ARGS}"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
