#!/bin/bash
# ncu launch list (device time of every kernel) of the timed part of a short bench run:
# the melt/warm-up launches are skipped (-s), then 3 bench steps = 60 MD steps are listed.
mkdir -p gpurun_out
ARGS="--no-extra --no-cpu --no-e2e --no-ab --steps 3 --warmup 3 --melt 0 ${BENCH_ARGS}"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-330} -c ${COUNT:-420} --csv \
    --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
python scripts/launch_summary.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt
