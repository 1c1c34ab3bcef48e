#!/bin/bash
# ncu captures of the LJ force kernel inside the bench workload (1 GPU).
mkdir -p gpurun_out
ARGS="--no-extra --no-cpu --no-e2e --steps 1 --warmup 3 --melt 100 ${BENCH_ARGS}"
# launch list: every kernel with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 300 --csv \
    --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"
# full capture of the force kernel (skip the melt phase launches of that kernel)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-k_force} -s 150 -c 2 \
    -f -o gpurun_out/prof_force python bench.py $ARGS > gpurun_out/prof_bench.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out
