#!/bin/bash
# ncu --set full capture of one kernel (KREGEX) inside a short bench run; OUT = report name
mkdir -p gpurun_out
ARGS="--no-extra --no-cpu --no-e2e --steps 1 --warmup 3 --melt ${MELT:-40} ${BENCH_ARGS}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s ${SKIP:-2} -c ${COUNT:-1} \
    -f -o gpurun_out/${OUT:-prof} python bench.py $ARGS > gpurun_out/${OUT:-prof}_bench.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -5
