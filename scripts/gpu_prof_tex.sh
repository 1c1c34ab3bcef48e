#!/bin/bash
# ncu full capture of the texture-assisted force kernel inside the bench workload (1 GPU)
mkdir -p gpurun_out
ARGS="--no-extra --no-cpu --no-e2e --no-ab --steps 1 --warmup 3 --melt 100"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force_full_tex -s 152 -c 2 \
    -f -o gpurun_out/prof_force_tex python bench.py $ARGS > gpurun_out/prof_bench.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/prof_force_tex.ncu-rep
