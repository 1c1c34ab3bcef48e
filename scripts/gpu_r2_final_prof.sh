#!/bin/bash
# round-2 profiles of the final code: launch list of two bench steps + ncu --set full captures of the
# FP64 and FP32 force sweeps and of the neighbour build
mkdir -p gpurun_out
B="python bench.py --no-extra --no-cpu --no-e2e --no-ab --no-checks --melt 20 --warmup 1 --steps 2"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_launches_bench.json 2> gpurun_out/r2_launches_bench.err
echo "launch list rc=$?"; wc -l gpurun_out/r2_launches.csv
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_force_full -s 40 -c 2 -f -o gpurun_out/r2_final_force_f64 $B > /dev/null 2>&1; echo "f64 rc=$?"
CBMD_PRECISION=32 timeout 600 $NCU -k regex:k_force_full_f32 -s 40 -c 2 -f -o gpurun_out/r2_final_force_f32 $B > /dev/null 2>&1; echo "f32 rc=$?"
timeout 600 $NCU -k regex:k_neigh_build -s 2 -c 1 -f -o gpurun_out/r2_final_neigh $B > /dev/null 2>&1; echo "neigh rc=$?"
ls -la gpurun_out/r2_final_*.ncu-rep
