#!/bin/bash
# round-2 experiment batch 1: force-kernel A/B legs + ncu captures of four of them
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt
timeout 600 experiments/force_r2 100 5 > gpurun_out/r2_force_variants.txt 2>&1
tail -5 gpurun_out/r2_force_variants.txt
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base function"
timeout 300 $NCU -k regex:'^k_d$' -s 2 -c 1 -f -o gpurun_out/r2_d0 experiments/force_r2 100 1 "d0 product" > gpurun_out/r2_ncu_d0.log 2>&1
timeout 300 $NCU -k regex:'^k_dp$' -s 2 -c 1 -f -o gpurun_out/r2_dp_blob experiments/force_r2 100 1 "dp persistent blob    rcp3 icmp u6 noalloc" > gpurun_out/r2_ncu_dp.log 2>&1
timeout 300 $NCU -k regex:'^k_s$' -s 2 -c 1 -f -o gpurun_out/r2_s0 experiments/force_r2 100 1 "s0 fp32 float4 LDG.128 u8" > gpurun_out/r2_ncu_s0.log 2>&1
timeout 300 $NCU -k regex:'^k_sp$' -s 2 -c 1 -f -o gpurun_out/r2_sp_blob experiments/force_r2 100 1 "sp persistent blob    float4 LDG.128 u8 noalloc" > gpurun_out/r2_ncu_sp.log 2>&1
ls -la gpurun_out/*.ncu-rep
