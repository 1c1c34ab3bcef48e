#!/bin/bash
set -x
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base function"
timeout 300 $NCU -k regex:'^k_s4$' -s 2 -c 1 -f -o gpurun_out/r2_s4 experiments/force_r2 100 1 "g s4 idx TEX, gathers LDG u3x4 (rows 8-class)" > gpurun_out/r2_ncu_s4.log 2>&1
timeout 300 $NCU -k regex:'^k_d4t$' -s 2 -c 1 -f -o gpurun_out/r2_d4t experiments/force_r2 100 1 "g d4t idx LDG.128, xy LDG.128 + z TEX u2x4 (rows 8-class)" > gpurun_out/r2_ncu_d4t.log 2>&1
tail -3 gpurun_out/r2_ncu_s4.log gpurun_out/r2_ncu_d4t.log
