#!/bin/bash
# neighbour-build correctness (gpu tests touching the list) + timing of the build kernels
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tstneighbor or neighbor or force_energy or halo or trajectory" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_neigh_build|k_pack_cand|k_cell_" -c 21 --csv --log-file gpurun_out/r2_neigh_launches.csv python bench.py --steps 1 --warmup 1 --melt 20 --no-extra --no-cpu --no-e2e --no-ab --no-checks > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_neigh_launches.csv")) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print(r[h.index("Kernel Name")][:32], r[h.index("Metric Name")][:30], r[h.index("Metric Value")])
PY
