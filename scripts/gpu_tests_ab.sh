#!/bin/bash
# parity tests incl. the grouped (8 lanes/atom) sweeps and the
# new file-format paths, then the A/B bench of the two sweep shapes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-extra --no-cpu --steps 3 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_ab.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_ab.json').read().strip().splitlines()[-1])
print("value %.4g  ms/step %.2f  e2e %.4g  force_kernel %.3f ms  frac %.3f" % (d["value"], d["ms_per_step"], (d["e2e"] or {}).get("value",0), d["roofline"]["avg_launch_ms"], d["roofline"]["frac"]))
print("buckets", {k: round(v,2) for k,v in d["time_buckets_ms"].items()})
print("extra", json.dumps(d["extra"], indent=1))
PY
