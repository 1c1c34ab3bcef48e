#!/bin/bash
# quick GPU regression: parity tests + short bench (no extra legs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --no-extra --no-cpu ${BENCH_ARGS} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print("value %.4g  ms/step %.2f  e2e %.4g  force_kernel %.3f ms  frac %.3f" % (d["value"], d["ms_per_step"], (d["e2e"] or {}).get("value",0), d["roofline"]["avg_launch_ms"], d["roofline"]["frac"]))
print("buckets", {k: round(v,2) for k,v in d["time_buckets_ms"].items()})
PY
