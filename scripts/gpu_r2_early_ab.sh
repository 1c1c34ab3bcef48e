#!/bin/bash
# 2-GPU box: GPU suite, then A/B of the refresh overlap (beside the integrator vs beside the interior tiles)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_early.log
for early in 1 0; do
  CBMD_EARLY=$early timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29830+early)) bench.py --gpus 2 --no-extra --no-checks --no-cpu --no-ab > gpurun_out/early_$early.json 2> gpurun_out/early_$early.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/early_$early.json').read().strip().splitlines()[-1])
print("EARLY=$early value %.4g ms/step %.3f e2e %.4g force_kernel_ms %.4f buckets %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_launch_ms'], {k: round(v,2) for k,v in d['time_buckets_ms'].items()}))
PY
done
