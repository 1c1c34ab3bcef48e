#!/bin/bash
# 4-GPU weak-scaling point ({2,2,1} decomposition, 4 M atoms per GPU)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29804 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
tail -2 gpurun_out/bench_n4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1])
print("N=4 value %.4g ms/step %.2f buckets %s" % (d['value'], d['ms_per_step'], {k: round(v,1) for k,v in d['time_buckets_ms'].items()}))
PY
