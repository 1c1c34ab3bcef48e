#!/bin/bash
# 8-GPU box: parity at 8 ranks, weak-scaling bench at 4 and 8, 16 M strong scaling + drift
mkdir -p gpurun_out
for h in "" "--half"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 tests/mp_parity.py $h 2>&1 | grep MP_PARITY | tee -a gpurun_out/mp_parity8.log
done
for n in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800+n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
print("N=$n value %.4g ms/step %.2f e2e %.4g buckets %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v,1) for k,v in d['time_buckets_ms'].items()}))
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29900 scripts/strong_scaling.py --cells-total 160 --steps 1000 2>gpurun_out/strong8.err | tee gpurun_out/strong8.json
