#!/bin/bash
# multi-GPU parity + weak-scaling bench; run under gpurun --gpus 8 (or fewer: NLIST)
mkdir -p gpurun_out
for n in ${NLIST:-2 4 8}; do
  for h in "" "--half"; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) tests/mp_parity.py $h 2>&1 | grep MP_PARITY
  done
done
for n in ${NLIST:-2 4 8}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800+n)) bench.py --gpus $n --steps ${STEPS:-3} --warmup 3 --no-e2e ${BENCH_ARGS} > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
print("N=$n value %.4g ms/step %.2f buckets %s" % (d['value'], d['ms_per_step'], {k: round(v,1) for k,v in d['time_buckets_ms'].items()}))
PY
done
