#!/bin/bash
# 2-GPU check: rank parity against the oracle's virtual ranks (full + half), multi-GPU pytest, weak-scaling bench
mkdir -p gpurun_out
for h in "" "--half"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 tests/mp_parity.py $h 2>&1 | grep -E "MP_PARITY|Error|error" | tee -a gpurun_out/mp_parity2.log
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29802 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -2 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print("N=2 value %.4g ms/step %.2f buckets %s" % (d['value'], d['ms_per_step'], {k: round(v,1) for k,v in d['time_buckets_ms'].items()}))
PY
