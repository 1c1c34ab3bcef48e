#!/bin/bash
# 8-GPU box: NCCL parity at 8 ranks (full + half list), then both bench arms at N=8
mkdir -p gpurun_out
for h in "" "--half"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 tests/mp_parity.py $h 2>&1 | grep MP_PARITY | tee -a gpurun_out/mp_parity8_final.log
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29841 bench.py --gpus 8 --impl reference > gpurun_out/bench_n8_reference.json 2> gpurun_out/bench_n8_reference.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29842 bench.py --gpus 8 > gpurun_out/bench_n8_cuda.json 2> gpurun_out/bench_n8_cuda.err
python - <<PY
import json
for f in ("reference", "cuda"):
    try:
        d=json.loads(open('gpurun_out/bench_n8_%s.json' % f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms/step %.2f e2e %.4g" % (d['value'], d['ms_per_step'], d['e2e']['value']), d.get('time_buckets_ms'))
        print({k: (v.get('value') if isinstance(v, dict) else v) for k, v in (d.get('extra') or {}).items()})
        print(d.get('parity_vs_oracle'))
    except Exception as e:
        print(f, "failed", e)
PY
tail -c 400 gpurun_out/bench_n8_cuda.err
