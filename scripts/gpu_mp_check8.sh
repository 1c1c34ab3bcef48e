#!/bin/bash
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29601 tests/mp_parity.py 2>&1 | grep MP_PARITY
$TR --master-port 29611 tests/mp_parity.py --half 2>&1 | grep MP_PARITY
for st in 1 3; do
  CBMD_HALO_STAGES=$st $TR --master-port 2962$st bench.py --gpus $N --no-extra --no-cpu --no-e2e --no-ab --no-checks 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('halo_stages $st', '%.4e'%d['value'], {a:round(b,2) for a,b in d['time_buckets_ms'].items()})"
done
