#!/bin/bash
# A/B of NCCL launch geometry for the overlapped halo (2 GPUs)
run() {
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N:-2} --master-addr 127.0.0.1 --master-port 29850 bench.py --gpus ${N:-2} --steps 3 --warmup 3 --no-e2e --melt 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*  ms/step %.2f buckets %s' % (d['ms_per_step'], {k: round(v,1) for k,v in d['time_buckets_ms'].items()}))"
}
run CBMD_OVERLAP=1
run CBMD_OVERLAP=1 NCCL_NTHREADS=128
run CBMD_OVERLAP=1 NCCL_NTHREADS=128 NCCL_MAX_NCHANNELS=2
run CBMD_OVERLAP=1 NCCL_NTHREADS=64 NCCL_MAX_NCHANNELS=4
run CBMD_OVERLAP=1 NCCL_MAX_NCHANNELS=1
