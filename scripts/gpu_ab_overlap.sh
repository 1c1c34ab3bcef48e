for cfg in "1 10" "1 0" "0 10" "0 0"; do set -- $cfg
CBMD_OVERLAP=$1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29850 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --thermo $2 --melt 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('overlap=$1 thermo=$2 ms/step %.2f buckets %s' % (d['ms_per_step'], {k: round(v,1) for k,v in d['time_buckets_ms'].items()}))"
done
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-extra --no-cpu --thermo 0 --melt 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=1 thermo=0 ms/step %.2f buckets %s' % (d['ms_per_step'], {k: round(v,1) for k,v in d['time_buckets_ms'].items()}))"
