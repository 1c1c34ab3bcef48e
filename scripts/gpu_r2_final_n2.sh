#!/bin/bash
# 2-GPU box: whole GPU suite (NCCL 2-rank parity + the hub's 2/4/8 thread-ranks on one GPU), then both
# bench arms at N=2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_n2.log
for impl in reference cuda; do
  extra=""; [ $impl = reference ] && extra="--impl reference"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29822 bench.py --gpus 2 $extra > gpurun_out/bench_n2_$impl.json 2> gpurun_out/bench_n2_$impl.err
  tail -c 1500 gpurun_out/bench_n2_$impl.json
done
