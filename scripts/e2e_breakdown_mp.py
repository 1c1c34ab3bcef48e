#!/usr/bin/env python
"""Where one end-to-end bench step spends its time at N ranks (torchrun): host wall clock per
stage with a stream sync after each.  python -m torch.distributed.run --nproc-per-node N ... scripts/e2e_breakdown_mp.py"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    import torch
    import torch.distributed as dist

    import cabanamd_b200 as cb
    from bench import build_sim, MD_PER_STEP
    from cabanamd_b200.capi import _dp

    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    uid = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [cb.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    args = argparse.Namespace(cutoff=2.5, guess=50)
    sim = build_sim(args, a.cells, False, world, rank, uid, local)
    sim.setup()
    sim.run(40, 0)
    ctx = sim.ctx
    g = ctx.get_atoms(fields="xvti")
    nl = g["n_local"]

    def pinned(x):
        t = torch.empty(x.shape, dtype=torch.from_numpy(x).dtype, pin_memory=True)
        t.numpy()[...] = x
        return t

    hx, hv, ht, hi = pinned(g["x"][:nl]), pinned(g["v"][:nl]), pinned(g["type"][:nl]), pinned(g["id"][:nl])
    stages = ["set_atoms", "exchange", "bin_sort", "exchange_halo", "neigh_build", "force", "run20", "get_atoms"]
    acc = {s: 0.0 for s in stages}

    def timed(name, fn):
        t0 = time.perf_counter()
        fn()
        ctx.sync()
        acc[name] += time.perf_counter() - t0

    for rep in range(a.reps + 1):
        if rep == 1:
            acc = {s: 0.0 for s in stages}
        if world > 1:
            dist.barrier()
        timed("set_atoms", lambda: ctx.set_atoms(hx.numpy(), hv.numpy(), None, ht.numpy(), hi.numpy()))
        timed("exchange", ctx.exchange)
        timed("bin_sort", lambda: ctx.bin_sort(sim.rn))
        timed("exchange_halo", lambda: ctx.exchange_halo(sim.rn))
        timed("neigh_build", lambda: ctx.neigh_build(sim.rn, False, 0, sim.guess))
        timed("force", lambda: (ctx.zero_force(), ctx.force(False)))
        sim.step = 0
        timed("run20", lambda: sim.run(MD_PER_STEP, 10))
        timed("get_atoms", lambda: ctx._ck(ctx.L.cbmd_get_atoms(ctx.h, 0, nl, _dp(hx.numpy()), _dp(hv.numpy()), None, None, None, None)))
    line = f"rank {rank}/{world} ms per e2e step: " + "  ".join(f"{s} {acc[s] / a.reps * 1e3:.2f}" for s in stages) \
        + f"  total {sum(acc.values()) / a.reps * 1e3:.2f}  cpus {sorted(os.sched_getaffinity(0))[:4]}..({len(os.sched_getaffinity(0))})"
    print(line, flush=True)
    if world > 1:
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
