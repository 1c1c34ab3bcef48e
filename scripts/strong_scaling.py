#!/usr/bin/env python
"""BASELINE.json configs[3]: LJ 16 M atoms strong scaling on 8 GPUs (2x2x2), 1000 steps,
thermo every 10, energy-drift record.  torchrun --nproc-per-node 8 scripts/strong_scaling.py
Prints one JSON line on rank 0 (also usable at 1/2/4 GPUs with --cells-total)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells-total", type=int, default=160, help="fcc cells per dim of the WHOLE box")
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--half", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    import cabanamd_b200 as cb
    import bench
    from cabanamd_b200.capi import dims_create

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
    grid = dims_create(world)
    assert all(a.cells_total % g == 0 for g in grid)
    # build_sim takes cells per GPU per dim times the grid: only cubic-per-rank boxes
    per = [a.cells_total // g for g in grid]
    assert per[0] == per[1] == per[2] or world in (2, 4), per
    args = argparse.Namespace(cutoff=2.5, guess=50)
    # non-cubic sub-boxes (2 or 4 ranks): emulate by a global box of cells_total^3
    sim = build_custom(args, a.cells_total, grid, a.half, world, rank, uid, local)
    sim.setup()
    sim.record_thermo()
    sim.run(100, 10)  # warm-up + melt
    dctx = True if world > 1 else None
    stream = torch.cuda.ExternalStream(sim.ctx.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bench.barrier(dctx)
    sim.ctx.sync()
    e0.record(stream)
    sim.run(a.steps, 10)
    sim.ctx.sync()
    e1.record(stream)
    torch.cuda.synchronize()
    sec = bench.max_over_ranks(e0.elapsed_time(e1) * 1e-3, dctx)
    if rank == 0:
        th = np.array(sim.thermo)
        etot = th[:, 2] + th[:, 3]
        run = etot[11:]  # the timed 1000 steps
        print(json.dumps({
            "config": f"LJ {sim.N} atoms strong scaling on {world} GPU(s) ({'x'.join(map(str, grid))}), "
                      f"{a.steps} steps, thermo/10, {'half' if a.half else 'full'} list",
            "n_gpus": world, "atoms": sim.N, "steps": a.steps, "seconds": sec,
            "atom_steps_per_s": sim.N * a.steps / sec, "ms_per_md_step": sec / a.steps * 1e3,
            "etot_first": float(run[0]), "etot_last": float(run[-1]),
            "energy_drift_per_atom": float(run[-1] - run[0]),
            "energy_rms_fluct": float(run.std()), "T_last": float(th[-1, 1]),
            "step0": [float(v) for v in th[0]],
        }), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def build_custom(args, cells_total, grid, half, nranks, rank, uid, device, temp=1.4, seed=87287):
    import bench
    from cabanamd_b200.harness import Simulation, create_velocities

    cells_global = (cells_total,) * 3
    a = (4.0 / 0.8442) ** (1.0 / 3.0)
    sim = Simulation(device=device, mass=(2.0,), cut=args.cutoff, skin=0.3, half=half,
                     exchange_rate=20, max_neigh_guess=args.guess, nranks=nranks, rank=rank, uid=uid)
    sim.set_box([0.0] * 3, [a * c for c in cells_global])
    x = bench.local_lattice(sim, cells_global, a)
    t = np.zeros(len(x), dtype=np.int32)
    n_before = sim.ctx.scan_sum_int(len(x)) - len(x) if nranks > 1 else 0
    ids = np.arange(1, len(x) + 1, dtype=np.int32) + n_before
    v = create_velocities(sim, x, t, temp, seed)
    sim.set_atoms(x, v, t, ids)
    T = sim.temperature()
    sim.ctx.set_velocities(v * np.sqrt(temp / T))
    assert sim.N == 4 * cells_total ** 3
    return sim


if __name__ == "__main__":
    main()
