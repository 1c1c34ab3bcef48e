#!/usr/bin/env python
"""Multi-GPU parity check, launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/mp_parity.py [--half]

N NCCL ranks run the decomposed MD loop; rank 0 compares (a) the thermo trace and (b) the
per-atom state by global id against the CPU oracle run with N virtual ranks AND with one
rank (SURVEY.md test T7), and (c) per-rank ghost sets against the oracle's."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def simulate_rank(world, rank, device, uid, half=False, steps=45, cells=8, precision=64):
    """One rank of the decomposed MD loop; uid is the NCCL id or a cabanamd_b200.Hub.  Returns this
    rank's final state (owned atoms, ghosts, thermo trace)."""
    import argparse as _ap

    from bench import build_sim

    a = _ap.Namespace(cutoff=2.5, guess=50, precision=precision)
    sim = build_sim(a, cells, half, world, rank, uid, device)
    sim.setup()
    sim.record_thermo()
    sim.run(steps, 5)
    g = sim.ctx.get_atoms()
    nl = g["n_local"]
    mine = dict(id=g["id"][:nl], x=g["x"][:nl], v=g["v"][:nl], f=g["f"][:nl],
                ghost_id=g["id"][nl:], ghost_x=g["x"][nl:], thermo=np.array(sim.thermo))
    return sim, mine


def compare_with_oracle(gathered, world, half=False, steps=45, cells=8, precision=64, busy_ranks=0):
    """Every rank's state (list indexed by rank) against the CPU oracle run with `world` virtual
    ranks (thermo, per-id x/v/f, per-rank ghost sets) and with one rank (thermo).  Returns
    (ok, worst); `worst` maps check name -> largest deviation."""
    import oracle_lib as O
    from cabanamd_b200.capi import dims_create

    ok = True
    worst = {}
    # torchrun exports OMP_NUM_THREADS=1; ranks busy-waiting in a collective meanwhile keep a core each
    O.lib().orc_set_threads(max(1, (os.cpu_count() or 1) - busy_ranks))
    grid = dims_create(world)
    cells3 = tuple(cells * k for k in grid)
    for nr in sorted({world, 1}, reverse=True):
        ref = O.Sim(mass=[2.0], half=half).create_lattice_fcc(cells=cells3, nranks=nr).setup()
        ref.record_thermo()
        ref.run(steps, 5)
        tg, to = np.array(gathered[0]["thermo"]), np.array(ref.thermo())
        if nr != world and half:
            # the reference's half-list PE weighs cross-rank pairs by 0.5, so
            # it depends on the decomposition: compare T and KE only
            tg, to = tg[:, [0, 1, 3]], to[:, [0, 1, 3]]
        worst[f"thermo_vs_{nr}rank"] = float(np.abs(tg - to).max())
        if nr != world:
            continue  # ids are numbered per rank at creation: only thermo is comparable
        ids = np.concatenate([d["id"] for d in gathered])
        order = np.argsort(ids)
        if not np.array_equal(ids[order], np.arange(1, len(ids) + 1)):
            worst["atoms_lost_or_duplicated"] = 1.0
            ok = False
            continue
        rid, rx, rv, rf = [], [], [], []
        for rk in range(nr):
            d = ref.get(rk)
            n = d["n_local"]
            rid.append(d["id"][:n]); rx.append(d["x"][:n]); rv.append(d["v"][:n]); rf.append(d["f"][:n])
        ro = np.argsort(np.concatenate(rid))
        for key, ours, theirs in (("x", "x", rx), ("v", "v", rv), ("f", "f", rf)):
            A = np.concatenate([d[ours] for d in gathered])[order]
            B = np.concatenate(theirs)[ro]
            if key == "x":  # same atom may sit one box length apart before the next wrap
                L = np.array(cells3) * ref.a
                diff = np.abs(A - B)
                diff = np.minimum(diff, np.abs(diff - L))
                worst[f"x_vs_{nr}rank"] = float(diff.max())
            else:
                worst[f"{key}_vs_{nr}rank"] = float(np.abs(A - B).max() / np.abs(B).max())
        # ghost SETS per rank: (owner id, position) multiset equal to the oracle's
        for rk in range(world):
            d = ref.get(rk)
            n = d["n_local"]
            want = np.concatenate([d["id"][n:, None].astype(np.float64), d["x"][n:]], axis=1)
            have = np.concatenate([gathered[rk]["ghost_id"][:, None].astype(np.float64),
                                   gathered[rk]["ghost_x"]], axis=1)
            if want.shape != have.shape:
                worst[f"ghost_count_rank{rk}"] = float(abs(len(want) - len(have)))
                ok = False
                continue
            want = want[np.lexsort(want.T[::-1])]
            have = have[np.lexsort(have.T[::-1])]
            if not np.array_equal(want[:, 0], have[:, 0]):
                worst[f"ghost_ids_rank{rk}"] = 1.0
                ok = False
            worst[f"ghost_x_rank{rk}"] = float(np.abs(want[:, 1:] - have[:, 1:]).max())
    loose = precision == 32  # FP32 force sweep: float round-off in f, amplified over the run
    tol = dict(thermo=2e-5 if loose else 1e-9, x=1e-4 if loose else 1e-9, v=1e-3 if loose else 1e-8,
               f=1e-3 if loose else 1e-8, ghost_x=1e-4 if loose else 1e-9)
    for k, v in worst.items():
        t = tol.get(k.split("_vs_")[0].split("_rank")[0], 0.0)
        if not (v <= t):
            ok = False
    return ok, worst


def run_parity(world, rank, local, uid, half=False, steps=45, cells=8, precision=64):
    """Decomposed MD loop on `world` NCCL ranks vs the CPU oracle's virtual ranks.  Collective:
    every rank calls it (torch.distributed initialised, backend nccl).  Returns (ok, worst) on
    every rank."""
    import torch.distributed as dist

    sim, mine = simulate_rank(world, rank, local, uid, half, steps, cells, precision)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    flag = [True, {}]
    if rank == 0:
        flag = list(compare_with_oracle(gathered, world, half, steps, cells, precision, busy_ranks=world - 1))
    dist.broadcast_object_list(flag, src=0)
    sim.ctx.close()
    return flag[0], flag[1]


def run_parity_threads(world, device=0, half=False, steps=45, cells=8, precision=64, timeout=120.0):
    """The same check with the `world` ranks as host threads of THIS process sharing one GPU,
    over the in-process transport (cabanamd_b200.Hub) instead of NCCL: identical kernels, plans
    and message contents; only the carrier of the messages differs."""
    import threading

    import cabanamd_b200 as cb

    hub = cb.Hub(world, timeout)
    out = [None] * world
    err = [None] * world

    def work(rank):
        try:
            sim, mine = simulate_rank(world, rank, device, hub, half, steps, cells, precision)
            out[rank] = mine
            sim.ctx.close()
        except BaseException as e:  # noqa: BLE001 - reported by the caller
            err[rank] = e

    threads = [threading.Thread(target=work, args=(r,), name=f"rank{r}") for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    bad = [(r, e) for r, e in enumerate(err) if e is not None]
    if bad:
        raise RuntimeError("rank(s) failed: " + "; ".join(f"{r}: {e!r}" for r, e in bad))
    hub.close()
    return compare_with_oracle(out, world, half, steps, cells, precision)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--half", action="store_true")
    ap.add_argument("--steps", type=int, default=45)
    ap.add_argument("--cells", type=int, default=8, help="fcc cells per dim per rank")
    ap.add_argument("--precision", type=int, default=64)
    ap.add_argument("--threads", type=int, default=0,
                    help="run this many ranks as host threads of one process on cuda:0 (in-process hub)")
    args = ap.parse_args()
    if args.threads:
        ok, worst = run_parity_threads(args.threads, 0, args.half, args.steps, args.cells, args.precision)
        print(("MP_PARITY_OK " if ok else "MP_PARITY_FAIL ") + f"thread-ranks={args.threads} half={args.half} "
              + " ".join(f"{k}={v:.2e}" for k, v in worst.items()), flush=True)
        sys.exit(0 if ok else 1)

    import torch
    import torch.distributed as dist

    import cabanamd_b200 as cb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [cb.Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ok, worst = run_parity(world, rank, local, box[0], args.half, args.steps, args.cells, args.precision)
    if rank == 0:
        print(("MP_PARITY_OK " if ok else "MP_PARITY_FAIL ") + f"ranks={world} half={args.half} "
              + " ".join(f"{k}={v:.2e}" for k, v in worst.items()), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
