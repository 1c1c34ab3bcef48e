#!/usr/bin/env python
"""bench.py — LJ atom-timesteps/s of the B200-native short-range MD step.

    python bench.py --gpus N --steps K --warmup W            (our CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm)

One bench "step" = one neighbour-list period of the MD loop = `exchange_rate` (20)
MD timesteps: 1 rebuild step (migrate/wrap, cell sort, ghost build, Verlet build) +
19 ghost-refresh steps, each with LJ force + both velocity-Verlet half steps, with
the reference's thermo output (T, PE, KE) every 10 MD steps, exactly the call order of
CbnMD::run (reference src/cabanamd_impl.h:285-399).  Workload: BASELINE.json
configs[2], 4 000 000 atoms per GPU (fcc 100^3 cells, rho*=0.8442, rc=2.5, skin 0.3,
full neighbour list, FP64), which at N=1 is the largest single-GPU configuration;
the other configs ride along under "extra": configs[1] (1 M atoms, full vs half list),
configs[3] (16.4 M atoms strong scaling + the 1000-step energy-drift check against the
oracle), configs[4] (rc = 5.0) and the FP32 force variant.  Data are synthetic: the deck
initialiser's lattice + hashed-RNG velocities, melted for 200 untimed MD steps so the timed
state is the LJ liquid.

The CPU oracle is only used here as the checker / baseline: the `cpu_baseline` leg,
`--impl reference`, and the parity / drift checks that ride along at N >= 1.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "LJ atom-timesteps/s"
UNIT = "atom-steps/s"
MD_PER_STEP = 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=100, help="fcc cells per dim per GPU (100 -> 4 M atoms)")
    ap.add_argument("--half", action="store_true", help="half neighbour list (Newton 3)")
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32],
                    help="arithmetic of the full-list force sweep (32 = FP32 variant)")
    ap.add_argument("--melt", type=int, default=200, help="untimed MD steps before warm-up")
    ap.add_argument("--thermo", type=int, default=10, help="thermo every n MD steps (0 = never)")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[1]/[3]/[4] and FP32 legs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ab", action="store_true", help="skip the A/B leg (gather=0)")
    ap.add_argument("--no-checks", action="store_true", help="skip the parity / drift checks vs the oracle")
    ap.add_argument("--cpu-cells", type=int, default=63, help="cpu_baseline sample: cells per dim")
    ap.add_argument("--strong-cells", type=int, default=160, help="configs[3]: global fcc cells per dim")
    ap.add_argument("--strong-steps", type=int, default=1000, help="configs[3]: MD steps")
    ap.add_argument("--cutoff", type=float, default=2.5)
    ap.add_argument("--guess", type=int, default=50)
    return ap.parse_args()


# --------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.index = index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
                for k, nm in enumerate(names):
                    if r[5 + k].strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


def local_lattice(sim, cells_global, a):
    """This rank's share of the fcc lattice (create_lattice, inputFile_impl.h:716-792):
    generate the cells overlapping the sub-box, keep lo <= x < hi."""
    d = sim.dom
    lo_c = np.floor(d["llo"] / a).astype(int) - 1
    hi_c = np.ceil(d["lhi"] / a).astype(int) + 1
    lo_c = np.maximum(lo_c, 0)
    hi_c = np.minimum(hi_c, np.array(cells_global))
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=np.float64)
    iz, iy, ix = np.meshgrid(np.arange(lo_c[2], hi_c[2]), np.arange(lo_c[1], hi_c[1]),
                             np.arange(lo_c[0], hi_c[0]), indexing="ij")
    cell = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(np.float64)
    x = (a * (1.0 * cell[:, None, :] + basis[None, :, :])).reshape(-1, 3)
    ghi = d["ghi"]
    m = np.all((x >= d["llo"]) & (x < d["lhi"]) & (x < ghi), axis=1)
    return np.ascontiguousarray(x[m])


def build_sim(args, cells_per_gpu, half, nranks, rank, uid, device, temp=1.4, seed=87287,
              cells_global=None, fast_velocities=False):
    """One rank of the deck's initial state.  cells_global (3-tuple) overrides the weak-scaling
    cells_per_gpu * grid; fast_velocities swaps the reference's hashed per-atom RNG (28 hash
    rounds per atom, vectorised numpy) for a seeded numpy draw — same distribution, only for
    throughput-only legs on very large systems."""
    from cabanamd_b200.capi import dims_create
    from cabanamd_b200.harness import Simulation, create_velocities

    grid = dims_create(nranks)
    if cells_global is None:
        cells_global = tuple(cells_per_gpu * g for g in grid)
    a = (4.0 / 0.8442) ** (1.0 / 3.0)
    sim = Simulation(device=device, mass=(2.0,), cut=args.cutoff, skin=0.3, half=half,
                     exchange_rate=MD_PER_STEP, max_neigh_guess=args.guess, nranks=nranks,
                     rank=rank, uid=uid, precision=getattr(args, "precision", 64))
    sim.set_box([0.0] * 3, [a * c for c in cells_global])
    x = local_lattice(sim, cells_global, a)
    t = np.zeros(len(x), dtype=np.int32)
    n_before = sim.ctx.scan_sum_int(len(x)) - len(x) if nranks > 1 else 0
    ids = np.arange(1, len(x) + 1, dtype=np.int32) + n_before
    if fast_velocities:
        v = (np.random.default_rng(seed + rank).random((len(x), 3)) - 0.5) / np.sqrt(2.0)
        tot = np.array([2.0 * len(x), *(2.0 * v).sum(0)])
        if nranks > 1:
            tot = np.array([sim.ctx.reduce_sum(q) for q in tot])
        v = v - tot[1:] / tot[0]
    else:
        v = create_velocities(sim, x, t, temp, seed)
    sim.set_atoms(x, v, t, ids)
    # rescale to the target temperature (inputFile_impl.h:851-865)
    T = sim.temperature()
    sim.ctx.set_velocities(v * np.sqrt(temp / T))
    expect = 4 * cells_global[0] * cells_global[1] * cells_global[2]
    assert sim.N == expect, (sim.N, expect)
    return sim


IN_LJ = """# 3d Lennard-Jones melt (the reference's input/in.lj with the box size and run length substituted)
units           lj
atom_style      atomic
newton          off
lattice         fcc 0.8442
region          box block 0 {c} 0 {c} 0 {c}
create_box      1 box
create_atoms    1 box
mass            1 2.0
velocity        all create 1.4 87287 loop geom
pair_style      lj/cut 2.5
pair_coeff      1 1 1.0 1.0 2.5
neighbor        0.3 bin
neigh_modify    every 20 one 50
comm_modify     cutoff * 20
fix             1 all nve
thermo          10
run             {steps}
"""


def cbnmd_leg(cells, steps):
    """Run the C++ driver on an in.lj deck and report what ITS summary prints (wall clock of the
    step loop, reference format `#Steps/s Atomsteps/s Atomsteps/(proc*s)`, cabanamd_impl.h:418-428)."""
    import subprocess
    import tempfile

    exe = os.path.join(ROOT, "cabanamd_b200", "lib", "cbnMD")
    if not os.path.exists(exe):
        return None
    with tempfile.TemporaryDirectory() as td:
        deck = os.path.join(td, "in.lj")
        with open(deck, "w") as f:
            f.write(IN_LJ.format(c=cells, steps=steps))
        out = os.path.join(td, "md.out")
        try:
            p = subprocess.run([exe, "-il", deck, "-o", out, "-e", os.path.join(td, "md.err")],
                               capture_output=True, text=True, cwd=td, timeout=600)
        except subprocess.TimeoutExpired:
            return {"error": "timeout"}
        txt = open(out).read() if os.path.exists(out) else ""
        if p.returncode != 0 or "#Steps/s" not in txt:
            return {"error": (p.stderr or txt)[-300:]}
        lines = txt.splitlines()
        k = max(i for i, ln in enumerate(lines) if ln.startswith("#Steps/s"))
        sps, aps, _ = (float(v) for v in lines[k + 1].split())
        thermo = [ln.split() for ln in lines[:k] if ln and ln[0].isdigit() and len(ln.split()) == 6]
        res = {"value": aps, "unit": UNIT, "steps_per_s": sps, "atoms": 4 * cells ** 3,
               "timing": "host wall clock around the step loop, as the reference prints it"}
        if thermo:
            res["etot_first_last"] = [float(thermo[0][3]), float(thermo[-1][3])]
        return res


def timed_steps(sim, md_steps, thermo, dist_ctx):
    """md_steps MD steps bracketed by barrier + synchronize, CUDA events on the context
    stream; returns seconds (max over ranks)."""
    import torch

    stream = torch.cuda.ExternalStream(sim.ctx.stream())
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier(dist_ctx)
    sim.ctx.sync()
    torch.cuda.synchronize()
    e0.record(stream)
    sim.run(md_steps, thermo)
    sim.ctx.sync()  # also applies the deferred final_integrate of the last step
    e1.record(stream)
    torch.cuda.synchronize()
    barrier(dist_ctx)
    sec = e0.elapsed_time(e1) * 1e-3
    return max_over_ranks(sec, dist_ctx)


def barrier(dist_ctx):
    if dist_ctx:
        import torch.distributed as dist

        dist.barrier()


def _reduce(val, dist_ctx, op):
    if not dist_ctx:
        return val
    import torch
    import torch.distributed as dist

    t = torch.tensor([val], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
    return float(t.item())


def max_over_ranks(val, dist_ctx):
    return _reduce(val, dist_ctx, "MAX")


def sum_over_ranks(val, dist_ctx):
    return _reduce(val, dist_ctx, "SUM")


def force_bytes(n_local, n_ghost, nn, half, precision=64):
    """Algorithmic bytes of one LJ force launch (SURVEY.md 8d / DESIGN.md): int32 indices, every
    position read once, forces written once; FP64: 24-byte positions/forces, FP32 variant: 12."""
    w = 24.0 if precision == 64 else 12.0
    if half:
        return n_local * (4.0 * nn + 8.0) + (n_local + n_ghost) * (w + 4.0 + 2.0 * w)
    return n_local * (4.0 * nn + w + 8.0) + (n_local + n_ghost) * (w + 4.0)


def measure_resident(args, sim, steps, warmup, dist_ctx, sample_clocks, melt=None):
    sim.run(args.melt if melt is None else melt, 0)
    # re-anchor the rebuild cadence so every bench step holds exactly one rebuild
    sim.step = 0
    sim.run(warmup * MD_PER_STEP, args.thermo)
    ctx = sim.ctx
    ctx.timing_enable(True)
    ctx.timing_reset()
    l0 = ctx.launch_count()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))) if sample_clocks else None
    if sampler:
        sampler.start()
    sec = timed_steps(sim, steps * MD_PER_STEP, args.thermo, dist_ctx)
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count() - l0
    tm = ctx.timing()
    ctx.timing_enable(False)
    nl, ng = ctx.counts()
    tot, mx = ctx.neigh_sizes()
    return dict(sec=sec, launches=launches, timing=tm, n_local=nl, n_ghost=ng,
                nn=tot / max(nl, 1), max_neigh=mx, clocks=clocks)


def roofline_of(m, half, precision, peak, peak_src, kernel):
    fk_ms, fk_n = m["timing"]["force_kernel"]
    fb = force_bytes(m["n_local"], m["n_ghost"], m["nn"], half, precision)
    avg = fk_ms / max(fk_n, 1)
    achieved = fb / (avg * 1e-3) / 1e9 if fk_ms > 0 else 0.0
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": peak_src, "traffic": None,
            "bytes_per_launch": fb, "avg_launch_ms": avg, "launches": fk_n,
            "stored_neighbours_per_atom": m["nn"], "ghost_fraction": m["n_ghost"] / max(m["n_local"], 1),
            "step_share": fk_ms * 1e-3 / m["sec"]}


def measure_e2e(args, sim, steps, dist_ctx):
    """Same metric through the public C ABI with HOST buffers: every bench step uploads
    this rank's atoms (x, v, type, id) from pinned host memory, runs the init path + one list
    period, and downloads x, v, type, id (the atoms a rank owns change with migration, so the
    whole rows come back) and the thermo scalars.  Returns seconds (max over ranks), bytes and
    a stage breakdown (host clock with a stream sync after each stage, rank 0)."""
    import torch

    from cabanamd_b200.capi import _dp, _ip

    ctx = sim.ctx
    nl0, _ = ctx.counts()
    cap = nl0 + nl0 // 8 + 1024  # owned atoms drift a little between ranks

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype, pin_memory=True)

    hx, hv = pinned((cap, 3), torch.float64), pinned((cap, 3), torch.float64)
    ht, hi = pinned((cap,), torch.int32), pinned((cap,), torch.int32)
    thermo = args.thermo
    stream = torch.cuda.ExternalStream(ctx.stream())
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    state = {"n": 0, "h2d": 0, "d2h": 0}
    stage = {"h2d_upload": 0.0, "init_path": 0.0, "md_steps": 0.0, "d2h_download": 0.0}

    def download():
        n, _ = ctx.counts()
        assert n <= cap
        ctx._ck(ctx.L.cbmd_get_atoms(ctx.h, 0, n, _dp(hx.numpy()), _dp(hv.numpy()), None,
                                     _ip(ht.numpy()), _ip(hi.numpy()), None))
        state["n"] = n

    def one_step(record):
        n = state["n"]
        t0 = time.perf_counter()
        ctx.set_atoms(hx.numpy()[:n], hv.numpy()[:n], None, ht.numpy()[:n], hi.numpy()[:n])
        t1 = time.perf_counter()
        sim.setup()
        if record:
            ctx.sync()
        t2 = time.perf_counter()
        sim.run(MD_PER_STEP, thermo)
        if record:
            ctx.sync()
        t3 = time.perf_counter()
        download()
        t4 = time.perf_counter()
        state["h2d"] += n * (24 + 24 + 4 + 4)
        state["d2h"] += state["n"] * (24 + 24 + 4 + 4)
        if record:
            for k, dt in zip(stage, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                stage[k] += dt

    download()
    one_step(False)  # warm-up
    state["h2d"] = state["d2h"] = 0
    barrier(dist_ctx)
    ctx.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        one_step(False)
    ctx.sync()
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    barrier(dist_ctx)
    sec = max(e0.elapsed_time(e1) * 1e-3, wall)
    sec = max_over_ranks(sec, dist_ctx)
    n_thermo = (MD_PER_STEP // thermo) if thermo else 0
    h2d = state["h2d"] / steps
    d2h = state["d2h"] / steps + n_thermo * 3 * 8
    one_step(True)  # untimed: stage breakdown
    barrier(dist_ctx)
    return sec, sum_over_ranks(h2d, dist_ctx), sum_over_ranks(d2h, dist_ctx), \
        {k: v * 1e3 for k, v in stage.items()}


def oracle_threads(busy_ranks=0):
    """All host cores for the oracle (torchrun exports OMP_NUM_THREADS=1), less one per rank that
    busy-waits in a collective meanwhile: OpenMP barriers spin, and spinning on oversubscribed
    cores made the 8-virtual-rank drift check 15x slower than the 1-rank one (196 s against 13 s)."""
    import oracle_lib as O

    L = O.lib()
    L.orc_set_threads(max(1, (os.cpu_count() or 1) - busy_ranks))
    return L.orc_max_threads()


def cpu_baseline(cells, md_steps, half=False):
    """The oracle (a port of the reference algorithm: the real Kokkos/Cabana build is not
    available in this image) on the host cores; bounded sample."""
    import oracle_lib as O

    nthr = oracle_threads()
    s = O.Sim(mass=[2.0], half=half).create_lattice_fcc(cells=(cells,) * 3).setup()
    s.run(MD_PER_STEP, 10)  # warm: first rebuild
    t0 = time.perf_counter()
    s.run(md_steps, 10)
    dt = time.perf_counter() - t0
    return dict(value=s.natoms * md_steps / dt, unit=UNIT, cores=nthr, kind="port",
                sample=f"oracle (C++/OpenMP port of the reference step), fcc {cells}^3 cells = "
                       f"{s.natoms} atoms x {md_steps} MD steps, thermo/10, {nthr} threads",
                seconds=dt, timers=s.timers())


def run_reference(args):
    """CPU arm: the oracle port on all host cores (rank 0 only).  Each step = 20 MD steps of ONE
    GPU's share of the workload (the whole N-GPU system would not finish in minutes on the
    host); the sample shrinks further when steps+warmup is large."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O

    nthr = oracle_threads()
    budget = 3.0e9  # atom-steps: ~100 s at 3e7 atom-steps/s
    cells = min(args.cells, int((budget / (MD_PER_STEP * (args.steps + args.warmup)) / 4.0) ** (1.0 / 3.0)))
    cells = max(cells, 10)
    s = O.Sim(mass=[2.0], half=args.half).create_lattice_fcc(cells=(cells,) * 3).setup()
    for _ in range(args.warmup):
        s.run(MD_PER_STEP, args.thermo)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.run(MD_PER_STEP, args.thermo)
    dt = time.perf_counter() - t0
    val = s.natoms * args.steps * MD_PER_STEP / dt
    sample = (f"each step = {MD_PER_STEP} MD steps of a {s.natoms}-atom sample (fcc {cells}^3 = "
              f"{'one GPU share' if cells == args.cells else 'a bounded part of one GPU share'} of the "
              f"{args.gpus}-GPU workload), {nthr} OpenMP threads")
    cfg = workload_config(args, args.gpus)
    cfg["workload"] += f"; CPU arm runs a sample: {s.natoms} atoms on {nthr} host threads"
    cfg["cpu_sample_atoms"] = s.natoms
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthr, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU arm: oracle/ (OpenMP port of the reference algorithm); the Kokkos+Cabana+MPI "
                "reference cannot be built in this image",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    atoms = 4 * args.cells ** 3
    return {
        "workload": f"LJ liquid, {atoms} atoms/GPU (fcc {args.cells}^3 cells/GPU, rho*=0.8442, "
                    f"rc={args.cutoff}, skin 0.3, {'half' if args.half else 'full'} neighbour list, "
                    f"rebuild every {MD_PER_STEP}) — BASELINE.json configs[2]",
        "atoms_per_gpu": atoms, "atoms_total": atoms * n,
        "md_steps_per_bench_step": MD_PER_STEP, "thermo_every": args.thermo,
        "neighbor_list": "half" if args.half else "full",
        "force_precision": args.precision,
        "decomposition": {1: "1x1x1", 2: "2x1x1", 4: "2x2x1", 8: "2x2x2"}.get(n, str(n)),
        "l2_policy": "working set (x,v,f + neighbour table, >1 GB at 4 M atoms) exceeds the 126 MB L2; no flush",
    }


def fresh_uid(dist_ctx, rank):
    """Every context group needs its own NCCL unique id."""
    if not dist_ctx:
        return None
    import torch.distributed as dist

    import cabanamd_b200 as cb

    box = [cb.Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def drift_check(args, n, rank, local, dist_ctx, cells=40, md_steps=1000):
    """configs[3]'s 1000-step energy-drift check at a size the host can follow: the in.lj deck
    (fcc 40^3 = 256 000 atoms, thermo every 10) on the N GPUs and on the oracle with N virtual
    ranks.  Trajectories are chaotic, so the per-step traces are compared over the first 100
    steps and the drift / fluctuation of the total energy over the whole run."""
    from cabanamd_b200.capi import dims_create

    a = argparse.Namespace(**vars(args))
    a.precision, a.cutoff, a.guess = 64, 2.5, 50
    grid = dims_create(n)
    if any(cells % g for g in grid):
        return None
    sim = build_sim(a, 0, False, n, rank, fresh_uid(dist_ctx, rank), local, cells_global=(cells,) * 3)
    sim.setup()
    sim.record_thermo()
    sim.run(md_steps, 10)
    th = np.array(sim.thermo)
    sim.ctx.close()
    out = None
    if rank == 0:
        import oracle_lib as O

        nthr = oracle_threads(busy_ranks=n - 1)  # the other ranks spin in the collective that follows
        t0 = time.perf_counter()
        ref = O.Sim(mass=[2.0]).create_lattice_fcc(cells=(cells,) * 3, nranks=n).setup()
        ref.record_thermo()
        ref.run(md_steps, 10)
        to = np.array(ref.thermo())
        eg, eo = th[:, 2] + th[:, 3], to[:, 2] + to[:, 3]
        k = 11  # rows of steps 0..100
        out = {"deck": f"in.lj (fcc {cells}^3 = {4 * cells ** 3} atoms, {md_steps} steps, thermo/10), "
                       f"{n} GPU rank(s) vs oracle with {n} virtual rank(s)",
               "thermo_maxdiff_first_100_steps": float(np.abs(th[:k, 1:] - to[:k, 1:]).max()),
               "etot_drift_gpu": float(eg[-1] - eg[0]), "etot_drift_oracle": float(eo[-1] - eo[0]),
               "etot_maxdev_gpu": float(np.abs(eg - eg[0]).max()), "etot_maxdev_oracle": float(np.abs(eo - eo[0]).max()),
               "etot_rms_gpu": float(np.std(eg)), "etot_rms_oracle": float(np.std(eo)),
               "T_final_gpu": float(th[-1, 1]), "T_final_oracle": float(to[-1, 1]),
               "oracle_seconds": time.perf_counter() - t0, "oracle_threads": nthr}
        out["ok"] = bool(out["thermo_maxdiff_first_100_steps"] < 1e-8 and
                         abs(out["etot_maxdev_gpu"] - out["etot_maxdev_oracle"]) < 0.5 * max(out["etot_maxdev_oracle"], 1e-6)
                         and abs(out["T_final_gpu"] - out["T_final_oracle"]) < 0.01)
    return out


def strong_scaling_leg(args, n, rank, local, dist_ctx):
    """configs[3]: 16.4 M atoms (fcc 160^3) on the N GPUs, strong scaling (total work fixed),
    1000 MD steps with thermo every 10; reports throughput and the energy drift of that run."""
    from cabanamd_b200.capi import dims_create

    grid = dims_create(n)
    c = args.strong_cells
    if any(c % g for g in grid):
        return None
    a = argparse.Namespace(**vars(args))
    a.precision, a.cutoff, a.guess, a.half = 64, 2.5, 50, False
    sim = build_sim(a, 0, False, n, rank, fresh_uid(dist_ctx, rank), local, cells_global=(c,) * 3,
                    fast_velocities=True)
    sim.setup()
    sim.run(100, 0)  # melt + warm
    sim.step = 0
    sim.thermo = []
    sim.record_thermo()
    sec = timed_steps(sim, args.strong_steps, 10, dist_ctx)
    th = np.array(sim.thermo)
    e = th[:, 2] + th[:, 3]
    nl, ng = sim.ctx.counts()
    res = {"value": sim.N * args.strong_steps / sec, "unit": UNIT, "scaling": "strong", "atoms_total": sim.N,
           "atoms_per_gpu": sim.N // n, "md_steps": args.strong_steps, "ms_per_md_step": sec / args.strong_steps * 1e3,
           "decomposition": "x".join(str(g) for g in grid), "ghost_fraction_rank0": ng / max(nl, 1),
           "etot_drift": float(e[-1] - e[0]), "etot_maxdev": float(np.abs(e - e[0]).max()),
           "note": "BASELINE configs[3]; velocities from a seeded numpy draw (throughput leg), melted 100 steps; "
                   "the drift check against the oracle is extra['configs[3] drift check']"}
    sim.ctx.close()
    return res


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference "
                         "for the CPU arm")
    torch.cuda.set_device(local)
    dist_ctx = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist_ctx = True
    n = world
    assert n == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    # ---- correctness first: the decomposed loop against the oracle's virtual ranks (N > 1)
    parity = None
    if n > 1 and not args.no_checks:
        from mp_parity import run_parity

        parity = {}
        for half in (False, True):
            ok, worst = run_parity(n, rank, local, fresh_uid(dist_ctx, rank), half=half, steps=45, cells=8)
            parity["half" if half else "full"] = {"ok": bool(ok), **worst}
        parity["what"] = (f"{n} NCCL ranks, fcc 8^3 cells/rank, 45 MD steps (2 rebuilds), thermo + per-id x/v/f + per-rank "
                          f"ghost sets vs the oracle with {n} virtual ranks (tests/mp_parity.py)")

    sim = build_sim(args, args.cells, args.half, n, rank, fresh_uid(dist_ctx, rank), local)
    sim.setup()
    m = measure_resident(args, sim, args.steps, args.warmup, dist_ctx, sample_clocks=True)
    md_steps = args.steps * MD_PER_STEP
    atoms_total = sim.N
    value = atoms_total * md_steps / m["sec"]

    # roofline of the dominant kernel (LJ force), measured live with CUDA events
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"
    gather = int(os.environ.get("CBMD_GATHER", "1"))
    if args.half:
        kname = "k_force_half"
    elif args.precision == 32:
        kname = "k_force_full_f32 (float4 LDG.128, index stream through TEX)"
    else:
        kname = "k_force_full<1,...> (xy LDG.128 + z TEX)" if gather == 1 else "k_force_full<0,...> (32-byte records)"
    roofline = roofline_of(m, args.half, args.precision, peak, peak_src, kname)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "force_traffic.json")))
        key = f"{'half' if args.half else 'full'}_{args.cells}" + ("_f32" if args.precision == 32 else "")
        if key in tr:
            roofline["traffic"] = tr[key]
            roofline["traffic_source"] = tr.get("source", "ncu --set full capture, profiles/force_traffic.json (not re-measured in this run)")
    except Exception:
        pass
    roofline["fp64_pipe_bound_note"] = (
        "17 FP64 instructions per stored pair at 64 FP64 lanes/clk/SM bound this sweep at ~0.31 ms per 4 M atoms "
        "(frac 0.71); ncu sm__inst_executed_pipe_fp64 in profiles/") if args.precision == 64 and not args.half else None
    whole_step_bytes = 604.0  # SURVEY.md 8d, B/atom-step, full list
    buckets = {k: v[0] for k, v in m["timing"].items()}

    e2e = None
    if not args.no_e2e:
        ke = max(1, min(args.steps, 3))
        sec, h2d, d2h, stages = measure_e2e(args, sim, ke, dist_ctx)
        e2e = {"value": atoms_total * ke * MD_PER_STEP / sec, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "stage_ms_rank0": stages,
               "note": "per bench step: pinned-host upload of x,v,type,id -> init path (wrap/migrate, sort, "
                       "ghosts, Verlet build, force) -> 20 MD steps -> download x,v,type,id + thermo"}

    extra = {}
    sim.ctx.close()
    del sim
    md = lambda mm, st: st * MD_PER_STEP / mm["sec"]
    if n == 1 and not args.no_ab and not args.half and args.precision == 64:
        # A/B of the force kernel's gather path on the headline workload
        os.environ["CBMD_GATHER"] = "0"
        try:
            s1 = build_sim(args, args.cells, args.half, 1, 0, None, local)
            s1.setup()
            m1 = measure_resident(args, s1, args.steps, args.warmup, None, False)
            f_ms, f_n = m1["timing"]["force_kernel"]
            extra["A/B gather=0 (32-byte records, LDG.256 only)"] = {
                "value": s1.N * md(m1, args.steps), "unit": UNIT, "atoms": s1.N,
                "force_kernel_ms": f_ms / max(f_n, 1),
                "force_bucket_ms_per_100_md_steps": m1["timing"]["force"][0] * 100.0 / md_steps}
            s1.ctx.close()
            del s1
        finally:
            del os.environ["CBMD_GATHER"]
    roofline_fp32 = None
    if not args.no_extra and not args.half and args.precision == 64:
        # the FP32 force variant (option precision=32) on the headline workload, every N
        a3 = argparse.Namespace(**vars(args))
        a3.precision = 32
        s3 = build_sim(a3, args.cells, False, n, rank, fresh_uid(dist_ctx, rank), local)
        s3.setup()
        m3 = measure_resident(a3, s3, args.steps, args.warmup, dist_ctx, False)
        roofline_fp32 = roofline_of(m3, False, 32, peak, peak_src,
                                    "k_force_full_f32 (float4 LDG.128, index stream through TEX)")
        roofline_fp32["bytes_convention"] = "SURVEY 8d with 12-byte positions/forces (24 -> 12)"
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "force_traffic.json")))
            if f"full_{args.cells}_f32" in tr:
                roofline_fp32["traffic"] = tr[f"full_{args.cells}_f32"]
                roofline_fp32["traffic_source"] = tr.get("source")
        except Exception:
            pass
        extra["FP32 force variant (precision=32), same workload"] = {
            "value": s3.N * md(m3, args.steps), "unit": UNIT, "atoms": s3.N,
            "force_kernel_ms": roofline_fp32["avg_launch_ms"], "roofline_frac": roofline_fp32["frac"]}
        s3.ctx.close()
        del s3
    if n == 1 and not args.no_extra:
        for half in (False, True):
            a2 = argparse.Namespace(**vars(args))
            a2.half, a2.precision = half, 64
            s2 = build_sim(a2, 63, half, 1, 0, None, local)
            s2.setup()
            k2 = max(args.steps, 5)
            m2 = measure_resident(a2, s2, k2, max(args.warmup, 3), None, False)
            f_ms, f_n = m2["timing"]["force_kernel"]
            fb2 = force_bytes(m2["n_local"], m2["n_ghost"], m2["nn"], half)
            extra[f"configs[1] 1M atoms {'half' if half else 'full'} list"] = {
                "value": s2.N * md(m2, k2), "unit": UNIT,
                "atoms": s2.N, "force_kernel_ms": f_ms / max(f_n, 1),
                "force_kernel_GBps_algorithmic": fb2 / (f_ms / max(f_n, 1) * 1e-3) / 1e9,
                "stored_neighbours_per_atom": m2["nn"],
                "note": "1 M-atom position array (32 MB) fits the 126 MB L2"}
            s2.ctx.close()
        # configs[4]: long cutoff, rc = 5.0 sigma (~526 stored neighbours per atom), 1 M atoms
        a5 = argparse.Namespace(**vars(args))
        a5.half, a5.cutoff, a5.guess, a5.melt, a5.precision = False, 5.0, 600, 40, 64
        s5 = build_sim(a5, 63, False, 1, 0, None, local)
        s5.setup()
        m5 = measure_resident(a5, s5, 2, 1, None, False)
        f_ms, f_n = m5["timing"]["force_kernel"]
        fb5 = force_bytes(m5["n_local"], m5["n_ghost"], m5["nn"], False)
        extra["configs[4] 1M atoms rc=5.0 full list"] = {
            "value": s5.N * md(m5, 2), "unit": UNIT, "atoms": s5.N,
            "force_kernel_ms": f_ms / max(f_n, 1),
            "force_kernel_GBps_algorithmic": fb5 / (f_ms / max(f_n, 1) * 1e-3) / 1e9,
            "force_kernel_frac_of_peak": fb5 / (f_ms / max(f_n, 1) * 1e-3) / 1e9 / peak,
            "stored_neighbours_per_atom": m5["nn"], "ghost_fraction": m5["n_ghost"] / m5["n_local"]}
        s5.ctx.close()
    if n == 1 and not args.no_extra:
        # the product driver itself: the C++ step loop (cabanamd_b200/host, binary cbnMD) on configs[0]
        # (input/in.lj as shipped: 20^3 fcc cells) and on the headline workload — no Python in the loop
        for cells, steps in ((20, 2000), (args.cells, 200)):
            try:
                r = cbnmd_leg(cells, steps)
            except Exception as e:  # noqa: BLE001 - an extra leg must never take the bench line down
                r = {"error": repr(e)[:300]}
            if r:
                extra[f"cbnMD (C++ driver) in.lj, {4 * cells ** 3} atoms, {steps} steps"] = r
    if not args.no_extra:
        # configs[3]: 16.4 M atoms strong scaling on the N GPUs + the drift check vs the oracle
        ss = strong_scaling_leg(args, n, rank, local, dist_ctx)
        if ss:
            extra["configs[3] 16M atoms strong scaling, 1000 steps"] = ss
    if not args.no_checks:
        dc = drift_check(args, n, rank, local, dist_ctx)
        if dc:
            extra["configs[3] drift check"] = dc

    cpu = None
    if rank == 0 and n == 1 and not args.no_cpu:
        cpu = cpu_baseline(args.cpu_cells, 5 * MD_PER_STEP)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["sec"] / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == 64 else "f32",
            "data": "synthetic", "config": workload_config(args, n),
            "clocks": m["clocks"], "e2e": e2e, "gpu_launches": int(m["launches"]),
            "roofline": roofline, "roofline_fp32": roofline_fp32, "cpu_baseline": cpu,
            "parity_vs_oracle": parity,
            "whole_step": {"algorithmic_bytes_per_atom_step": whole_step_bytes,
                           "achieved_GBps_per_gpu": whole_step_bytes * value / n / 1e9,
                           "frac_of_peak": whole_step_bytes * value / n / 1e9 / peak},
            "time_buckets_ms": buckets, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if dist_ctx:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
