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
configs[1] (1 M atoms, full vs half list) is measured alongside and reported under
"extra".  Data are synthetic: the deck initialiser's lattice + hashed-RNG velocities,
melted for 200 untimed MD steps so the timed state is the LJ liquid.

The CPU oracle is only used here for the `cpu_baseline` leg and `--impl reference`.
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
    ap.add_argument("--melt", type=int, default=200, help="untimed MD steps before warm-up")
    ap.add_argument("--thermo", type=int, default=10, help="thermo every n MD steps (0 = never)")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[1] 1 M-atom legs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ab", action="store_true", help="skip the A/B legs (gather=0, nb_group=8)")
    ap.add_argument("--cpu-cells", type=int, default=40, help="cpu_baseline sample: cells per dim")
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


def build_sim(args, cells_per_gpu, half, nranks, rank, uid, device, temp=1.4, seed=87287):
    from cabanamd_b200.capi import dims_create
    from cabanamd_b200.harness import Simulation, create_velocities

    grid = dims_create(nranks)
    cells_global = tuple(cells_per_gpu * g for g in grid)
    a = (4.0 / 0.8442) ** (1.0 / 3.0)
    sim = Simulation(device=device, mass=(2.0,), cut=args.cutoff, skin=0.3, half=half,
                     exchange_rate=MD_PER_STEP, max_neigh_guess=args.guess, nranks=nranks,
                     rank=rank, uid=uid)
    sim.set_box([0.0] * 3, [a * c for c in cells_global])
    x = local_lattice(sim, cells_global, a)
    t = np.zeros(len(x), dtype=np.int32)
    n_before = sim.ctx.scan_sum_int(len(x)) - len(x) if nranks > 1 else 0
    ids = np.arange(1, len(x) + 1, dtype=np.int32) + n_before
    v = create_velocities(sim, x, t, temp, seed)
    sim.set_atoms(x, v, t, ids)
    # rescale to the target temperature (inputFile_impl.h:851-865)
    T = sim.temperature()
    sim.ctx.set_velocities(v * np.sqrt(temp / T))
    expect = 4 * cells_global[0] * cells_global[1] * cells_global[2]
    assert sim.N == expect, (sim.N, expect)
    return sim


def timed_steps(sim, k, thermo, dist_ctx):
    """K bench steps bracketed by barrier + synchronize, CUDA events on the context
    stream; returns seconds (max over ranks)."""
    import torch

    stream = torch.cuda.ExternalStream(sim.ctx.stream())
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier(dist_ctx)
    sim.ctx.sync()
    torch.cuda.synchronize()
    e0.record(stream)
    sim.run(k * MD_PER_STEP, thermo)
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


def max_over_ranks(val, dist_ctx):
    if not dist_ctx:
        return val
    import torch
    import torch.distributed as dist

    t = torch.tensor([val], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(val, dist_ctx):
    if not dist_ctx:
        return val
    import torch
    import torch.distributed as dist

    t = torch.tensor([val], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def force_bytes(n_local, n_ghost, nn, half):
    """Algorithmic bytes of one LJ force launch (SURVEY.md 8d / DESIGN.md): int32
    indices, FP64 positions/forces, every position read once."""
    if half:
        return n_local * (4.0 * nn + 8.0) + (n_local + n_ghost) * (28.0 + 48.0)
    return n_local * (4.0 * nn + 24.0 + 8.0) + (n_local + n_ghost) * (24.0 + 4.0)


def measure_resident(args, sim, steps, warmup, dist_ctx, sample_clocks):
    sim.run(args.melt, 0)
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
    sec = timed_steps(sim, steps, args.thermo, dist_ctx)
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count() - l0
    tm = ctx.timing()
    ctx.timing_enable(False)
    nl, ng = ctx.counts()
    tot, mx = ctx.neigh_sizes()
    return dict(sec=sec, launches=launches, timing=tm, n_local=nl, n_ghost=ng,
                nn=tot / max(nl, 1), max_neigh=mx, clocks=clocks)


def measure_e2e(args, sim, steps, dist_ctx):
    """Same metric through the public C ABI with HOST buffers: every bench step uploads
    the atoms (x, v, type, id) from pinned host memory, runs the init path + one list
    period, and downloads x, v and the thermo scalars."""
    import torch

    ctx = sim.ctx
    g = ctx.get_atoms(fields="xvti")
    nl = g["n_local"]

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t

    hx, hv = pinned(g["x"][:nl]), pinned(g["v"][:nl])
    ht, hi = pinned(g["type"][:nl]), pinned(g["id"][:nl])
    thermo = args.thermo
    stream = torch.cuda.ExternalStream(ctx.stream())
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    from cabanamd_b200.capi import _dp, _ip

    def one_step():
        ctx.set_atoms(hx.numpy(), hv.numpy(), None, ht.numpy(), hi.numpy())
        sim.setup()
        sim.run(MD_PER_STEP, thermo)
        ctx._ck(ctx.L.cbmd_get_atoms(ctx.h, 0, nl, _dp(hx.numpy()), _dp(hv.numpy()), None, None,
                                     None, None))

    one_step()  # warm-up
    barrier(dist_ctx)
    ctx.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        one_step()
    ctx.sync()
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    barrier(dist_ctx)
    sec = max(e0.elapsed_time(e1) * 1e-3, wall)
    sec = max_over_ranks(sec, dist_ctx)
    h2d = nl * (24 + 24 + 4 + 4)
    n_thermo = (MD_PER_STEP // thermo) if thermo else 0
    d2h = nl * 48 + n_thermo * 3 * 8
    return sec, sum_over_ranks(h2d, dist_ctx), sum_over_ranks(d2h, dist_ctx)


def cpu_baseline(cells, md_steps, threads=None, half=False):
    """The oracle (a port of the reference algorithm: the real Kokkos/Cabana build is not
    available in this image) on the host cores; bounded sample."""
    import oracle_lib as O

    L = O.lib()
    if threads:
        L.orc_set_threads(threads)
    nthr = L.orc_max_threads()
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O

    L = O.lib()
    nthr = L.orc_max_threads()
    cells = args.cpu_cells
    s = O.Sim(mass=[2.0], half=args.half).create_lattice_fcc(cells=(cells,) * 3).setup()
    for _ in range(args.warmup):
        s.run(MD_PER_STEP, args.thermo)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.run(MD_PER_STEP, args.thermo)
    dt = time.perf_counter() - t0
    val = s.natoms * args.steps * MD_PER_STEP / dt
    sample = (f"each step = {MD_PER_STEP} MD steps of a {s.natoms}-atom sample (fcc {cells}^3) of "
              f"the workload, {nthr} OpenMP threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthr, "kind": "port",
                         "sample": sample},
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
        "decomposition": {1: "1x1x1", 2: "2x1x1", 4: "2x2x1", 8: "2x2x2"}.get(n, str(n)),
        "l2_policy": "working set (x,v,f + neighbour table, >1 GB at 4 M atoms) exceeds the 126 MB L2; no flush",
    }


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
    uid = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist_ctx = True
        import cabanamd_b200 as cb

        box = [cb.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    n = world
    assert n == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    sim = build_sim(args, args.cells, args.half, n, rank, uid, local)
    sim.setup()
    m = measure_resident(args, sim, args.steps, args.warmup, dist_ctx, sample_clocks=True)
    md_steps = args.steps * MD_PER_STEP
    atoms_total = sim.N
    value = atoms_total * md_steps / m["sec"]

    # roofline of the dominant kernel (LJ force), measured live with CUDA events
    fk_ms, fk_n = m["timing"]["force_kernel"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"
    fb = force_bytes(m["n_local"], m["n_ghost"], m["nn"], args.half)
    achieved = fb / (fk_ms / max(fk_n, 1) * 1e-3) / 1e9 if fk_ms > 0 else 0.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "force_traffic.json")))
        key = f"{'half' if args.half else 'full'}_{args.cells}"
        traffic = tr.get(key)
    except Exception:
        pass
    gather = int(os.environ.get("CBMD_GATHER", "1"))
    kname = "k_force_half" if args.half else ("k_force_full<1,...> (xy LDG.128 + z TEX)" if gather == 1
                                              else "k_force_full<0,...> (32-byte records)")
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": traffic,
                "bytes_per_launch": fb, "avg_launch_ms": fk_ms / max(fk_n, 1), "launches": fk_n,
                "stored_neighbours_per_atom": m["nn"], "ghost_fraction": m["n_ghost"] / m["n_local"],
                "step_share": fk_ms * 1e-3 / m["sec"]}
    whole_step_bytes = 604.0  # SURVEY.md 8d, B/atom-step, full list
    buckets = {k: v[0] for k, v in m["timing"].items()}

    e2e = None
    if not args.no_e2e:
        sec, h2d, d2h = measure_e2e(args, sim, max(1, min(args.steps, 3)), dist_ctx)
        ke = max(1, min(args.steps, 3))
        e2e = {"value": atoms_total * ke * MD_PER_STEP / sec, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "note": "per bench step: pinned-host upload of x,v,type,id -> init path (wrap, sort, "
                       "ghosts, Verlet build, force) -> 20 MD steps -> download x,v + thermo"}

    extra = {}
    sim.ctx.close()
    del sim
    if n == 1 and not args.no_ab:
        # A/B of the force kernel's gather path / sweep shape on the headline workload
        # (DESIGN.md 3.1): 32-byte records by LDG.256 (no texture path), and 8 lanes per atom
        for label, env in (("A/B gather=0 (32-byte records, LDG.256 only)", {"CBMD_GATHER": "0"}),):
            os.environ.update(env)
            try:
                s1 = build_sim(args, args.cells, args.half, 1, 0, None, local)
                s1.setup()
                m1 = measure_resident(args, s1, args.steps, args.warmup, None, False)
                f_ms, f_n = m1["timing"]["force_kernel"]
                extra[label] = {
                    "value": s1.N * md_steps / m1["sec"], "unit": UNIT, "atoms": s1.N,
                    "force_kernel_ms": f_ms / max(f_n, 1),
                    "force_bucket_ms_per_100_md_steps": m1["timing"]["force"][0] * 100.0 / md_steps}
                s1.ctx.close()
                del s1
            finally:
                for k in env:
                    del os.environ[k]
    if n == 1 and not args.no_extra:
        for half in (False, True):
            a2 = argparse.Namespace(**vars(args))
            a2.half = half
            s2 = build_sim(a2, 63, half, 1, 0, None, local)
            s2.setup()
            m2 = measure_resident(a2, s2, max(args.steps, 5), max(args.warmup, 3), None, False)
            f_ms, f_n = m2["timing"]["force_kernel"]
            fb2 = force_bytes(m2["n_local"], m2["n_ghost"], m2["nn"], half)
            extra[f"configs[1] 1M atoms {'half' if half else 'full'} list"] = {
                "value": s2.N * max(args.steps, 5) * MD_PER_STEP / m2["sec"], "unit": UNIT,
                "atoms": s2.N, "force_kernel_ms": f_ms / max(f_n, 1),
                "force_kernel_GBps_algorithmic": fb2 / (f_ms / max(f_n, 1) * 1e-3) / 1e9,
                "stored_neighbours_per_atom": m2["nn"],
                "note": "1 M-atom position array (32 MB) fits the 126 MB L2"}
            s2.ctx.close()
        # configs[4]: long cutoff, rc = 5.0 sigma (~526 stored neighbours per atom), 1 M atoms
        a5 = argparse.Namespace(**vars(args))
        a5.half, a5.cutoff, a5.guess, a5.melt = False, 5.0, 600, 40
        s5 = build_sim(a5, 63, False, 1, 0, None, local)
        s5.setup()
        m5 = measure_resident(a5, s5, 2, 1, None, False)
        f_ms, f_n = m5["timing"]["force_kernel"]
        fb5 = force_bytes(m5["n_local"], m5["n_ghost"], m5["nn"], False)
        extra["configs[4] 1M atoms rc=5.0 full list"] = {
            "value": s5.N * 2 * MD_PER_STEP / m5["sec"], "unit": UNIT, "atoms": s5.N,
            "force_kernel_ms": f_ms / max(f_n, 1),
            "force_kernel_GBps_algorithmic": fb5 / (f_ms / max(f_n, 1) * 1e-3) / 1e9,
            "stored_neighbours_per_atom": m5["nn"], "ghost_fraction": m5["n_ghost"] / m5["n_local"]}
        s5.ctx.close()

    cpu = None
    if rank == 0 and n == 1 and not args.no_cpu:
        cpu = cpu_baseline(args.cpu_cells, 2 * MD_PER_STEP)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["sec"] / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, n),
            "clocks": m["clocks"], "e2e": e2e, "gpu_launches": int(m["launches"]),
            "roofline": roofline, "cpu_baseline": cpu,
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
