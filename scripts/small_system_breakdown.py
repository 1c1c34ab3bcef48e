#!/usr/bin/env python
"""Where a step of a SMALL system (in.lj as shipped, 32 000 atoms; 256 000 atoms) spends its time:
host wall clock per MD step of the regular loop (rebuild every 20, thermo every 10) through the Python
harness — stepwise module calls against cbmd_md_steps with and without the CUDA graph, timing buckets
on and off — and through the cbnMD driver.  Run on a GPU box: python scripts/small_system_breakdown.py"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import bench


def harness(cells, timing, batch, graph, steps=2000):
    a = argparse.Namespace(cutoff=2.5, guess=50, precision=64)
    sim = bench.build_sim(a, cells, False, 1, 0, None, 0)
    sim.ctx.set_option("graph_steps", graph)
    sim.setup()
    sim.run(100, 10, batch=batch)
    c = sim.ctx
    c.timing_enable(timing)
    c.sync()
    t0 = time.perf_counter()
    sim.run(steps, 10, batch=batch)
    c.sync()
    us = (time.perf_counter() - t0) / steps * 1e6
    c.close()
    return us


def cbnmd(cells, steps, env):
    exe = os.path.join(ROOT, "cabanamd_b200", "lib", "cbnMD")
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "in.lj"), "w").write(bench.IN_LJ.format(c=cells, steps=steps))
        subprocess.run([exe, "-il", "in.lj", "-o", "md.out", "-e", "md.err"], cwd=td, check=True,
                       capture_output=True, env=dict(os.environ, **env))
        lines = open(os.path.join(td, "md.out")).read().splitlines()
        k = max(i for i, ln in enumerate(lines) if ln.startswith("#Steps/s"))
        perf = [ln for ln in lines if "PERFORMANCE" in ln]
        return float(lines[k + 1].split()[0]), perf[-1] if perf else ""


if __name__ == "__main__":
    for cells in (20, 40):
        n = 4 * cells ** 3
        for timing in (0, 1):
            row = {name: round(harness(cells, timing, b, g), 1)
                   for name, b, g in (("stepwise", False, 1), ("md_steps", True, 0), ("md_steps+graph", True, 1))}
            print(f"harness {n} atoms, us per MD step, timing buckets {'on' if timing else 'off'}: {row}", flush=True)
        for name, env in (("stepwise", {"CBMD_BATCH_STEPS": "0"}), ("md_steps", {"CBMD_GRAPH": "0"}),
                          ("md_steps+graph", {})):
            sps, perf = cbnmd(cells, 2000, env)
            print(f"cbnMD {n} atoms {name}: {1e6 / sps:.1f} us/step = {n * sps:.3e} atom-steps/s | {perf}", flush=True)
