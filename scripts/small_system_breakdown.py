#!/usr/bin/env python
"""Where a step of a SMALL system (in.lj as shipped, 32 000 atoms) spends its time: host wall clock per
plain step, per rebuild step and per thermo step through the Python harness, and the cbnMD driver with
and without thermo output.  Run on a GPU box: python scripts/small_system_breakdown.py"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import argparse

import numpy as np

import bench


def harness(cells, timing):
    a = argparse.Namespace(cutoff=2.5, guess=50, precision=64)
    sim = bench.build_sim(a, cells, False, 1, 0, None, 0)
    sim.setup()
    sim.run(100, 0)
    c = sim.ctx
    c.timing_enable(timing) if hasattr(c, "timing_enable") else None
    c.sync()
    out = {}
    # plain steps
    sim.exchange_rate = 10 ** 9
    t0 = time.perf_counter(); sim.run(2000, 0); c.sync(); out["plain_us"] = (time.perf_counter() - t0) / 2000 * 1e6
    # thermo every step
    t0 = time.perf_counter(); sim.run(300, 1); c.sync(); out["thermo_step_us"] = (time.perf_counter() - t0) / 300 * 1e6
    # rebuild every step
    sim.exchange_rate = 1
    t0 = time.perf_counter(); sim.run(300, 0); c.sync(); out["rebuild_step_us"] = (time.perf_counter() - t0) / 300 * 1e6
    c.close()
    return out


def cbnmd(cells, steps, thermo):
    exe = os.path.join(ROOT, "cabanamd_b200", "lib", "cbnMD")
    with tempfile.TemporaryDirectory() as td:
        deck = bench.IN_LJ.format(c=cells, steps=steps).replace("thermo          10", f"thermo          {thermo}")
        open(os.path.join(td, "in.lj"), "w").write(deck)
        subprocess.run([exe, "-il", "in.lj", "-o", "md.out", "-e", "md.err"], cwd=td, check=True, capture_output=True)
        lines = open(os.path.join(td, "md.out")).read().splitlines()
        k = max(i for i, ln in enumerate(lines) if ln.startswith("#Steps/s"))
        perf = [ln for ln in lines if "PERFORMANCE" in ln]
        return float(lines[k + 1].split()[0]), perf[-1] if perf else ""


if __name__ == "__main__":
    for cells in (20, 40):
        for timing in (0, 1):
            print(f"harness {4 * cells ** 3} atoms, timing buckets {'on' if timing else 'off'}:",
                  {k: round(v, 1) for k, v in harness(cells, timing).items()}, flush=True)
        for thermo in (10, 1000000):
            sps, perf = cbnmd(cells, 2000, thermo)
            print(f"cbnMD {4 * cells ** 3} atoms thermo {thermo}: {1e6 / sps:.1f} us/step | {perf}", flush=True)
