#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (oracle/oracle.hpp).

The reference itself cannot be built in this image (needs Kokkos + Cabana + MPI), so these
fixtures freeze the ORACLE's outputs on fixed seeded inputs: they pin the oracle against
regressions and give the GPU tests committed vectors to compare with.  Inputs follow the
reference's own unit tests where they exist (tstNeighbor.hpp:289-304 configuration).
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]
import oracle_lib as O  # noqa: E402


def neighbor_fixture():
    rng = np.random.default_rng(342343901)
    n, n_ghost, rc = 1000, 200, 2.32
    lo, hi = -5.3 * rc, 4.7 * rc
    x = rng.uniform(lo, hi, (n, 3))
    out = dict(x=x, n_local=np.int32(n - n_ghost), rc=np.float64(rc), lo=np.float64(lo), hi=np.float64(hi))
    for half in (False, True):
        nl = O.NeighList()
        nl.brute(x, n - n_ghost, rc, half)
        c, o, nb = nl.arrays()
        rows = np.concatenate([np.sort(nb[o[i]:o[i + 1]]) for i in range(n - n_ghost)]) if len(nb) else nb
        tag = "half" if half else "full"
        out[f"counts_{tag}"] = c.astype(np.int32)
        out[f"rows_{tag}"] = rows.astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "neighbor_tstneighbor.npz"), **out)


def md_fixture():
    out = {}
    for half in (False, True):
        s = O.Sim(mass=[2.0], half=half).create_lattice_fcc(cells=(8, 8, 8)).setup()
        d0 = s.get()
        n = d0["n_local"]
        tag = "half" if half else "full"
        if not half:
            out["x0"], out["v0"], out["id0"] = d0["x"][:n], d0["v"][:n], d0["id"][:n]
        s.record_thermo()
        s.run(100, 10)
        d = s.get()
        n = d["n_local"]
        order = np.argsort(d["id"][:n])
        out[f"thermo_{tag}"] = np.array(s.thermo())
        out[f"x100_{tag}"] = d["x"][:n][order]
        out[f"f100_{tag}"] = d["f"][:n][order]
        out[f"pe_corrected_{tag}"] = np.float64(s.potential(True))
    np.savez_compressed(os.path.join(HERE, "md_fcc8_100steps.npz"), **out)


if __name__ == "__main__":
    neighbor_fixture()
    md_fixture()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
