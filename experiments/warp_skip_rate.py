#!/usr/bin/env python
"""Round-1 verdict leg (b): "decide 'outside the cutoff' from a cheap FP32 test and keep FP64 only for the
in-range pairs; the saving only materialises when whole warps skip, so report the measured warp-level
skip rate".  This counts it (set logic on the CPU oracle's state, no timing): a melted LJ liquid, Verlet
rows as the product stores them (32-atom tiles, rows ascending in atom index, four entries per int4),
and for each (tile, int4 slot) whether ALL 32 lanes x 4 entries — or all 32 lanes of one entry — lie
beyond the force cutoff, 0 / 10 / 19 steps after the rebuild.  Second table: the best case for the idea,
rows sorted by distance at build time (which gives up the index order the gathers like, section 3.2).

    python experiments/warp_skip_rate.py [cells=20]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle_lib as O


def rows_of(sim):
    c, o, n = sim.list()
    d = sim.get()
    nl = d["n_local"]
    return d["x"], nl, c[:nl], o, n


def skip_rates(x, nl, counts, offs, neigh, rc2, order_by_distance_of=None):
    rows = int(counts.max())
    rows4 = (rows + 3) // 4
    tiles = (nl + 31) // 32
    # table[i, k] = r^2 of entry k (inf for padding)
    r2 = np.full((tiles * 32, rows4 * 4), np.inf)
    for i in range(nl):
        j = neigh[offs[i]:offs[i + 1]]
        if order_by_distance_of is not None:
            d0 = ((order_by_distance_of[j] - order_by_distance_of[i]) ** 2).sum(1)
            j = j[np.argsort(d0, kind="stable")]
        r2[i, :len(j)] = ((x[j] - x[i]) ** 2).sum(1)
    listed = np.isfinite(r2)
    outside = r2 > rc2                     # padding counts as outside (the kernel guards it anyway)
    t = outside.reshape(tiles, 32, rows4 * 4)
    lt = listed.reshape(tiles, 32, rows4 * 4)
    warp_entry_all_out = t.all(axis=1)     # [tile, entry]: every lane's entry k is outside
    warp_entry_any_listed = lt.any(axis=1)
    slot_all_out = warp_entry_all_out.reshape(tiles, rows4, 4).all(axis=2)
    slot_any_listed = warp_entry_any_listed.reshape(tiles, rows4, 4).any(axis=2)
    pairs = listed.sum()
    return dict(
        stored_per_atom=pairs / nl,
        inside_per_atom=(listed & ~outside).sum() / nl,
        lane_skip=float((listed & outside).sum() / pairs),
        warp_entry_skip=float((warp_entry_all_out & warp_entry_any_listed).sum() / warp_entry_any_listed.sum()),
        warp_int4_skip=float((slot_all_out & slot_any_listed).sum() / slot_any_listed.sum()),
    )


def main():
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    O.lib().orc_set_threads(os.cpu_count() or 1)
    sim = O.Sim(mass=[2.0]).create_lattice_fcc(cells=(cells,) * 3).setup()
    sim.run(200, 0)   # melt; the last rebuild is at step 200
    rc2 = 2.5 ** 2
    x0 = sim.get()["x"].copy()
    print(f"LJ liquid, {4 * cells ** 3} atoms, rc 2.5, skin 0.3, rebuild every 20 steps; rows in index order (product layout)")
    print("steps after rebuild | stored/atom inside/atom | lanes outside | warp skips one entry | warp skips a whole int4 (4 entries)")
    out = []
    for after in (0, 10, 19):
        sim.run(after - (out[-1][0] if out else 0), 0)
        x, nl, c, o, n = rows_of(sim)
        r = skip_rates(x, nl, c, o, n, rc2)
        out.append((after, r))
        print(f"{after:>19} | {r['stored_per_atom']:.1f} {r['inside_per_atom']:.1f} | {r['lane_skip']:.3f} | "
              f"{r['warp_entry_skip']:.5f} | {r['warp_int4_skip']:.5f}")
        rs = skip_rates(x, nl, c, o, n, rc2, order_by_distance_of=x0)
        print(f"{'rows by distance':>19} | {'':>9} | {rs['lane_skip']:.3f} | {rs['warp_entry_skip']:.5f} | {rs['warp_int4_skip']:.5f}")


if __name__ == "__main__":
    main()
