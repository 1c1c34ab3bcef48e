"""CPU model of the force kernel's L1 traffic: distinct 128-byte position lines per warp
gather request, for the current layout (each lane walks its own ascending list) and for
row schedules that insert idle slots so that the lanes of a warp stay close together in
the tile's sorted union of lines.  Uses the oracle's sorted atoms + Verlet list of a melted
LJ liquid, which the GPU path reproduces index for index.  Not part of the product."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import oracle_lib as O

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
sim = O.Sim(mass=[2.0]).create_lattice_fcc(cells=(cells,) * 3).setup()
sim.run(steps, 0)
cnt, off, nb = sim.list()
nl = len(off) - 1
print("atoms", nl, "ghosts", len(cnt) - nl, "nn", nb.size / nl)
APL = 4  # atoms per 128-byte line (32-byte records)
rng = np.random.default_rng(0)
cand = rng.permutation(nl // 32)
tiles = [t for t in cand if nb[off[32 * t]:off[32 * t + 32]].max() < nl][:300]
print("interior tiles sampled", len(tiles))

def lists_of(t):
    return [np.sort(nb[off[i]:off[i + 1]]) // APL for i in range(32 * t, 32 * t + 32)]

def current(t):
    lists = lists_of(t)
    rows = max(len(l) for l in lists)
    tot = 0
    for k in range(rows):
        tot += len({int(l[k]) for l in lists if k < len(l)})
    return rows, tot

def paced(t, W, eps):
    lists = lists_of(t)
    union = np.unique(np.concatenate(lists))
    U = len(union)
    ranks = [np.searchsorted(union, l) for l in lists]
    n = [len(l) for l in lists]
    R = int(np.ceil(max(n) * (1 + eps)))
    pos = [0] * 32
    rows = tot = 0
    while any(pos[l] < n[l] for l in range(32)):
        limit = (rows + 1) * U / R + W
        used = set()
        for l in range(32):
            if pos[l] < n[l] and ranks[l][pos[l]] < limit:
                used.add(int(ranks[l][pos[l]]))
                pos[l] += 1
        rows += 1
        tot += len(used)
    return rows, tot

r0 = np.array([current(t) for t in tiles])
print(f"current      : rows/tile {r0[:,0].mean():6.1f}  lines/row {r0[:,1].sum()/r0[:,0].sum():5.2f}  lines/tile {r0[:,1].mean():7.1f}")
for eps in (0.0, 0.1, 0.2):
    for W in (1, 2, 4, 8):
        r = np.array([paced(t, W, eps) for t in tiles])
        print(f"paced W{W:2d} e{eps:.1f}: rows/tile {r[:,0].mean():6.1f}  lines/row {r[:,1].sum()/r[:,0].sum():5.2f}  lines/tile {r[:,1].mean():7.1f}"
              f"   rows x{r[:,0].mean()/r0[:,0].mean():.2f}  L1 x{r[:,1].mean()/r0[:,1].mean():.2f}")
