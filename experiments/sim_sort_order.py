"""How the atom ORDER changes the distinct 128-byte lines per warp gather of the force
kernel (model only; see sim_row_schedule.py).  Periodic LJ liquid from the oracle; lists
rebuilt here with a k-d tree for every candidate ordering."""
import sys, os
import numpy as np
from scipy.spatial import cKDTree
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import oracle_lib as O

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 24
sim = O.Sim(mass=[2.0]).create_lattice_fcc(cells=(cells,) * 3).setup()
sim.run(100, 0)
d = sim.get()
n = d["n_local"]
x = d["x"][:n].copy()
L = cells * (4 / 0.8442) ** (1 / 3)
x %= L
RN, APL = 2.8, 4
tree = cKDTree(x, boxsize=L)
pairs = tree.query_pairs(RN, output_type="ndarray")
print("atoms", n, "nn", 2 * len(pairs) / n)

def morton(ix, iy, iz):
    def spread(v):
        v = v.astype(np.uint64) & 0x3FF
        v = (v | (v << 16)) & 0x30000FF
        v = (v | (v << 8)) & 0x300F00F
        v = (v | (v << 4)) & 0x30C30C3
        v = (v | (v << 2)) & 0x9249249
        return v
    return spread(ix) | (spread(iy) << 1) | (spread(iz) << 2)

def order_key(kind):
    nc = int(L // RN); cs = L / nc
    c = np.minimum((x / cs).astype(int), nc - 1)
    cell = (c[:, 0] * nc + c[:, 1]) * nc + c[:, 2]
    if kind == "cell":
        return np.lexsort((np.arange(n), cell))
    if kind == "cell+sub2":
        s = np.minimum(((x - c * cs) / (cs / 2)).astype(int), 1)
        return np.lexsort((np.arange(n), (s[:, 0] * 2 + s[:, 1]) * 2 + s[:, 2], cell))
    if kind == "cell+sub3":
        s = np.minimum(((x - c * cs) / (cs / 3)).astype(int), 2)
        return np.lexsort((np.arange(n), (s[:, 0] * 3 + s[:, 1]) * 3 + s[:, 2], cell))
    if kind == "cell+zsort":
        return np.lexsort((x[:, 2], cell))
    if kind.startswith("grid"):
        h = float(kind[4:]); m = int(L // h); hs = L / m
        g = np.minimum((x / hs).astype(int), m - 1)
        return np.lexsort((np.arange(n), (g[:, 0] * m + g[:, 1]) * m + g[:, 2]))
    if kind.startswith("morton"):
        h = float(kind[6:]); m = int(L // h); hs = L / m
        g = np.minimum((x / hs).astype(int), m - 1)
        return np.lexsort((np.arange(n), morton(g[:, 0], g[:, 1], g[:, 2])))
    if kind.startswith("colz"):  # thin (x,y) columns, fully z-sorted inside
        h = float(kind[4:]); m = int(L // h); hs = L / m
        g = np.minimum((x[:, :2] / hs).astype(int), m - 1)
        return np.lexsort((x[:, 2], g[:, 0] * m + g[:, 1]))
    raise ValueError(kind)

rng = np.random.default_rng(0)
def measure(kind, ntiles=300):
    perm = order_key(kind)              # new[i] = old[perm[i]]
    inv = np.empty(n, int); inv[perm] = np.arange(n)
    a, b = inv[pairs[:, 0]], inv[pairs[:, 1]]
    src = np.concatenate([a, b]); dst = np.concatenate([b, a])
    o = np.lexsort((dst, src)); src, dst = src[o], dst[o]
    off = np.searchsorted(src, np.arange(n + 1))
    tiles = rng.choice(n // 32, size=ntiles, replace=False)
    rows = lines = req_lanes = 0
    for t in tiles:
        ls = [dst[off[i]:off[i + 1]] // APL for i in range(32 * t, 32 * t + 32)]
        R = max(len(l) for l in ls)
        rows += R
        for k in range(R):
            lines += len({int(l[k]) for l in ls if k < len(l)})
    return rows / ntiles, lines / rows, lines / ntiles

for kind in ("cell", "cell+sub2", "cell+sub3", "cell+zsort", "grid1.4", "grid0.93", "morton1.4", "morton0.7",
             "colz1.4", "colz0.93"):
    r, lpr, lpt = measure(kind)
    print(f"{kind:12s} rows/tile {r:6.1f}  lines/row {lpr:5.2f}  lines/tile {lpt:7.1f}")

print("\nlanes sharing an atom (G lanes per atom, 32/G atoms per warp): L1 lines and FP64 warp-rows PER ATOM")
print(f"{'thread/atom cell':22s} lines/atom {1602.1/32:6.1f}  warp-rows/atom {80.2/32:5.2f}")
def measure_group(kind, G, ntiles=300):
    perm = order_key(kind)
    inv = np.empty(n, int); inv[perm] = np.arange(n)
    a, b = inv[pairs[:, 0]], inv[pairs[:, 1]]
    src = np.concatenate([a, b]); dst = np.concatenate([b, a])
    o = np.lexsort((dst, src)); src, dst = src[o], dst[o]
    off = np.searchsorted(src, np.arange(n + 1))
    A = 32 // G
    groups = rng.choice(n // A, size=ntiles, replace=False)
    rows = lines = 0
    for gidx in groups:
        ls = [dst[off[i]:off[i + 1]] // APL for i in range(A * gidx, A * gidx + A)]
        R = max((len(l) + G - 1) // G for l in ls)
        rows += R
        for k in range(R):
            s = set()
            for l in ls:
                s.update(l[k * G:(k + 1) * G].tolist())
            lines += len(s)
    return lines / (ntiles * A), rows / (ntiles * A)
for kind in ("cell", "cell+sub2", "grid1.4", "morton1.4", "morton0.7", "colz1.4", "colz0.93", "colz0.7"):
    for G in (4, 8, 16, 32):
        l, r = measure_group(kind, G)
        print(f"{kind:12s} G={G:2d}    lines/atom {l:6.1f}  warp-rows/atom {r:5.2f}")

print("\nquarter-warp model (LDG.256: one wavefront per distinct 128-byte line among each aligned group of 8 lanes)")
def measure_quarter(kind, natoms=4000):
    perm = order_key(kind)
    inv = np.empty(n, int); inv[perm] = np.arange(n)
    a, b = inv[pairs[:, 0]], inv[pairs[:, 1]]
    src = np.concatenate([a, b]); dst = np.concatenate([b, a])
    o = np.lexsort((dst, src)); src, dst = src[o], dst[o]
    off = np.searchsorted(src, np.arange(n + 1))
    # (1) one lane per atom: 8 consecutive atoms, k-th entries
    t1 = r1 = 0
    for t in rng.choice(n // 8, size=natoms // 8, replace=False):
        ls = [dst[off[i]:off[i + 1]] // APL for i in range(8 * t, 8 * t + 8)]
        R = max(len(l) for l in ls)
        for k in range(R):
            t1 += len({int(l[k]) for l in ls if k < len(l)})
        r1 += sum(len(l) for l in ls)
    # (2) 8 lanes per atom: 8 consecutive entries of one atom
    t8 = r8 = 0
    for i in rng.choice(n, size=natoms, replace=False):
        l = dst[off[i]:off[i + 1]] // APL
        for k in range(0, len(l), 8):
            t8 += len(set(l[k:k + 8].tolist()))
        r8 += len(l)
    return t1 / r1, t8 / r8
for kind in ("cell", "cell+sub2", "cell+sub3", "cell+zsort", "grid1.4", "grid0.93", "morton1.4", "morton0.7", "colz1.4", "colz0.93", "colz0.7"):
    w1, w8 = measure_quarter(kind)
    print(f"{kind:12s} wavefronts per pair: lane/atom {w1:5.3f} (x32 = {32*w1:5.1f}/request)   8 lanes/atom {w8:5.3f} (x32 = {32*w8:5.1f}/request)")
