"""Where the end-to-end bench step spends its time (4 M atoms, one GPU): host->device
upload, init path (setup), 20 MD steps, device->host download."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from cabanamd_b200.capi import _dp

args = bench.parse()
sim = bench.build_sim(args, args.cells, False, 1, 0, None, 0)
sim.setup(); sim.run(100, 0)
ctx = sim.ctx
g = ctx.get_atoms(fields="xvti"); nl = g["n_local"]
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True); t.numpy()[...] = a; return t
hx, hv, ht, hi = pinned(g["x"][:nl]), pinned(g["v"][:nl]), pinned(g["type"][:nl]), pinned(g["id"][:nl])
T = {"upload": 0.0, "setup": 0.0, "run20": 0.0, "download": 0.0}
def tick(name, fn):
    ctx.sync(); t0 = time.perf_counter(); fn(); ctx.sync(); T[name] += time.perf_counter() - t0
for it in range(4):
    if it == 1:
        for k in T: T[k] = 0.0
    tick("upload", lambda: ctx.set_atoms(hx.numpy(), hv.numpy(), None, ht.numpy(), hi.numpy()))
    tick("setup", sim.setup)
    tick("run20", lambda: sim.run(20, 10))
    tick("download", lambda: ctx._ck(ctx.L.cbmd_get_atoms(ctx.h, 0, nl, _dp(hx.numpy()), _dp(hv.numpy()), None, None, None, None)))
print({k: round(v / 3 * 1e3, 2) for k, v in T.items()}, "ms per e2e step; H2D MB", nl * 56 / 1e6, "D2H MB", nl * 48 / 1e6)
