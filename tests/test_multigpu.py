"""N>1 coverage.  GPU: torchrun tests/mp_parity.py on 2 (and 4/8 when present) GPUs over NCCL; on ANY
GPU box (one device is enough) the same comparison with the ranks as host threads of one process over
the in-process hub (cbmd_hub_create) — migration, 6-phase ghost build, one-stage and staged halo refresh,
reverse force fold and the scalar collectives at 2/4/8 ranks against the oracle's virtual ranks.
CPU: world_size-2 gloo test of the host-side decomposition logic (lattice split, id scan,
momentum reduction) against the oracle's virtual ranks."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("n", [2, 4, 8])
def test_multigpu_matches_oracle(n, half):
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(29511 + n + (50 if half else 0)),
           os.path.join(ROOT, "tests", "mp_parity.py")] + (["--half"] if half else [])
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0 and "MP_PARITY_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


def _hub_parity(n, **kw):
    sys.path[:0] = [p for p in (ROOT, os.path.join(ROOT, "tests")) if p not in sys.path]
    from mp_parity import run_parity_threads

    ok, worst = run_parity_threads(n, 0, **kw)
    assert ok, worst
    return worst


@pytest.mark.gpu
@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("n", [2, 4, 8])
def test_ranks_on_one_gpu_match_oracle(n, half):
    """2/4/8 ranks ({2,1,1}, {2,2,1}, {2,2,2}) sharing cuda:0: thermo, per-id x/v/f and per-rank ghost
    sets against the oracle with the same number of virtual ranks, thermo also against one rank."""
    worst = _hub_parity(n, half=half)
    assert f"thermo_vs_{n}rank" in worst and f"f_vs_{n}rank" in worst and "ghost_x_rank0" in worst


@pytest.mark.gpu
def test_ranks_on_one_gpu_three_in_a_row_and_fp32():
    """Three ranks in x (distinct +x and -x peers, unlike the 2-rank torus) and the FP32 sweep at 4."""
    _hub_parity(3, half=False)
    _hub_parity(3, half=True)
    _hub_parity(4, half=False, precision=32)


@pytest.mark.gpu
def test_ranks_on_one_gpu_staged_halo_and_no_overlap():
    """The 3-stage refresh (option halo_stages 3) and the non-overlapped step give the same answers."""
    os.environ["CBMD_HALO_STAGES"] = "3"
    try:
        _hub_parity(4, half=False)
        _hub_parity(8, half=True)
    finally:
        del os.environ["CBMD_HALO_STAGES"]
    os.environ["CBMD_OVERLAP"] = "0"
    try:
        _hub_parity(4, half=False)
    finally:
        del os.environ["CBMD_OVERLAP"]
    # one-stage refresh beside the integrator (boundary tiles first) instead of beside the interior force tiles
    os.environ["CBMD_EARLY"] = "1"
    try:
        _hub_parity(4, half=False)
        _hub_parity(2, half=True)
    finally:
        del os.environ["CBMD_EARLY"]


@pytest.mark.gpu
def test_hub_peer_that_never_arrives_fails_instead_of_hanging():
    import cabanamd_b200 as cb

    hub = cb.Hub(2, timeout=1.0)
    c = cb.Context(0)
    c.comm_init(2, 0, hub)
    with pytest.raises(cb.CbmdError, match="hub"):
        c.reduce_sum_int(1)
    c.close()
    hub.close()


# ---------------------------------------------------------------- CPU, gloo, world_size 2
WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = sys.argv[1]
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from cabanamd_b200.capi import make_domain, dims_create
from cabanamd_b200.harness import velocity_geom_uniform
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cells_per = 6
grid = dims_create(world)
cells = tuple(cells_per * g for g in grid)
a = (4.0 / 0.8442) ** (1.0 / 3.0)
dom = make_domain([0.0] * 3, [a * c for c in cells], world, rank, 20.0)

class S:  # the two attributes bench.local_lattice reads
    pass
s = S(); s.dom = dom
from bench import local_lattice
x = local_lattice(s, cells, a)
# MPI_Scan of the local counts -> id offsets (inputFile_impl.h:778-784)
counts = [None] * world
dist.all_gather_object(counts, len(x))
offset = sum(counts[:rank])
ids = np.arange(1, len(x) + 1) + offset
m = np.full(len(x), 2.0)
v = (velocity_geom_uniform(x, 87287) - 0.5) / np.sqrt(m)[:, None]
tot = torch.tensor([m.sum(), *(m[:, None] * v).sum(0)], dtype=torch.float64)
dist.all_reduce(tot)
v = v - (tot[1:] / tot[0]).numpy()
out = [None] * world
dist.all_gather_object(out, dict(x=x, v=v, ids=ids, dom={k: np.asarray(val) for k, val in dom.items()}))
if rank == 0:
    np.save(sys.argv[2], np.array(out, dtype=object), allow_pickle=True)
dist.barrier()
dist.destroy_process_group()
"""


def test_gloo_two_rank_decomposition_matches_oracle(tmp_path):
    import oracle_lib as O

    out = tmp_path / "ranks.npy"
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(w), ROOT, str(out)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    ranks = np.load(out, allow_pickle=True)

    ref = O.Sim(mass=[2.0]).create_lattice_fcc(cells=(12, 6, 6), nranks=2, temp=1.4)
    all_ids = np.concatenate([r["ids"] for r in ranks])
    assert np.array_equal(np.sort(all_ids), np.arange(1, 4 * 12 * 6 * 6 + 1))
    vs, rvs = [], []
    for rk in range(2):
        d, dom = ref.get(rk), ref.domain(rk)
        n = d["n_local"]
        for k in ("llo", "lhi", "ghost_lo", "ghost_hi", "grid", "pos"):
            assert np.array_equal(ranks[rk]["dom"][k], dom[k]), k
        # same atoms, same order, same ids, bit-identical positions
        assert np.array_equal(ranks[rk]["x"], d["x"][:n])
        assert np.array_equal(ranks[rk]["ids"], d["id"][:n])
        vs.append(ranks[rk]["v"]); rvs.append(d["v"][:n])
    # velocities before the temperature rescale differ from the oracle's final ones by one factor
    v, rv = np.concatenate(vs), np.concatenate(rvs)
    scale = (rv * v).sum() / (v * v).sum()
    assert np.allclose(rv, scale * v, rtol=1e-12, atol=1e-14)
    assert abs(v.sum(0)).max() < 1e-9
