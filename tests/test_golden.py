"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from
the CPU oracle; the reference itself cannot be built here).  CPU: the oracle still
reproduces them.  GPU: the CUDA path reproduces them through the C ABI."""
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def split_rows(counts, rows, n_local):
    off = np.concatenate([[0], np.cumsum(counts[:n_local])])
    return [rows[off[i]:off[i + 1]] for i in range(n_local)]


@pytest.mark.parametrize("half", [False, True])
def test_oracle_reproduces_neighbor_fixture(half):
    g = load("neighbor_tstneighbor.npz")
    tag = "half" if half else "full"
    x, n_local, rc = g["x"], int(g["n_local"]), float(g["rc"])
    # Verlet (cell) build of the oracle == frozen brute-force sets
    nl = O.NeighList().build(x, n_local, rc, half, [float(g["lo"])] * 3, [float(g["hi"])] * 3)
    c, _, _ = nl.arrays()
    assert np.array_equal(c, g[f"counts_{tag}"])
    for a, b in zip(nl.rows_sorted(), split_rows(g[f"counts_{tag}"], g[f"rows_{tag}"], n_local)):
        assert np.array_equal(a, b)
    # and the criterion itself, restated in numpy: d^2 <= rc^2 inclusive, i != j (tstNeighbor.hpp:97,130)
    d = x[:n_local, None, :] - x[None, :, :]
    d2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]
    m = d2 <= rc * rc
    m[np.arange(n_local), np.arange(n_local)] = False
    if half:
        xi, xj = x[:n_local, None, :], x[None, :, :]
        m &= (xj[..., 0] > xi[..., 0]) | ((xj[..., 0] == xi[..., 0]) & (
            (xj[..., 1] > xi[..., 1]) | ((xj[..., 1] == xi[..., 1]) & (xj[..., 2] > xi[..., 2]))))
    assert np.array_equal(m.sum(1), g[f"counts_{tag}"][:n_local])


@pytest.mark.parametrize("half", [False, True])
def test_oracle_reproduces_md_fixture(half):
    g = load("md_fcc8_100steps.npz")
    tag = "half" if half else "full"
    s = O.Sim(mass=[2.0], half=half).create_lattice_fcc(cells=(8, 8, 8)).setup()
    d0 = s.get()
    n = d0["n_local"]
    assert np.array_equal(d0["x"][:n], g["x0"]) and np.array_equal(d0["v"][:n], g["v0"])
    s.record_thermo()
    s.run(100, 10)
    assert np.abs(np.array(s.thermo()) - g[f"thermo_{tag}"]).max() < 1e-11
    # in.lj-style known answer at step 0 on the perfect lattice (SURVEY 8c iii)
    t0 = g[f"thermo_{tag}"][0]
    assert abs(t0[1] - 1.4) < 1e-12
    if not half:
        assert abs(t0[2] - (-6.332812)) < 5e-7


@pytest.mark.gpu
@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("layout", [0, 1])
def test_gpu_reproduces_neighbor_fixture(half, layout):
    import cabanamd_b200 as cb

    g = load("neighbor_tstneighbor.npz")
    tag = "half" if half else "full"
    x, n_local, rc = g["x"], int(g["n_local"]), float(g["rc"])
    ctx = cb.Context(0)
    ctx.set_domain([float(g["lo"])] * 3, [float(g["hi"])] * 3)
    ctx.set_atoms(x[:n_local])
    ctx.append_ghosts(x[n_local:])
    ctx.neigh_build(rc, half, layout, 100)
    counts, offsets, neigh = ctx.neigh_get()
    assert np.array_equal(counts, g[f"counts_{tag}"])  # bit-exact sets, ghost rows empty
    want = split_rows(g[f"counts_{tag}"], g[f"rows_{tag}"], n_local)
    for i in range(n_local):
        assert np.array_equal(np.sort(neigh[offsets[i]:offsets[i + 1]]), want[i])
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("half", [False, True])
def test_gpu_reproduces_md_fixture(half):
    from cabanamd_b200.harness import Simulation

    g = load("md_fcc8_100steps.npz")
    tag = "half" if half else "full"
    a = (4.0 / 0.8442) ** (1.0 / 3.0)
    sim = Simulation(device=0, half=half)
    sim.set_box([0.0] * 3, [8 * a] * 3)
    sim.set_atoms(g["x0"], g["v0"], np.zeros(len(g["x0"]), dtype=np.int32), g["id0"])
    sim.setup()
    sim.record_thermo()
    sim.run(100, 10)
    got = np.array(sim.thermo)
    assert np.abs(got - g[f"thermo_{tag}"]).max() < 1e-9   # T, PE/N, KE/N over 100 steps
    d = sim.ctx.get_atoms()
    n = d["n_local"]
    order = np.argsort(d["id"][:n])
    f, x = d["f"][:n][order], d["x"][:n][order]
    L = 8 * a
    dx = np.abs(x - g[f"x100_{tag}"])
    assert np.minimum(dx, np.abs(dx - L)).max() < 1e-9
    # FP64 force tolerance of the north star: 1e-10 relative (to the largest force), after
    # 100 steps of trajectory divergence allow 1e-8
    assert np.abs(f - g[f"f100_{tag}"]).max() / np.abs(g[f"f100_{tag}"]).max() < 1e-8
    pe_c = sim.potential(corrected=True)
    assert abs(pe_c - float(g[f"pe_corrected_{tag}"])) < 1e-7 * abs(pe_c)
    sim.ctx.close()
