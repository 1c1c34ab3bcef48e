"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol
include/cbmd_c_api.h declares, and fails loudly (no fallback) without a GPU."""
import os

import pytest

import cabanamd_b200 as cb


def test_library_exports_every_declared_symbol():
    L = cb.load_library()
    names = cb.declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert missing == []


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.CbmdError) as e:
        cb.Context(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for dp, _, fs in os.walk(os.path.join(root, "cabanamd_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if "oracle_lib" in txt or "liboracle" in txt or "oracle/" in txt or "oracle.hpp" in txt:
                    bad.append(f)
    assert bad == []


def test_domain_helper_matches_oracle():
    import numpy as np
    import oracle_lib as O
    from cabanamd_b200.capi import make_domain, dims_create

    s = O.Sim(mass=[2.0]).create_lattice_fcc(cells=(12, 10, 8), nranks=8)
    glo = np.zeros(3)
    ghi = s.a * np.array([12.0, 10.0, 8.0])
    for rk in range(8):
        d, o = make_domain(glo, ghi, 8, rk, 20.0), s.domain(rk)
        for k in ("llo", "lhi", "ghost_lo", "ghost_hi", "grid", "pos"):
            assert np.array_equal(d[k], o[k]), (rk, k)
    assert dims_create(12) == (3, 2, 2)


@pytest.mark.parametrize("rows", [1, 7, 8, 50, 97])
def test_verlet_table_addressing(rows):
    """The table layout the build, force, energy and CSR-export kernels share (nb_entry in
    cbmd_internal.cuh): every (atom, entry) has its own slot inside the table, a 32-atom tile
    is one contiguous block, and a warp reads four entries per lane as one 512-byte request."""
    import ctypes as C

    import numpy as np

    L = cb.load_library()
    L.cbmd_table_offset.restype = C.c_int64
    L.cbmd_table_offset.argtypes = [C.c_int] * 3
    L.cbmd_table_size.restype = C.c_int64
    L.cbmd_table_size.argtypes = [C.c_int] * 2
    n_atoms = 75  # rounds up to 96 = 3 tiles
    size = L.cbmd_table_size(n_atoms, rows)
    rows4 = (rows + 3) // 4
    assert size == 96 * 4 * rows4                                # rows round up to a multiple of 4
    off = np.array([[L.cbmd_table_offset(i, n, rows) for n in range(rows)] for i in range(96)])
    assert off.min() >= 0 and off.max() < size
    assert len(np.unique(off)) == off.size                       # no two entries share a slot
    block = size // 3
    for t in range(3):                                           # tile = contiguous block
        o = off[32 * t:32 * t + 32]
        assert o.min() >= t * block and o.max() < (t + 1) * block
    # entries 4k..4k+3 of a lane are one aligned int4; the 32 lanes' int4s of group k are
    # one contiguous 512-byte run: lane = atom & 31
    for n in range(rows):
        assert np.array_equal(off[:32, n] - off[0, n], 4 * np.arange(32))
        assert off[0, n] % 128 == n % 4
        if n % 4:
            assert np.array_equal(off[:, n], off[:, n - 1] + 1)


def test_hub_objects_are_host_side_and_checked():
    """cbmd_hub_create / cbmd_hub_destroy need no device: creation, argument checks, destruction."""
    h = cb.Hub(4, timeout=1.0)
    assert h.nranks == 4
    h.close()
    h.close()  # idempotent
    with pytest.raises(cb.CbmdError, match="nranks"):
        cb.Hub(0)
    with pytest.raises(cb.CbmdError, match="nranks"):
        cb.Hub(65)


class _Recorder:
    """Stands in for capi.Context: records the module calls of the step loop."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def f(*a, **k):
            self.calls.append((name,) + tuple(a))
            return {"neigh_build": 50, "sum_mv2": 1.0, "reduce_sum": 1.0, "energy": (0.0, 0.0)}.get(name, 0)

        return f


def _expand(calls, half):
    """cbmd_md_steps(n, half) = n times the six module calls (include/cbmd_c_api.h)."""
    out = []
    for c in calls:
        if c[0] == "md_steps":
            step = [("integrate_initial",), ("update_halo",), ("zero_force",), ("force", half)]
            step += [("update_force",)] if half else []
            step += [("integrate_final",)]
            out += step * c[1]
        else:
            out.append(c)
    return out


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("nsteps,thermo", [(67, 10), (45, 0), (20, 1), (19, 7), (1, 0), (40, 20)])
def test_step_loop_batching_keeps_the_call_sequence(nsteps, thermo, half):
    """harness.Simulation.run hands stretches of plain steps to cbmd_md_steps; expanded, the sequence of
    module calls is exactly the stepwise one (the order of CbnMD::run, cabanamd_impl.h:285-399), and no
    rebuild or thermo step ever lands inside a stretch."""
    from cabanamd_b200.harness import Simulation

    seqs = []
    for batch in (False, True):
        sim = Simulation.__new__(Simulation)
        sim.ctx = _Recorder()
        sim.half, sim.rn, sim.exchange_rate = half, 2.8, 20
        sim.guess, sim.layout, sim.step, sim.N = 50, 0, 13, 100  # starts mid-period
        sim.mvv2e, sim.boltz, sim.thermo, sim.fuse_energy, sim.nranks = 1.0, 1.0, [], True, 1
        sim.run(nsteps, thermo, batch=batch)
        assert sim.step == 13 + nsteps
        seqs.append(sim.ctx.calls)
    stepwise, batched = seqs
    assert _expand(batched, half) == stepwise
    if nsteps > 3 and thermo != 1:
        assert any(c[0] == "md_steps" for c in batched)
    assert all(c[1] >= 1 for c in batched if c[0] == "md_steps")
