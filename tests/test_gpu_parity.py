"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Tolerances (north_star): neighbour sets bit-exact
after sorting; per-atom forces <= 1e-10 relative (FP64); integrator and ghost
refresh bit-identical; thermo to round-off."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

FORCE_RTOL = 1e-10


@pytest.fixture(scope="module")
def cb():
    import cabanamd_b200 as cb

    return cb


def tst_neighbor_config(seed=342343901):
    """unit_test/tstNeighbor.hpp:262-304 with the reference's own positions: the Kokkos XorShift64
    pool stream of the Serial backend (oracle/oracle.hpp KokkosXorShift64)."""
    rc = 2.32
    lo, hi = -5.3 * rc, 4.7 * rc
    x = O.kokkos_positions(seed, 1000, lo, hi)
    return x, 800, rc, lo, hi


def csr_sorted(counts, offsets, neigh, n_rows):
    """All rows of a CSR list, each sorted: one flat array (vectorised; for whole-list equality)."""
    c = np.asarray(counts[:n_rows], dtype=np.int64)
    tot = int(offsets[n_rows])
    rid = np.repeat(np.arange(n_rows, dtype=np.int64), c)
    key = rid * (1 << 31) + np.asarray(neigh[:tot], dtype=np.int64)
    key.sort()
    return key


def gpu_rows(ctx):
    counts, offsets, neigh = ctx.neigh_get()
    nl, _ = ctx.counts()
    return counts, [np.sort(neigh[offsets[i]:offsets[i + 1]]) for i in range(nl)]


@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("guess", [100, 7])
def test_tstneighbor_bitexact(cb, half, layout, guess):
    """unit_test/tstNeighbor.hpp:322-379 restated: set == O(N^2) brute force (and the
    exact half criterion), ghost rows empty; guess=7 exercises the 2-D regrow path."""
    x, n_local, rc, lo, hi = tst_neighbor_config()
    ctx = cb.Context(0)
    ctx.set_domain([lo] * 3, [hi] * 3)
    ctx.set_atoms(x[:n_local])
    ctx.append_ghosts(x[n_local:])
    g = ctx.neigh_build(rc, half, layout, guess)
    ref = O.NeighList().brute(x, n_local, rc, half)
    rc_counts, _, _ = ref.arrays()
    counts, rows = gpu_rows(ctx)
    assert np.array_equal(counts, rc_counts)
    assert np.all(counts[n_local:] == 0)
    for a, b in zip(rows, ref.rows_sorted()):
        assert np.array_equal(a, b)
    assert g >= counts.max()
    tot, mx = ctx.neigh_sizes()
    assert tot == counts.sum() and mx == counts.max()


def test_neighbor_edge_cases(cb):
    ctx = cb.Context(0)
    ctx.set_domain([0.0] * 3, [10.0] * 3)
    # empty system
    ctx.set_atoms(np.zeros((0, 3)))
    ctx.neigh_build(2.8, False, 0, 10)
    c, o, n = ctx.neigh_get()
    assert len(c) == 0 and len(n) == 0
    # one atom, and coincident-coordinate ties for the half discriminator
    x = np.array([[1.0, 1.0, 1.0], [1.0, 1.0, 2.0], [1.0, 2.0, 1.0], [3.8, 1.0, 1.0],
                  [3.8000001, 1.0, 1.0], [9.9, 9.9, 9.9]])
    for half in (False, True):
        ctx.set_atoms(x)
        ctx.neigh_build(2.8, half, 0, 2)
        ref = O.NeighList().brute(x, len(x), 2.8, half)
        counts, rows = gpu_rows(ctx)
        for a, b in zip(rows, ref.rows_sorted()):
            assert np.array_equal(a, b)
    # exact-cutoff pair: d^2 == rc^2 is inside (inclusive), list is stale after set_atoms
    x = np.array([[0.0, 0.0, 0.0], [3.0, 4.0, 0.0]])
    ctx.set_atoms(x)
    with pytest.raises(cb.CbmdError):
        ctx.force(False)
    ctx.neigh_build(5.0, False, 0, 4)
    counts, rows = gpu_rows(ctx)
    assert counts.tolist() == [1, 1]


def test_integrator_bitexact_and_reversible(cb):
    """unit_test/tstIntegrator.hpp:83-137 + bit equality with the oracle."""
    rng = np.random.default_rng(11)
    n = 1000
    x0 = rng.uniform(0, 20, size=(n, 3))
    v0 = rng.uniform(-1, 1, size=(n, 3))
    f0 = rng.uniform(-1, 1, size=(n, 3))
    t = rng.integers(0, 2, size=n).astype(np.int32)
    mass = [1.0, 3.0]
    ctx = cb.Context(0)
    ctx.set_units(1.0, 1.0, 0.005)
    ctx.set_mass(mass)
    ctx.set_domain([0.0] * 3, [20.0] * 3)
    ctx.set_atoms(x0, v0, f0, t)
    x, v = x0, v0
    for _ in range(100):
        ctx.integrate_initial()
        ctx.integrate_final()
        x, v = O.integrate(0, x, v, f0, t, mass)
        x, v = O.integrate(1, x, v, f0, t, mass)
    a = ctx.get_atoms()
    assert np.array_equal(a["x"], x) and np.array_equal(a["v"], v)
    assert np.array_equal(a["f"], f0) and np.array_equal(a["type"], t)
    ctx.set_velocities(-a["v"])
    for _ in range(100):
        ctx.integrate_initial()
        ctx.integrate_final()
    b = ctx.get_atoms()
    assert np.allclose(b["x"].astype(np.float32), x0.astype(np.float32), rtol=5e-7, atol=0)


def melted_state(cells=(10, 10, 10), steps=60, half=False):
    """A liquid-like configuration from the oracle (lattice + a few dozen steps)."""
    s = O.Sim(mass=[2.0], half=half).create_lattice_fcc(cells=cells).setup()
    s.run(steps)
    return s


def test_binning_matches_oracle(cb):
    s = melted_state((8, 8, 8), 25)
    d = s.get()
    n = d["n_local"]
    dom = s.domain()
    ctx = cb.Context(0)
    ctx.set_domain(dom["llo"], dom["lhi"])
    ctx.set_atoms(d["x"][:n], d["v"][:n], d["f"][:n], d["type"][:n], d["id"][:n])
    nbin, mn, mx = ctx.bin_sort(2.8)
    perm, onbin, omn, omx = O.binning(d["x"][:n], dom["llo"], dom["lhi"], 2.8)
    assert np.array_equal(nbin, onbin) and np.array_equal(mn, omn) and np.array_equal(mx, omx)
    assert np.array_equal(ctx.permutation(), perm)
    a = ctx.get_atoms()
    for k in ("x", "v", "f"):
        assert np.array_equal(a[k], d[k][:n][perm])
    assert np.array_equal(a["id"], d["id"][:n][perm])


# gather paths of the FP64 full-list sweep: 1 = mirror (xy LDG.128 + z TEX), 0 = 32-byte records
@pytest.mark.parametrize("gather", [1, 0])
@pytest.mark.parametrize("half", [False, True])
def test_force_energy_on_oracle_state(cb, half, gather):
    """Same atoms (owned + ghosts from the oracle's 6-phase build): neighbour sets
    bit-exact, forces <= 1e-10 relative, energy to round-off, for every sweep shape /
    gather path of the force kernel."""
    s = melted_state((10, 10, 10), 60, half)
    d = s.get()
    n, ng = d["n_local"], d["n_ghost"]
    dom = s.domain()
    lj1, lj2, cutsq = s.tables
    ctx = cb.Context(0)
    ctx.set_mass([2.0])
    ctx.set_lj(lj1, lj2, cutsq)
    ctx.set_domain(dom["llo"], dom["lhi"])
    ctx.set_atoms(d["x"][:n], d["v"][:n], None, d["type"][:n], d["id"][:n])
    ctx.append_ghosts(d["x"][n:], d["type"][n:], d["id"][n:])
    ctx.set_option("gather", gather)
    ctx.neigh_build(2.8, half, 0, 50)
    counts, rows = gpu_rows(ctx)
    ocounts, ooff, oneigh = s.list()
    assert np.array_equal(counts, ocounts)
    for i in range(n):
        assert np.array_equal(rows[i], np.sort(oneigh[ooff[i]:ooff[i + 1]]))
    # oracle force on its own list, before the reverse ghost fold (raw kernel output)
    ol = O.NeighList().set(n, n + ng, ocounts, ooff, oneigh)
    f_ref = ol.force(d["x"], d["type"], half, lj1, lj2, cutsq)
    ctx.zero_force()
    ctx.force(half)
    a = ctx.get_atoms()
    scale = np.abs(f_ref).max()
    assert np.abs(a["f"] - f_ref).max() <= FORCE_RTOL * scale
    nz = np.abs(f_ref) > 1e-3 * scale
    assert (np.abs(a["f"] - f_ref)[nz] / np.abs(f_ref)[nz]).max() <= FORCE_RTOL
    pe, pe_c = ctx.energy(half)
    e_ref = ol.energy(d["x"], d["type"], half, lj1, lj2, cutsq)
    e_cor = ol.energy(d["x"], d["type"], half, lj1, lj2, cutsq, corrected=True)
    assert abs(pe - e_ref) <= 1e-12 * abs(e_ref)
    assert abs(pe_c - e_cor) <= 1e-12 * abs(e_cor)
    # accumulate semantics: a second compute without zeroing doubles f
    ctx.force(half)
    b = ctx.get_atoms()
    assert np.abs(b["f"] - 2 * a["f"]).max() <= 1e-12 * scale
    if not half:
        # the two gather paths run the same arithmetic in the same order: identical bits
        ctx.set_option("gather", 1 - gather)
        ctx.zero_force()
        ctx.force(half)
        assert np.array_equal(ctx.get_atoms()["f"], a["f"])


@pytest.mark.parametrize("gather", [1, 0])
def test_multitype_force(cb, gather):
    rng = np.random.default_rng(5)
    s = melted_state((8, 8, 8), 40)
    d = s.get()
    n, ng = d["n_local"], d["n_ghost"]
    nt = 3
    t = rng.integers(0, nt, size=n + ng).astype(np.int32)
    lj1 = rng.uniform(20, 60, size=(nt, nt))
    lj1 = (lj1 + lj1.T) / 2
    lj2 = rng.uniform(10, 30, size=(nt, nt))
    lj2 = (lj2 + lj2.T) / 2
    cutsq = rng.uniform(4.0, 6.25, size=(nt, nt))
    cutsq = (cutsq + cutsq.T) / 2
    dom = s.domain()
    ctx = cb.Context(0)
    ctx.set_mass([1.0, 2.0, 3.0])
    ctx.set_lj(lj1, lj2, cutsq)
    ctx.set_domain(dom["llo"], dom["lhi"])
    ctx.set_atoms(d["x"][:n], None, None, t[:n])
    ctx.append_ghosts(d["x"][n:], t[n:])
    ctx.set_option("gather", gather)
    ctx.neigh_build(2.8, False, 0, 90)
    oc, oo, on = s.list()
    ol = O.NeighList().set(n, n + ng, oc, oo, on)
    f_ref = ol.force(d["x"], t, False, lj1, lj2, cutsq)
    ctx.zero_force()
    ctx.force(False)
    a = ctx.get_atoms()
    assert np.abs(a["f"] - f_ref).max() <= FORCE_RTOL * np.abs(f_ref).max()
    pe, _ = ctx.energy(False)
    e_ref = ol.energy(d["x"], t, False, lj1, lj2, cutsq)
    assert abs(pe - e_ref) <= 1e-12 * abs(e_ref)
    # fused force + energy sweep of the multi-type kernel, and the other gather path: same bits
    ctx.zero_force()
    ctx.request_energy()
    ctx.force(False)
    pe2, _ = ctx.energy(False)
    assert abs(pe2 - e_ref) <= 1e-12 * abs(e_ref)
    assert np.abs(ctx.get_atoms()["f"] - a["f"]).max() <= 1e-13 * np.abs(f_ref).max()
    ctx.set_option("gather", 1 - gather)
    ctx.zero_force()
    ctx.force(False)
    assert np.array_equal(ctx.get_atoms()["f"], a["f"])


def numpy_virial(x, t, counts, offsets, neigh, lj1, lj2, cutsq, once):
    """sum of rsq * fpair over the listed pairs inside the cutoff; `once`: every pair is stored
    once (half list), otherwise twice (full list -> factor 1/2)."""
    i = np.repeat(np.arange(len(counts)), counts)
    j = neigh[: offsets[-1]]
    d = x[i] - x[j]
    rsq = (d * d).sum(1)
    k1, k2, kc = lj1[t[i], t[j]], lj2[t[i], t[j]], cutsq[t[i], t[j]]
    m = rsq < kc
    r2 = 1.0 / rsq[m]
    r6 = r2 ** 3
    w = (rsq[m] * (r6 * (k1[m] * r6 - k2[m])) * r2).sum()
    return w if once else 0.5 * w


@pytest.mark.parametrize("half", [False, True])
def test_virial_matches_closed_form(cb, half):
    """The pair virial (extension, north_star) from the fused sweep and from the stand-alone sweep
    against a numpy evaluation of sum r_ij . f_ij over the oracle's list, and against the
    finite-difference identity W = -3 V dU/dV on the unshifted LJ energy (uniform scaling)."""
    s = melted_state((8, 8, 8), 40, half)
    d = s.get()
    n, ng = d["n_local"], d["n_ghost"]
    dom = s.domain()
    lj1, lj2, cutsq = s.tables
    oc, oo, on = s.list()
    w_ref = numpy_virial(d["x"], d["type"], oc[:n], oo, on, lj1, lj2, cutsq, half)
    ctx = cb.Context(0)
    ctx.set_mass([2.0])
    ctx.set_lj(lj1, lj2, cutsq)
    ctx.set_domain(dom["llo"], dom["lhi"])
    ctx.set_atoms(d["x"][:n], None, None, d["type"][:n], d["id"][:n])
    ctx.append_ghosts(d["x"][n:], d["type"][n:], d["id"][n:])
    ctx.neigh_build(2.8, half, 0, 50)
    w_alone = ctx.virial(half)                 # stand-alone sweep
    ctx.zero_force()
    ctx.request_energy()
    ctx.force(half)
    w_fused = ctx.virial(half)                 # cached by the fused sweep
    assert abs(w_alone - w_ref) <= 1e-11 * abs(w_ref)
    assert abs(w_fused - w_ref) <= 1e-11 * abs(w_ref)
    # virial theorem for pair forces: W = sum_i x_i . f_i over owned + ghost contributions;
    # with a half list the ghost rows carry their share (before the reverse fold)
    a = ctx.get_atoms()
    if half:
        w_xf = (a["x"] * a["f"]).sum()
        assert abs(w_xf - w_ref) <= 1e-9 * abs(w_ref)


def test_half_list_pull_sweep_is_atomics_free_and_deterministic(cb):
    """Newton-3 sweep: the default pulls the j side from the transposed list (no atomics).  It must
    agree with the oracle's half-list forces (owned AND ghost rows, before the reverse fold) to
    1e-10, conserve momentum, give bit-identical results run after run, and match the round-1
    RED.ADD.F64 scatter kernel (option half_kernel 0) to round-off."""
    s = melted_state((10, 10, 10), 60, True)
    d = s.get()
    n, ng = d["n_local"], d["n_ghost"]
    dom = s.domain()
    lj1, lj2, cutsq = s.tables
    oc, oo, on = s.list()
    ol = O.NeighList().set(n, n + ng, oc, oo, on)
    f_ref = ol.force(d["x"], d["type"], True, lj1, lj2, cutsq)
    e_ref = ol.energy(d["x"], d["type"], True, lj1, lj2, cutsq)
    out = {}
    for kernel in (1, 1, 0):
        ctx = cb.Context(0)
        ctx.set_mass([2.0])
        ctx.set_lj(lj1, lj2, cutsq)
        ctx.set_domain(dom["llo"], dom["lhi"])
        ctx.set_atoms(d["x"][:n], d["v"][:n], None, d["type"][:n], d["id"][:n])
        ctx.append_ghosts(d["x"][n:], d["type"][n:], d["id"][n:])
        ctx.set_option("half_kernel", kernel)
        ctx.neigh_build(2.8, True, 0, 50)
        ctx.zero_force()
        ctx.request_energy()
        ctx.force(True)
        f = ctx.get_atoms()["f"]
        pe, _ = ctx.energy(True)
        scale = np.abs(f_ref).max()
        assert np.abs(f - f_ref).max() <= FORCE_RTOL * scale, kernel
        assert np.abs(f.sum(0)).max() <= 1e-9 * scale          # Newton 3: owned + ghost rows cancel
        assert abs(pe - e_ref) <= 1e-12 * abs(e_ref)
        out.setdefault(kernel, []).append(f)
        ctx.close()
    assert np.array_equal(out[1][0], out[1][1])                 # same bits, run after run
    assert np.abs(out[1][0] - out[0][0]).max() <= 1e-12 * np.abs(f_ref).max()


FORCE_RTOL_F32 = 1e-5  # north_star: per-atom forces to 1e-5 relative in FP32


@pytest.mark.parametrize("ntypes", [1, 3])
def test_force_fp32_variant_matches_float_oracle(cb, ntypes):
    """Option precision=32 (the reference's T_X_FLOAT/T_F_FLOAT = float build for the force
    evaluation, types.h:133-148): float positions, FP32 pair terms and sums.  Checked against
    the float instantiation of the oracle's full-list sweep on the same list, 1e-5 relative;
    and against the FP64 sweep, where the difference is the float rounding of the positions."""
    rng = np.random.default_rng(11)
    s = melted_state((10, 10, 10), 60, False)
    d = s.get()
    n, ng = d["n_local"], d["n_ghost"]
    dom = s.domain()
    if ntypes == 1:
        lj1, lj2, cutsq = s.tables
        t = np.zeros(n + ng, dtype=np.int32)
        mass = [2.0]
    else:
        t = rng.integers(0, ntypes, size=n + ng).astype(np.int32)
        lj1 = rng.uniform(20, 60, size=(ntypes, ntypes)); lj1 = (lj1 + lj1.T) / 2
        lj2 = rng.uniform(10, 30, size=(ntypes, ntypes)); lj2 = (lj2 + lj2.T) / 2
        cutsq = rng.uniform(4.0, 6.25, size=(ntypes, ntypes)); cutsq = (cutsq + cutsq.T) / 2
        mass = [1.0, 2.0, 3.0]
    ctx = cb.Context(0)
    ctx.set_mass(mass)
    ctx.set_lj(lj1, lj2, cutsq)
    ctx.set_domain(dom["llo"], dom["lhi"])
    ctx.set_atoms(d["x"][:n], None, None, t[:n], d["id"][:n])
    ctx.append_ghosts(d["x"][n:], t[n:], d["id"][n:])
    ctx.set_option("precision", 32)
    ctx.neigh_build(2.8, False, 0, 90)
    oc, oo, on = s.list()
    ol = O.NeighList().set(n, n + ng, oc, oo, on)
    f32 = ol.force_f32(d["x"], t, lj1, lj2, cutsq)
    f64 = ol.force(d["x"], t, False, lj1, lj2, cutsq)
    ctx.zero_force()
    ctx.request_energy()
    ctx.force(False)
    a = ctx.get_atoms()
    scale = np.abs(f32).max()
    assert np.abs(a["f"][:n] - f32[:n]).max() <= FORCE_RTOL_F32 * scale
    assert np.all(a["f"][n:] == 0.0)                      # full list: ghost rows untouched
    # against FP64: only the float rounding of x (2^-24 * |x| ~ 1e-6) and of the sums
    assert np.abs(a["f"][:n] - f64[:n]).max() <= 2e-3 * np.abs(f64).max()
    pe, _ = ctx.energy(False)
    e32 = ol.energy_f32(d["x"], t, lj1, lj2, cutsq)
    assert abs(pe - e32) <= 1e-5 * abs(e32)
    # accumulate semantics: a second sweep without zeroing doubles f
    ctx.force(False)
    b = ctx.get_atoms()
    assert np.abs(b["f"] - 2 * a["f"]).max() <= 1e-12 * scale
    # back to FP64: the 1e-10 bar again on the same context
    ctx.set_option("precision", 64)
    ctx.zero_force()
    ctx.force(False)
    c = ctx.get_atoms()
    assert np.abs(c["f"] - f64).max() <= FORCE_RTOL * np.abs(f64).max()


def test_fp32_variant_trajectory_conserves_energy(cb):
    """NVE run with the FP32 force sweep (FP64 integration state): total energy stays within
    the float force noise, and T/PE track the FP64 run over the first steps."""
    from cabanamd_b200.harness import Simulation

    ref = O.Sim(mass=[2.0]).create_lattice_fcc(cells=(8, 8, 8))
    d, dom = ref.get(), ref.domain()
    out = {}
    for prec in (64, 32):
        sim = Simulation(device=0)
        sim.ctx.set_option("precision", prec)
        sim.set_box(dom["llo"], dom["lhi"])
        sim.set_atoms(d["x"], d["v"], d["type"], d["id"])
        sim.setup()
        sim.record_thermo()
        sim.run(200, 10)
        out[prec] = np.array(sim.thermo)
    e64 = out[64][:, 2] + out[64][:, 3]
    e32 = out[32][:, 2] + out[32][:, 3]
    assert np.abs(e32 - e32[0]).max() < 5e-4               # drift of the FP32-force trajectory
    assert np.abs(e64 - e64[0]).max() < 5e-4
    assert np.abs(out[32][:3, 1:] - out[64][:3, 1:]).max() < 1e-5   # same physics at the start


def canon(x, ids):
    o = np.argsort(ids, kind="stable")
    return x[o], ids[o]


@pytest.mark.parametrize("half", [False, True])
def test_halo_build_and_update_match_oracle(cb, half):
    """T6: ghost set == oracle's 6-phase set (same order: ordered compaction), and
    update_halo is bitwise owner + shift."""
    from cabanamd_b200.harness import Simulation

    s = melted_state((8, 8, 8), 19, half)
    d = s.get()
    n = d["n_local"]
    dom = s.domain()
    sim = Simulation(half=half)
    sim.set_box(dom["llo"], dom["lhi"])
    sim.set_atoms(d["x"][:n], d["v"][:n], d["type"][:n], d["id"][:n])
    # oracle: one more step hits the rebuild at step 20
    s.run(1)
    sim.ctx.set_atoms(d["x"][:n], d["v"][:n], d["f"][:n], d["type"][:n], d["id"][:n])
    sim.step = 19
    sim.run(1)
    o = s.get()
    a = sim.ctx.get_atoms()
    assert a["n_local"] == o["n_local"] and a["n_ghost"] == o["n_ghost"]
    assert np.array_equal(a["id"], o["id"])  # same owned order AND same ghost order
    assert np.array_equal(a["x"], o["x"])  # bitwise: integrate + wrap + sort + ghosts
    # three more steps: update_halo path
    s.run(3)
    sim.run(3)
    o = s.get()
    a = sim.ctx.get_atoms()
    assert np.array_equal(a["id"], o["id"])
    assert np.abs(a["x"] - o["x"]).max() < 1e-12
    # ghosts are exact images of their owners
    nl = a["n_local"]
    L = dom["lhi"] - dom["llo"]
    pos_of = {i: k for k, i in enumerate(a["id"][:nl])}
    own = np.array([pos_of[i] for i in a["id"][nl:]])
    k = np.round((a["x"][nl:] - a["x"][own]) / L)
    assert np.abs(k).max() == 1
    assert np.array_equal(a["x"][nl:], np.where(k == 0, a["x"][own], a["x"][own] + k * L))


@pytest.mark.parametrize("half", [False, True])
def test_trajectory_and_thermo_match_oracle(cb, half):
    """T8 (short form; the 1000-step form runs in bench/): same initial state, 100
    steps with rebuilds every 20: thermo agrees to round-off growth, ids identical."""
    from cabanamd_b200.harness import Simulation

    s0 = O.Sim(mass=[2.0], half=half).create_lattice_fcc(cells=(10, 10, 10))
    d = s0.get()
    dom = s0.domain()
    s0.setup()
    sim = Simulation(half=half)
    sim.set_box(dom["llo"], dom["lhi"])
    sim.set_atoms(d["x"], d["v"], d["type"], d["id"])
    sim.setup()
    s0.record_thermo()
    sim.record_thermo()
    s0.run(100, 10)
    sim.run(100, 10)
    tg, to = np.array(sim.thermo), np.array(s0.thermo())
    assert np.array_equal(tg[:, 0], to[:, 0])
    assert np.abs(tg[:, 1:] - to[:, 1:]).max() < 1e-9
    # step-0 values agree to summation-order round-off: the oracle adds ~2e5 pair
    # energies sequentially into a sum of magnitude 2.5e4 (half-ulp 1.8e-12 per add),
    # the GPU uses a two-level tree, so per-atom PE may differ by a few 1e-12
    assert np.abs(tg[0, 1:] - to[0, 1:]).max() < 2e-11
    a, o = sim.ctx.get_atoms(), s0.get()
    assert np.array_equal(np.sort(a["id"][: a["n_local"]]), np.sort(o["id"][: o["n_local"]]))
    xa, _ = canon(a["x"][: a["n_local"]], a["id"][: a["n_local"]])
    xo, _ = canon(o["x"][: o["n_local"]], o["id"][: o["n_local"]])
    assert np.abs(xa - xo).max() < 1e-8


@pytest.mark.parametrize("half", [False, True])
def test_batched_plain_steps_are_bit_identical_to_stepwise(cb, half):
    """cbmd_md_steps (stretches of plain steps in one call, replayed from a CUDA graph on one rank) against
    the six module calls per step: same kernels, arguments and order, so x, v, f and the thermo trace are
    bit-identical — with the graph (default) and without it (option graph_steps 0)."""
    from cabanamd_b200.harness import Simulation

    s0 = O.Sim(mass=[2.0], half=half).create_lattice_fcc(cells=(8, 8, 8))
    d, dom = s0.get(), s0.domain()
    runs = []
    for mode in ("stepwise", "graph", "nograph"):
        sim = Simulation(half=half)
        if mode == "nograph":
            sim.ctx.set_option("graph_steps", 0)
        sim.set_box(dom["llo"], dom["lhi"])
        sim.set_atoms(d["x"], d["v"], d["type"], d["id"])
        sim.setup()
        sim.record_thermo()
        l0 = sim.ctx.launch_count()
        sim.run(67, 10, batch=(mode != "stepwise"))  # ends inside a stretch, 3 rebuilds, 6 thermo steps
        a = sim.ctx.get_atoms()
        runs.append((a, np.array(sim.thermo), sim.ctx.launch_count() - l0))
        sim.ctx.close()
    ref = runs[0]
    for a, th, _ in runs[1:]:
        assert np.array_equal(th, ref[1])
        for k in ("x", "v", "f", "id"):
            assert np.array_equal(a[k], ref[0][k]), k
    # the replayed steps are counted as the launches they contain
    assert runs[1][2] == runs[2][2] == runs[0][2]


def test_inlj_step0_known_answer_gpu(cb):
    """T5: in.lj (256 000 atoms) step-0 thermo: 1.400000 / -6.332812 / -4.232820."""
    from cabanamd_b200.harness import Simulation, fcc_lattice

    s0 = O.Sim(mass=[2.0]).create_lattice_fcc(cells=(40, 40, 40))
    d = s0.get()
    x, a = fcc_lattice((40, 40, 40))
    assert np.array_equal(x, d["x"])  # harness lattice == oracle lattice, bitwise
    dom = s0.domain()
    sim = Simulation()
    sim.set_box(dom["llo"], dom["lhi"])
    sim.set_atoms(d["x"], d["v"], d["type"], d["id"])
    sim.setup()
    T, pe, ke = sim.temperature(), sim.potential() / sim.N, sim.kinetic() / sim.N
    assert f"{T:.6f}" == "1.400000"
    assert f"{pe:.6f}" == "-6.332812"
    assert f"{pe + ke:.6f}" == "-4.232820"
    counts, _, _ = sim.ctx.neigh_get()
    assert np.all(counts[:256000] == 78)
    assert sim.guess == int(78 * 1.1)


def test_newton3_properties_full_size(cb):
    """Size-independent properties at the bench size (1 M atoms): sum f = 0, half ==
    full after the reverse fold, neighbour symmetry via counts."""
    from cabanamd_b200.harness import Simulation, fcc_lattice

    x, a = fcc_lattice((63, 63, 63))
    rng = np.random.default_rng(1)
    x = x + rng.uniform(-0.05, 0.05, size=x.shape)
    L = 63 * a
    x = np.mod(x, L)
    out = {}
    for half in (False, True):
        sim = Simulation(half=half, max_neigh_guess=100)
        sim.set_box([0.0] * 3, [L] * 3)
        sim.set_atoms(x, np.zeros_like(x))
        sim.setup()
        g = sim.ctx.get_atoms(fields="fi")
        nl = g["n_local"]
        o = np.argsort(g["id"][:nl])
        out[half] = g["f"][:nl][o]
        tot, mx = sim.ctx.neigh_sizes()
        out[("tot", half)] = tot
    scale = np.abs(out[False]).max()
    assert np.abs(out[False].sum(axis=0)).max() < 1e-9 * scale * 1000
    assert np.abs(out[False] - out[True]).max() < 1e-10 * scale
    assert out[("tot", False)] == 2 * out[("tot", True)]


@pytest.mark.parametrize("half", [False, True])
def test_fused_energy_matches_standalone(cb, half):
    """cbmd_request_energy: the PE accumulated inside the force sweep equals the
    stand-alone compute_energy sweep, and the cache is dropped as soon as atoms move."""
    s = melted_state((8, 8, 8), 40, half)
    d = s.get()
    n, ng = d["n_local"], d["n_ghost"]
    dom = s.domain()
    lj1, lj2, cutsq = s.tables
    ctx = cb.Context(0)
    ctx.set_mass([2.0])
    ctx.set_lj(lj1, lj2, cutsq)
    ctx.set_domain(dom["llo"], dom["lhi"])
    ctx.set_atoms(d["x"][:n], d["v"][:n], None, d["type"][:n], d["id"][:n])
    ctx.append_ghosts(d["x"][n:], d["type"][n:], d["id"][n:])
    ctx.neigh_build(2.8, half, 0, 90)
    ctx.zero_force()
    ctx.force(half)
    f_plain = ctx.get_atoms()["f"]
    pe0, pec0 = ctx.energy(half)           # stand-alone sweep
    l0 = ctx.launch_count()
    ctx.zero_force()
    ctx.request_energy()
    ctx.force(half)
    l1 = ctx.launch_count()
    pe1, pec1 = ctx.energy(half)           # served from the fused sweep: no new kernel
    assert ctx.launch_count() == l1 and l1 > l0
    assert abs(pe1 - pe0) <= 5e-13 * abs(pe0) and abs(pec1 - pec0) <= 5e-13 * abs(pec0)
    f_fused = ctx.get_atoms()["f"]
    if not half:  # half-list forces go through FP64 atomics: order-dependent last bits
        assert np.array_equal(f_fused, f_plain)
    else:
        assert np.abs(f_fused - f_plain).max() <= 1e-12 * np.abs(f_plain).max()
    e_ref = O.NeighList().set(n, n + ng, *s.list()).energy(d["x"], d["type"], half, lj1, lj2, cutsq)
    assert abs(pe1 - e_ref) <= 1e-12 * abs(e_ref)
    # the hint is one-shot and the cache dies when positions change
    ctx.integrate_initial()
    l2 = ctx.launch_count()
    ctx.energy(half)
    assert ctx.launch_count() > l2


def test_neighbor_build_hard_cases(cb):
    """Cases aimed at the staged / FP32-prefiltered build: (a) a dense blob whose
    27-cell stencil exceeds one staging chunk, (b) pairs within 1e-12 of the cutoff on
    both sides (FP32-ambiguous -> exact FP64 fallback), (c) a perfect lattice (equal-x
    ties for the half discriminator), (d) a box thinner than the cutoff."""
    rng = np.random.default_rng(77)
    # (a) 3000 atoms inside a 3x3x3 sigma blob + 500 spread out; rc 1.0
    blob = rng.uniform(4.0, 7.0, size=(3000, 3))
    rest = rng.uniform(0.0, 12.0, size=(500, 3))
    x = np.concatenate([blob, rest])
    for half in (False, True):
        ctx = cb.Context(0)
        ctx.set_domain([0.0] * 3, [12.0] * 3)
        ctx.set_atoms(x[:3200])
        ctx.append_ghosts(x[3200:])
        ctx.neigh_build(1.0, half, 0, 64)
        ref = O.NeighList().brute(x, 3200, 1.0, half)
        counts, rows = gpu_rows(ctx)
        assert np.array_equal(counts, ref.arrays()[0])
        for a, b in zip(rows, ref.rows_sorted()):
            assert np.array_equal(a, b)
    # (b) shells of partners at r = rc*(1 +- k*2^-50) around random centres
    rc = 2.5
    centres = rng.uniform(3.0, 17.0, size=(40, 3))
    pts = [centres]
    for k in range(-6, 7):
        u = rng.normal(size=(40, 3))
        u /= np.linalg.norm(u, axis=1)[:, None]
        pts.append(centres + u * rc * (1.0 + k * 2.0 ** -50))
    x = np.concatenate(pts)
    for half in (False, True):
        ctx = cb.Context(0)
        ctx.set_domain([0.0] * 3, [20.0] * 3)
        ctx.set_atoms(x)
        ctx.neigh_build(rc, half, 1, 8)
        ref = O.NeighList().brute(x, len(x), rc, half)
        counts, rows = gpu_rows(ctx)
        assert np.array_equal(counts, ref.arrays()[0])
        for a, b in zip(rows, ref.rows_sorted()):
            assert np.array_equal(a, b)
    # (c) perfect sc lattice, spacing 1.0, rc exactly sqrt(2) -> many d^2 == rc^2 and x ties
    g = np.arange(8, dtype=np.float64)
    x = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3) + 0.5
    for half in (False, True):
        ctx = cb.Context(0)
        ctx.set_domain([0.0] * 3, [8.0] * 3)
        ctx.set_atoms(x)
        ctx.neigh_build(np.sqrt(2.0), half, 0, 20)
        ref = O.NeighList().brute(x, len(x), np.sqrt(2.0), half)
        counts, rows = gpu_rows(ctx)
        assert np.array_equal(counts, ref.arrays()[0])
        for a, b in zip(rows, ref.rows_sorted()):
            assert np.array_equal(a, b)
    # (d) slab: z extent 1.5 < rc 2.8, long in x
    x = rng.uniform([0, 0, 0], [300.0, 9.0, 1.5], size=(4000, 3))
    ctx = cb.Context(0)
    ctx.set_domain([0.0] * 3, [300.0, 9.0, 1.5])
    ctx.set_atoms(x)
    ctx.neigh_build(2.8, False, 0, 40)
    ref = O.NeighList().brute(x, len(x), 2.8, False)
    counts, rows = gpu_rows(ctx)
    assert np.array_equal(counts, ref.arrays()[0])
    for a, b in zip(rows, ref.rows_sorted()):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("half", [False, True])
def test_1000_step_thermo_and_energy_drift_match_oracle(cb, half):
    """T8 (north star: "thermo energies must agree over 1000 steps with matching energy
    drift"): 4 000 atoms, 1000 NVE steps, thermo every 10.  The trajectories are chaotic,
    so round-off differences grow; the thermo trace still agrees to 1e-6 and the total
    energy drift (last - first ETot, and its rms fluctuation) agrees to 1e-7."""
    from cabanamd_b200.harness import Simulation

    s0 = O.Sim(mass=[2.0], half=half).create_lattice_fcc(cells=(10, 10, 10))
    d, dom = s0.get(), s0.domain()
    s0.setup()
    sim = Simulation(half=half)
    sim.set_box(dom["llo"], dom["lhi"])
    sim.set_atoms(d["x"], d["v"], d["type"], d["id"])
    sim.setup()
    s0.record_thermo()
    sim.record_thermo()
    s0.run(1000, 10)
    sim.run(1000, 10)
    tg, to = np.array(sim.thermo), np.array(s0.thermo())
    assert tg.shape == to.shape == (101, 4)
    assert np.abs(tg[:, 1:] - to[:, 1:]).max() < 1e-6
    eg, eo = tg[:, 2] + tg[:, 3], to[:, 2] + to[:, 3]
    drift_g, drift_o = eg[-1] - eg[0], eo[-1] - eo[0]
    assert abs(drift_g - drift_o) < 1e-7
    assert abs(eg.std() - eo.std()) < 1e-7
    if not half:  # full-list PE is the physical one: NVE conserves ETot to ~1e-4 per atom
        assert abs(drift_g) < 5e-4


def test_long_cutoff_variant_matches_oracle(cb):
    """BASELINE configs[4]: rc = 5.0 sigma, skin 0.3 (~526 stored neighbours per atom, row
    capacity regrown from the 'one 600' guess): sets bit-exact, forces 1e-10, thermo."""
    from cabanamd_b200.harness import Simulation

    s0 = O.Sim(mass=[2.0], cut=5.0).create_lattice_fcc(cells=(8, 8, 8))
    d, dom = s0.get(), s0.domain()
    s0.setup()
    sim = Simulation(cut=5.0, max_neigh_guess=600)
    sim.set_box(dom["llo"], dom["lhi"])
    sim.set_atoms(d["x"], d["v"], d["type"], d["id"])
    sim.setup()
    c, o, n = sim.ctx.neigh_get()
    oc, oo, on = s0.list()
    assert np.array_equal(c, oc) and c[: len(oo) - 1].mean() > 500
    nl = len(oo) - 1
    assert np.array_equal(csr_sorted(c, o, n, nl), csr_sorted(oc, oo, on, nl))   # every row
    s0.record_thermo()
    sim.record_thermo()
    s0.run(40, 10)
    sim.run(40, 10)
    assert np.abs(np.array(sim.thermo) - np.array(s0.thermo())).max() < 1e-9
    a, b = sim.ctx.get_atoms(), s0.get()
    fa = a["f"][: a["n_local"]][np.argsort(a["id"][: a["n_local"]])]
    fb = b["f"][: b["n_local"]][np.argsort(b["id"][: b["n_local"]])]
    assert np.abs(fa - fb).max() <= 1e-9 * np.abs(fb).max()


@pytest.mark.parametrize("half", [False, True])
def test_config1_one_million_atoms_vs_oracle(cb, half):
    """BASELINE configs[1] at its full size (fcc 63^3 = 1 000 188 atoms, rc 2.5, skin 0.3), full and
    half list: every neighbour row equal to the oracle's, forces 1e-10, thermo after a list period
    (one rebuild) to round-off."""
    from cabanamd_b200.harness import Simulation

    s0 = O.Sim(mass=[2.0], half=half).create_lattice_fcc(cells=(63, 63, 63))
    d, dom = s0.get(), s0.domain()
    s0.setup()
    sim = Simulation(half=half)
    sim.set_box(dom["llo"], dom["lhi"])
    sim.set_atoms(d["x"], d["v"], d["type"], d["id"])
    sim.setup()
    c, o, n = sim.ctx.neigh_get()
    oc, oo, on = s0.list()
    nl = len(oo) - 1
    assert nl == 1000188 and np.array_equal(c, oc)
    assert np.array_equal(csr_sorted(c, o, n, nl), csr_sorted(oc, oo, on, nl))
    s0.record_thermo()
    sim.record_thermo()
    s0.run(25, 5)
    sim.run(25, 5)
    assert np.abs(np.array(sim.thermo) - np.array(s0.thermo())).max() < 1e-9
    a, b = sim.ctx.get_atoms(fields="fi"), s0.get()
    fa = a["f"][: a["n_local"]][np.argsort(a["id"][: a["n_local"]])]
    fb = b["f"][: b["n_local"]][np.argsort(b["id"][: b["n_local"]])]
    assert np.abs(fa - fb).max() <= 1e-9 * np.abs(fb).max()


def test_config4_one_million_atoms_long_cutoff_vs_oracle(cb):
    """BASELINE configs[4] at its full size (1 000 188 atoms, rc 5.0, ~530 stored neighbours per
    atom): row lengths of every atom, every 97th row as a set, forces 1e-10 and thermo."""
    from cabanamd_b200.harness import Simulation

    s0 = O.Sim(mass=[2.0], cut=5.0).create_lattice_fcc(cells=(63, 63, 63))
    d, dom = s0.get(), s0.domain()
    s0.setup()
    sim = Simulation(cut=5.0, max_neigh_guess=600)
    sim.set_box(dom["llo"], dom["lhi"])
    sim.set_atoms(d["x"], d["v"], d["type"], d["id"])
    sim.setup()
    c, o, n = sim.ctx.neigh_get()
    oc, oo, on = s0.list()
    nl = len(oo) - 1
    assert np.array_equal(c, oc) and c[:nl].mean() > 500
    for i in range(0, nl, 97):
        assert np.array_equal(np.sort(n[o[i]:o[i + 1]]), np.sort(on[oo[i]:oo[i + 1]]))
    s0.record_thermo()
    sim.record_thermo()
    s0.run(5, 5)
    sim.run(5, 5)
    assert np.abs(np.array(sim.thermo) - np.array(s0.thermo())).max() < 1e-9
    a, b = sim.ctx.get_atoms(fields="fi"), s0.get()
    fa = a["f"][: a["n_local"]][np.argsort(a["id"][: a["n_local"]])]
    fb = b["f"][: b["n_local"]][np.argsort(b["id"][: b["n_local"]])]
    assert np.abs(fa - fb).max() <= 1e-9 * np.abs(fb).max()


def test_position_mirror_follows_every_way_atoms_move(cb):
    """The texture-assisted force kernel reads a split copy of the positions that the
    integrator and the halo refresh maintain; every other path (re-upload with a larger
    capacity, migration, cell sort, ghost rebuild, switching the option off and on) must
    leave it detectably stale.  After each such sequence the forces have to be bit-identical
    to the record-gather kernel's (same arithmetic, same order)."""
    from cabanamd_b200.harness import Simulation

    def forces(sim, gather):
        sim.ctx.set_option("gather", gather)
        sim.ctx.zero_force()
        sim.ctx.force(False)
        a = sim.ctx.get_atoms()
        return a["f"][: a["n_local"]].copy()

    def check(sim, what):
        f_tex = forces(sim, 1)
        f_rec = forces(sim, 0)
        assert np.array_equal(f_tex, f_rec), what
        sim.ctx.set_option("gather", 1)

    small = melted_state((6, 6, 6), 30)
    d, dom = small.get(), small.domain()
    n = d["n_local"]
    sim = Simulation(device=0)
    sim.set_box(dom["llo"], dom["lhi"])
    sim.set_atoms(d["x"][:n], d["v"][:n], d["type"][:n], d["id"][:n])
    sim.setup()
    check(sim, "after setup")
    sim.run(7, 0)                       # integrator + halo refresh keep the mirror current
    check(sim, "after 7 steps")
    sim.ctx.set_option("gather", 0)     # mirror not maintained while the option is off
    sim.run(5, 0)
    check(sim, "after 5 steps with the option off")
    sim.run(25, 0)                      # crosses a rebuild: migration, sort, ghost rebuild
    check(sim, "after a rebuild step")
    # a larger system through the same context: capacity grows, mirror + texture are re-made
    big = melted_state((9, 9, 9), 20)
    d, dom = big.get(), big.domain()
    n = d["n_local"]
    sim.set_box(dom["llo"], dom["lhi"])
    sim.set_atoms(d["x"][:n], d["v"][:n], d["type"][:n], d["id"][:n])
    sim.setup()
    check(sim, "after re-upload with a larger capacity")
    sim.run(21, 0)
    check(sim, "after 21 more steps")
    # and the trajectory itself still matches the oracle
    big.run(21)
    a, b = sim.ctx.get_atoms(), big.get()
    xa = a["x"][: a["n_local"]][np.argsort(a["id"][: a["n_local"]])]
    xb = b["x"][: b["n_local"]][np.argsort(b["id"][: b["n_local"]])]
    L = np.array(dom["lhi"]) - np.array(dom["llo"])
    dx = (xa - xb + L / 2) % L - L / 2
    assert np.abs(dx).max() < 1e-9


def test_lammps_bench_lj_step100_known_answer_gpu(cb):
    """The CUDA path against the published LAMMPS `bench/in.lj` log, step 100:
    Temp 0.7574531, E_pair -5.7585055, TotEng -4.6223613 (see tests/test_oracle.py)."""
    from test_oracle import LAMMPS_BENCH_LJ_STEP100, unshifted_thermo
    from cabanamd_b200.harness import Simulation

    ref = O.Sim(mass=[1.0]).create_lattice_fcc(cells=(20, 20, 20), temp=1.44)
    d, dom = ref.get(), ref.domain()
    sim = Simulation(device=0, mass=(1.0,))
    sim.set_box(dom["llo"], dom["lhi"])
    sim.set_atoms(d["x"], d["v"], d["type"], d["id"])
    sim.setup()
    sim.run(100, 0)
    n = sim.N
    _, off, nb = sim.ctx.neigh_get()
    x_all = sim.ctx.get_atoms(fields="x")["x"]
    got = unshifted_thermo(x_all, off, nb, n, sim.temperature(), sim.potential() / n, sim.kinetic() / n)
    assert got == LAMMPS_BENCH_LJ_STEP100
