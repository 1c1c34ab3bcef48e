"""ctypes access to oracle/liboracle.so — the CPU checker (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (cabanamd_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_int64)


def dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def lp(a):
    return None if a is None else a.ctypes.data_as(c_lp)


def build():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle.hpp", "oracle_capi.cpp", "Makefile")]
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    vp = C.c_void_p
    L.orc_list_new.restype = vp
    L.orc_list_free.argtypes = [vp]
    L.orc_list_total.argtypes = [vp]
    L.orc_list_total.restype = C.c_int64
    L.orc_list_max.argtypes = [vp]
    L.orc_list_copy.argtypes = [vp, c_ip, c_lp, c_ip]
    L.orc_list_set.argtypes = [vp, C.c_int, C.c_int, c_ip, c_lp, c_ip]
    L.orc_neigh_build.argtypes = [vp, c_dp, C.c_int, C.c_int, C.c_double, C.c_int, c_dp, c_dp]
    L.orc_neigh_brute.argtypes = [vp, c_dp, C.c_int, C.c_int, C.c_double, C.c_int]
    L.orc_force_lj.argtypes = [vp, c_dp, c_ip, c_dp, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp]
    L.orc_energy_lj.argtypes = [vp, c_dp, c_ip, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp]
    L.orc_energy_lj.restype = C.c_double
    L.orc_force_lj_f32.argtypes = [vp, c_dp, c_ip, c_dp, C.c_int, C.c_int, c_dp, c_dp, c_dp]
    L.orc_energy_lj_f32.argtypes = [vp, c_dp, c_ip, C.c_int, C.c_int, c_dp, c_dp, c_dp]
    L.orc_energy_lj_f32.restype = C.c_double
    L.orc_integrate.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_ip, C.c_int, C.c_int, c_dp,
                                C.c_double, C.c_double]
    L.orc_binning.argtypes = [c_dp, C.c_int, c_dp, c_dp, C.c_double, C.c_double, C.c_double,
                              C.c_int, c_ip, c_ip, c_dp, c_dp]
    L.orc_velocity_geom.argtypes = [C.c_int, c_dp, c_dp]
    L.orc_dims_create.argtypes = [C.c_int, c_ip]
    L.orc_kokkos_positions.argtypes = [C.c_ulonglong, C.c_int, C.c_double, C.c_double, c_dp]
    L.orc_sim_new.restype = vp
    L.orc_sim_new.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp, C.c_double, C.c_double, C.c_int,
                              C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
    L.orc_sim_free.argtypes = [vp]
    L.orc_sim_create_lattice_fcc.argtypes = [vp, C.c_double, c_dp, c_dp, C.c_int, C.c_double,
                                             C.c_int]
    L.orc_sim_set_atoms.argtypes = [vp, c_dp, c_dp, C.c_int, C.c_int, c_dp, c_dp, c_ip, c_ip]
    for name in ("setup", "binning", "exchange_halo", "update_halo", "neighbor", "force",
                 "initial_integrate", "final_integrate", "record_thermo"):
        getattr(L, "orc_sim_" + name).argtypes = [vp]
        getattr(L, "orc_sim_" + name).restype = None
    L.orc_sim_run.argtypes = [vp, C.c_int, C.c_int]
    L.orc_sim_exchange.argtypes = [vp]
    L.orc_sim_natoms.argtypes = [vp]
    L.orc_sim_nranks.argtypes = [vp]
    L.orc_sim_nlocal.argtypes = [vp, C.c_int]
    L.orc_sim_nghost.argtypes = [vp, C.c_int]
    L.orc_sim_domain.argtypes = [vp, C.c_int, c_dp, c_dp, c_dp, c_dp, c_ip, c_ip]
    L.orc_sim_get.argtypes = [vp, C.c_int, c_dp, c_dp, c_dp, c_ip, c_ip]
    L.orc_sim_list_total.argtypes = [vp, C.c_int]
    L.orc_sim_list_total.restype = C.c_int64
    L.orc_sim_list_copy.argtypes = [vp, C.c_int, c_ip, c_lp, c_ip]
    for name in ("temperature", "kinetic"):
        getattr(L, "orc_sim_" + name).argtypes = [vp]
        getattr(L, "orc_sim_" + name).restype = C.c_double
    L.orc_sim_potential.argtypes = [vp, C.c_int]
    L.orc_sim_potential.restype = C.c_double
    L.orc_sim_nthermo.argtypes = [vp]
    L.orc_sim_thermo.argtypes = [vp, C.c_int, c_ip, c_dp, c_dp, c_dp]
    L.orc_sim_timers.argtypes = [vp, c_dp]
    L.orc_set_threads.argtypes = [C.c_int]
    _LIB = L
    return L


def lj_tables(ntypes=1, eps=1.0, sigma=1.0, cut=2.5):
    """force_lj_cabana_neigh_impl.h:62-89 for a single (eps, sigma, cut) on every pair."""
    lj1 = np.full((ntypes, ntypes), 48.0 * eps * sigma ** 12.0)
    lj2 = np.full((ntypes, ntypes), 24.0 * eps * sigma ** 6.0)
    cutsq = np.full((ntypes, ntypes), cut * cut)
    return lj1, lj2, cutsq


class NeighList:
    """Host CSR list (counts[n_total], offsets[n_local+1], neigh[total])."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.orc_list_new()
        self.n_local = self.n_total = 0

    def __del__(self):
        try:
            self.L.orc_list_free(self.h)
        except Exception:
            pass

    def build(self, x, n_local, r, half, gmin, gmax):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.n_local, self.n_total = n_local, x.shape[0]
        gmin = np.ascontiguousarray(gmin, dtype=np.float64)
        gmax = np.ascontiguousarray(gmax, dtype=np.float64)
        self.L.orc_neigh_build(self.h, dp(x), n_local, x.shape[0], r, int(half), dp(gmin), dp(gmax))
        return self

    def brute(self, x, n_local, r, half):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.n_local, self.n_total = n_local, x.shape[0]
        self.L.orc_neigh_brute(self.h, dp(x), n_local, x.shape[0], r, int(half))
        return self

    def set(self, n_local, n_total, counts, offsets, neigh):
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        neigh = np.ascontiguousarray(neigh, dtype=np.int32)
        self.n_local, self.n_total = n_local, n_total
        self.L.orc_list_set(self.h, n_local, n_total, ip(counts), lp(offsets), ip(neigh))
        return self

    def arrays(self):
        tot = self.L.orc_list_total(self.h)
        counts = np.zeros(self.n_total, dtype=np.int32)
        offsets = np.zeros(self.n_local + 1, dtype=np.int64)
        neigh = np.zeros(max(tot, 1), dtype=np.int32)
        self.L.orc_list_copy(self.h, ip(counts), lp(offsets), ip(neigh))
        return counts, offsets, neigh[:tot]

    def rows_sorted(self):
        counts, offsets, neigh = self.arrays()
        return [np.sort(neigh[offsets[i]:offsets[i + 1]]) for i in range(self.n_local)]

    def force(self, x, type_, half, lj1, lj2, cutsq, f=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        if f is None:
            f = np.zeros_like(x)
        nt = lj1.shape[0]
        self.L.orc_force_lj(self.h, dp(x), ip(type_), dp(f), self.n_local, int(half), nt,
                            dp(np.ascontiguousarray(lj1)), dp(np.ascontiguousarray(lj2)),
                            dp(np.ascontiguousarray(cutsq)))
        return f

    def force_f32(self, x, type_, lj1, lj2, cutsq):
        """Full-list sweep of the reference's float build (T_X_FLOAT = T_F_FLOAT = float)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        f = np.zeros_like(x)
        nt = lj1.shape[0]
        self.L.orc_force_lj_f32(self.h, dp(x), ip(type_), dp(f), self.n_local, nt,
                                dp(np.ascontiguousarray(lj1)), dp(np.ascontiguousarray(lj2)),
                                dp(np.ascontiguousarray(cutsq)))
        return f

    def energy_f32(self, x, type_, lj1, lj2, cutsq):
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        nt = lj1.shape[0]
        return self.L.orc_energy_lj_f32(self.h, dp(x), ip(type_), self.n_local, nt,
                                        dp(np.ascontiguousarray(lj1)), dp(np.ascontiguousarray(lj2)),
                                        dp(np.ascontiguousarray(cutsq)))

    def energy(self, x, type_, half, lj1, lj2, cutsq, corrected=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        nt = lj1.shape[0]
        return self.L.orc_energy_lj(self.h, dp(x), ip(type_), self.n_local, int(half),
                                    int(corrected), nt, dp(np.ascontiguousarray(lj1)),
                                    dp(np.ascontiguousarray(lj2)), dp(np.ascontiguousarray(cutsq)))


def kokkos_positions(seed, n, lo, hi):
    """createAtoms of the reference's tstNeighbor (Kokkos XorShift64 pool, Serial backend)."""
    x = np.empty((n, 3))
    lib().orc_kokkos_positions(seed, n, lo, hi, dp(x))
    return x


def integrate(which, x, v, f, type_, mass, dt=0.005, mvv2e=1.0):
    L = lib()
    x = np.array(x, dtype=np.float64, copy=True)
    v = np.array(v, dtype=np.float64, copy=True)
    f = np.ascontiguousarray(f, dtype=np.float64)
    type_ = np.ascontiguousarray(type_, dtype=np.int32)
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    L.orc_integrate(which, dp(x), dp(v), dp(f), ip(type_), x.shape[0], mass.shape[0], dp(mass),
                    dt, mvv2e)
    return x, v


def binning(x, llo, lhi, delta, halo_depth=1):
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    perm = np.zeros(n, dtype=np.int32)
    nbin = np.zeros(3, dtype=np.int32)
    bmin = np.zeros(3)
    bmax = np.zeros(3)
    llo = np.ascontiguousarray(llo, dtype=np.float64)
    lhi = np.ascontiguousarray(lhi, dtype=np.float64)
    L.orc_binning(dp(x), n, dp(llo), dp(lhi), delta, delta, delta, halo_depth, ip(perm), ip(nbin),
                  dp(bmin), dp(bmax))
    return perm, nbin, bmin, bmax


class Sim:
    """The reference step loop over in-process virtual ranks (oracle.hpp: struct Sim)."""

    def __init__(self, ntypes=1, mass=(2.0,), eps=1.0, sigma=1.0, cut=2.5, skin=0.3, half=False,
                 exchange_rate=20, ghost_cutoff=20.0, dt=0.005, mvv2e=1.0, boltz=1.0,
                 force_cutoff=None, tables=None):
        self.L = lib()
        if tables is not None:  # explicit per-type-pair lj1, lj2, cutsq (multi-type decks)
            lj1, lj2, cutsq = (np.ascontiguousarray(t, dtype=np.float64) for t in tables)
        else:
            lj1, lj2, cutsq = lj_tables(ntypes, eps, sigma, cut)
        self.tables = (lj1, lj2, cutsq)
        mass = np.ascontiguousarray(mass, dtype=np.float64)
        fc = cut if force_cutoff is None else force_cutoff
        self.h = self.L.orc_sim_new(ntypes, dp(mass), dp(lj1), dp(lj2), dp(cutsq), fc, skin,
                                    int(half), exchange_rate, ghost_cutoff, dt, mvv2e, boltz)
        self.half = half

    def __del__(self):
        try:
            self.L.orc_sim_free(self.h)
        except Exception:
            pass

    def create_lattice_fcc(self, density=0.8442, cells=(40, 40, 40), nranks=1, temp=1.4,
                           seed=87287):
        a = (4.0 / density) ** (1.0 / 3.0)
        blo = np.zeros(3)
        bhi = np.array(cells, dtype=np.float64)
        self.L.orc_sim_create_lattice_fcc(self.h, a, dp(blo), dp(bhi), nranks, temp, seed)
        self.a = a
        return self

    def set_atoms(self, glo, ghi, nranks, x, v, type_, id_):
        glo = np.ascontiguousarray(glo, dtype=np.float64)
        ghi = np.ascontiguousarray(ghi, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        id_ = np.ascontiguousarray(id_, dtype=np.int32)
        self.L.orc_sim_set_atoms(self.h, dp(glo), dp(ghi), nranks, x.shape[0], dp(x), dp(v),
                                 ip(type_), ip(id_))
        return self

    def setup(self):
        self.L.orc_sim_setup(self.h)
        return self

    def run(self, nsteps, thermo_rate=0):
        self.L.orc_sim_run(self.h, nsteps, thermo_rate)

    def __getattr__(self, name):
        if name in ("binning", "exchange_halo", "update_halo", "neighbor", "force",
                    "initial_integrate", "final_integrate", "record_thermo"):
            fn = getattr(self.L, "orc_sim_" + name)
            return lambda: fn(self.h)
        raise AttributeError(name)

    def exchange(self):
        return self.L.orc_sim_exchange(self.h)

    @property
    def natoms(self):
        return self.L.orc_sim_natoms(self.h)

    @property
    def nranks(self):
        return self.L.orc_sim_nranks(self.h)

    def nlocal(self, rk=0):
        return self.L.orc_sim_nlocal(self.h, rk)

    def nghost(self, rk=0):
        return self.L.orc_sim_nghost(self.h, rk)

    def domain(self, rk=0):
        a = [np.zeros(3) for _ in range(4)]
        g = np.zeros(3, dtype=np.int32)
        p = np.zeros(3, dtype=np.int32)
        self.L.orc_sim_domain(self.h, rk, dp(a[0]), dp(a[1]), dp(a[2]), dp(a[3]), ip(g), ip(p))
        return dict(llo=a[0], lhi=a[1], ghost_lo=a[2], ghost_hi=a[3], grid=g, pos=p)

    def get(self, rk=0):
        n = self.nlocal(rk) + self.nghost(rk)
        x = np.zeros((n, 3))
        v = np.zeros((n, 3))
        f = np.zeros((n, 3))
        t = np.zeros(n, dtype=np.int32)
        i = np.zeros(n, dtype=np.int32)
        self.L.orc_sim_get(self.h, rk, dp(x), dp(v), dp(f), ip(t), ip(i))
        return dict(x=x, v=v, f=f, type=t, id=i, n_local=self.nlocal(rk), n_ghost=self.nghost(rk))

    def list(self, rk=0):
        n_local, n_total = self.nlocal(rk), self.nlocal(rk) + self.nghost(rk)
        tot = self.L.orc_sim_list_total(self.h, rk)
        counts = np.zeros(n_total, dtype=np.int32)
        offsets = np.zeros(n_local + 1, dtype=np.int64)
        neigh = np.zeros(max(tot, 1), dtype=np.int32)
        self.L.orc_sim_list_copy(self.h, rk, ip(counts), lp(offsets), ip(neigh))
        return counts, offsets, neigh[:tot]

    def temperature(self):
        return self.L.orc_sim_temperature(self.h)

    def kinetic(self):
        return self.L.orc_sim_kinetic(self.h)

    def potential(self, corrected=False):
        return self.L.orc_sim_potential(self.h, int(corrected))

    def thermo(self):
        out = []
        for k in range(self.L.orc_sim_nthermo(self.h)):
            s = C.c_int()
            T = C.c_double()
            pe = C.c_double()
            ke = C.c_double()
            self.L.orc_sim_thermo(self.h, k, C.byref(s), C.byref(T), C.byref(pe), C.byref(ke))
            out.append((s.value, T.value, pe.value, ke.value))
        return out

    def timers(self):
        t = np.zeros(5)
        self.L.orc_sim_timers(self.h, dp(t))
        return dict(force=t[0], neigh=t[1], comm=t[2], integrate=t[3], other=t[4])
