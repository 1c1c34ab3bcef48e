"""ctypes binding of include/cbmd_c_api.h (test / bench harness only)."""
import ctypes as C
import os
import re

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_int64)
vp = C.c_void_p


class CbmdError(RuntimeError):
    pass


def library_path():
    return os.path.join(_PKG, "lib", "libcbmd_cuda.so")


def header_path():
    return os.path.join(_ROOT, "include", "cbmd_c_api.h")


def declared_symbols():
    """Every function include/cbmd_c_api.h declares."""
    txt = open(header_path()).read()
    return sorted(set(re.findall(r"\b(cbmd_[a-z0-9_]+)\s*\(", txt)))


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def _lp(a):
    return None if a is None else a.ctypes.data_as(c_lp)


def load_library():
    """Load libcbmd_cuda.so; raises (never falls back) when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise CbmdError(
            f"{path} not found: build it with `make -C cabanamd_b200/csrc` "
            "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    L.cbmd_last_error.restype = C.c_char_p
    L.cbmd_version.restype = C.c_char_p
    L.cbmd_create.argtypes = [C.POINTER(vp), C.c_int]
    L.cbmd_destroy.argtypes = [vp]
    L.cbmd_sync.argtypes = [vp]
    L.cbmd_set_units.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    L.cbmd_set_mass.argtypes = [vp, C.c_int, c_dp]
    L.cbmd_set_domain.argtypes = [vp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_ip, c_ip]
    L.cbmd_set_atoms.argtypes = [vp, C.c_int, c_dp, c_dp, c_dp, c_ip, c_ip, c_dp]
    L.cbmd_append_ghosts.argtypes = [vp, C.c_int, c_dp, c_ip, c_ip]
    L.cbmd_get_atoms.argtypes = [vp, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_ip, c_ip, c_dp]
    L.cbmd_set_velocities.argtypes = [vp, C.c_int, c_dp]
    L.cbmd_get_counts.argtypes = [vp, c_ip, c_ip]
    L.cbmd_integrate_initial.argtypes = [vp]
    L.cbmd_integrate_final.argtypes = [vp]
    L.cbmd_bin_sort.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int, c_ip, c_dp, c_dp]
    L.cbmd_get_permutation.argtypes = [vp, c_ip]
    L.cbmd_neigh_build.argtypes = [vp, C.c_double, C.c_int, C.c_int, C.c_int, c_ip]
    L.cbmd_neigh_sizes.argtypes = [vp, c_lp, c_ip]
    L.cbmd_neigh_get.argtypes = [vp, c_ip, c_lp, c_ip]
    L.cbmd_set_lj.argtypes = [vp, C.c_int, c_dp, c_dp, c_dp]
    L.cbmd_zero_force.argtypes = [vp]
    L.cbmd_force_lj.argtypes = [vp, C.c_int]
    L.cbmd_energy_lj.argtypes = [vp, C.c_int, c_dp, c_dp]
    L.cbmd_virial_lj.argtypes = [vp, C.c_int, c_dp]
    L.cbmd_request_energy.argtypes = [vp]
    L.cbmd_md_steps.argtypes = [vp, C.c_int, C.c_int]
    L.cbmd_comm_unique_id.argtypes = [vp]
    L.cbmd_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.cbmd_comm_rank.argtypes = [vp, c_ip, c_ip]
    L.cbmd_hub_create.argtypes = [C.POINTER(vp), C.c_int, C.c_double]
    L.cbmd_hub_destroy.argtypes = [vp]
    L.cbmd_comm_init_hub.argtypes = [vp, vp, C.c_int]
    L.cbmd_exchange.argtypes = [vp, c_ip]
    L.cbmd_exchange_halo.argtypes = [vp, C.c_double]
    L.cbmd_update_halo.argtypes = [vp]
    L.cbmd_update_force.argtypes = [vp]
    L.cbmd_reduce_sum_double.argtypes = [vp, c_dp, C.c_int]
    L.cbmd_reduce_sum_int.argtypes = [vp, c_ip, C.c_int]
    L.cbmd_reduce_max_double.argtypes = [vp, c_dp, C.c_int]
    L.cbmd_reduce_max_int.argtypes = [vp, c_ip, C.c_int]
    L.cbmd_scan_sum_int.argtypes = [vp, c_ip, C.c_int]
    L.cbmd_sum_mv2.argtypes = [vp, c_dp]
    L.cbmd_stream.argtypes = [vp]
    L.cbmd_stream.restype = vp
    L.cbmd_launch_count.argtypes = [vp]
    L.cbmd_launch_count.restype = C.c_int64
    L.cbmd_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.cbmd_timing_enable.argtypes = [vp, C.c_int]
    L.cbmd_timing_get.argtypes = [vp, C.c_int, c_dp, c_lp]
    L.cbmd_timing_reset.argtypes = [vp]
    _LIB = L
    return L


def dims_create(n):
    """MPI_Dims_create(n, 3): balanced factors, non-increasing (system.h:154-155)."""
    best = (n, 1, 1)
    for a in range(1, n + 1):
        if n % a:
            continue
        for b in range(1, n // a + 1):
            if (n // a) % b:
                continue
            t = tuple(sorted((a, b, n // a // b), reverse=True))
            if (t[0] - t[2], t[0]) < (best[0] - best[2], best[0]):
                best = t
    return best


def make_domain(glo, ghi, nranks, rank, ghost_cutoff):
    """SystemCommon::create_domain (system.h:149-205,251-271) for one rank."""
    glo = np.asarray(glo, dtype=np.float64)
    ghi = np.asarray(ghi, dtype=np.float64)
    grid = np.array(dims_create(nranks), dtype=np.int32)
    pos = np.zeros(3, dtype=np.int32)
    r = rank
    pos[2] = r % grid[2]
    r //= grid[2]
    pos[1] = r % grid[1]
    r //= grid[1]
    pos[0] = r
    cell = (ghi - glo) / (100 * grid)
    halo = int(np.ceil(ghost_cutoff / cell.min()))
    off = 100 * pos
    return dict(glo=glo, ghi=ghi, grid=grid, pos=pos,
                llo=glo + cell * off, lhi=glo + cell * (off + 100),
                ghost_lo=glo + cell * (off - halo), ghost_hi=glo + cell * (off + 100 + halo))


class Hub:
    """In-process transport between `nranks` contexts, each driven by its own host thread
    (cbmd_hub_create).  Lets one GPU run the decomposed path at 2/4/8 ranks."""

    def __init__(self, nranks, timeout=120.0):
        self.L = load_library()
        self.nranks = int(nranks)
        h = C.c_void_p()
        if self.L.cbmd_hub_create(C.byref(h), self.nranks, float(timeout)) != 0:
            raise CbmdError(self.L.cbmd_last_error().decode())
        self.h = h

    def close(self):
        if self.h:
            if self.L.cbmd_hub_destroy(self.h) != 0:
                raise CbmdError(self.L.cbmd_last_error().decode())
            self.h = None


class Context:
    """One device context == one rank's System + modules (thin, 1:1 over the C ABI)."""

    def __init__(self, device=0):
        self.L = load_library()
        h = vp()
        if self.L.cbmd_create(C.byref(h), device) != 0:
            raise CbmdError(self.L.cbmd_last_error().decode())
        self.h = h
        self.half = False

    def close(self):
        if getattr(self, "h", None):
            self.L.cbmd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise CbmdError(self.L.cbmd_last_error().decode())

    # ---- System
    def set_units(self, boltz=1.0, mvv2e=1.0, dt=0.005):
        self._ck(self.L.cbmd_set_units(self.h, boltz, mvv2e, dt))

    def set_mass(self, mass):
        m = np.ascontiguousarray(mass, dtype=np.float64)
        self._ck(self.L.cbmd_set_mass(self.h, len(m), _dp(m)))

    def set_domain(self, glo, ghi, llo=None, lhi=None, ghost_lo=None, ghost_hi=None, grid=None,
                   pos=None):
        f = lambda a, d: np.ascontiguousarray(d if a is None else a, dtype=np.float64)
        glo, ghi = f(glo, None), f(ghi, None)
        llo, lhi = f(llo, glo), f(lhi, ghi)
        ghost_lo, ghost_hi = f(ghost_lo, llo), f(ghost_hi, lhi)
        grid = np.ascontiguousarray([1, 1, 1] if grid is None else grid, dtype=np.int32)
        pos = np.ascontiguousarray([0, 0, 0] if pos is None else pos, dtype=np.int32)
        self._ck(self.L.cbmd_set_domain(self.h, _dp(glo), _dp(ghi), _dp(llo), _dp(lhi),
                                        _dp(ghost_lo), _dp(ghost_hi), _ip(grid), _ip(pos)))

    def set_atoms(self, x, v=None, f=None, type_=None, id_=None, q=None):
        c = lambda a, t: None if a is None else np.ascontiguousarray(a, dtype=t)
        x = c(x, np.float64)
        v, f, q = c(v, np.float64), c(f, np.float64), c(q, np.float64)
        type_, id_ = c(type_, np.int32), c(id_, np.int32)
        self._ck(self.L.cbmd_set_atoms(self.h, x.shape[0], _dp(x), _dp(v), _dp(f), _ip(type_),
                                       _ip(id_), _dp(q)))

    def append_ghosts(self, x, type_=None, id_=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = None if type_ is None else np.ascontiguousarray(type_, dtype=np.int32)
        id_ = None if id_ is None else np.ascontiguousarray(id_, dtype=np.int32)
        self._ck(self.L.cbmd_append_ghosts(self.h, x.shape[0], _dp(x), _ip(type_), _ip(id_)))

    def counts(self):
        a, b = C.c_int(), C.c_int()
        self._ck(self.L.cbmd_get_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def get_atoms(self, first=0, count=None, fields="xvftiq"):
        nl, ng = self.counts()
        if count is None:
            count = nl + ng - first
        out = {}
        x = np.zeros((count, 3)) if "x" in fields else None
        v = np.zeros((count, 3)) if "v" in fields else None
        f = np.zeros((count, 3)) if "f" in fields else None
        t = np.zeros(count, dtype=np.int32) if "t" in fields else None
        i = np.zeros(count, dtype=np.int32) if "i" in fields else None
        q = np.zeros(count) if "q" in fields else None
        self._ck(self.L.cbmd_get_atoms(self.h, first, count, _dp(x), _dp(v), _dp(f), _ip(t),
                                       _ip(i), _dp(q)))
        out.update(x=x, v=v, f=f, type=t, id=i, q=q, n_local=nl, n_ghost=ng)
        return out

    def set_velocities(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        self._ck(self.L.cbmd_set_velocities(self.h, v.shape[0], _dp(v)))

    # ---- Integrator
    def integrate_initial(self):
        self._ck(self.L.cbmd_integrate_initial(self.h))

    def integrate_final(self):
        self._ck(self.L.cbmd_integrate_final(self.h))

    # ---- Binning
    def bin_sort(self, dx, dy=None, dz=None, halo_depth=1):
        dy = dx if dy is None else dy
        dz = dx if dz is None else dz
        nbin = np.zeros(3, dtype=np.int32)
        mn, mx = np.zeros(3), np.zeros(3)
        self._ck(self.L.cbmd_bin_sort(self.h, dx, dy, dz, halo_depth, _ip(nbin), _dp(mn), _dp(mx)))
        return nbin, mn, mx

    def permutation(self):
        nl, _ = self.counts()
        p = np.zeros(nl, dtype=np.int32)
        self._ck(self.L.cbmd_get_permutation(self.h, _ip(p)))
        return p

    # ---- Neighbor
    def neigh_build(self, rcut, half=False, layout=0, max_neigh_guess=50):
        g = C.c_int()
        self._ck(self.L.cbmd_neigh_build(self.h, rcut, int(half), layout, max_neigh_guess,
                                         C.byref(g)))
        self.half = bool(half)
        return g.value

    def neigh_sizes(self):
        t, m = C.c_int64(), C.c_int()
        self._ck(self.L.cbmd_neigh_sizes(self.h, C.byref(t), C.byref(m)))
        return t.value, m.value

    def neigh_get(self):
        nl, ng = self.counts()
        tot, _ = self.neigh_sizes()
        counts = np.zeros(nl + ng, dtype=np.int32)
        offsets = np.zeros(nl + 1, dtype=np.int64)
        neigh = np.zeros(max(tot, 1), dtype=np.int32)
        self._ck(self.L.cbmd_neigh_get(self.h, _ip(counts), _lp(offsets), _ip(neigh)))
        return counts, offsets, neigh[:tot]

    # ---- Force
    def set_lj(self, lj1, lj2, cutsq):
        a = [np.ascontiguousarray(t, dtype=np.float64) for t in (lj1, lj2, cutsq)]
        self._ck(self.L.cbmd_set_lj(self.h, a[0].shape[0], _dp(a[0]), _dp(a[1]), _dp(a[2])))

    def zero_force(self):
        self._ck(self.L.cbmd_zero_force(self.h))

    def force(self, half=None):
        self._ck(self.L.cbmd_force_lj(self.h, int(self.half if half is None else half)))

    def request_energy(self):
        self._ck(self.L.cbmd_request_energy(self.h))

    def energy(self, half=None):
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.cbmd_energy_lj(self.h, int(self.half if half is None else half),
                                       C.byref(a), C.byref(b)))
        return a.value, b.value

    def virial(self, half=None):
        """Scalar pair virial sum r_ij . f_ij of this rank (extension; see cbmd_c_api.h)."""
        a = C.c_double()
        self._ck(self.L.cbmd_virial_lj(self.h, int(self.half if half is None else half), C.byref(a)))
        return a.value

    def md_steps(self, nsteps, half=False):
        """nsteps plain MD steps (no rebuild, no thermo) in one call; graph replay on one rank."""
        self._ck(self.L.cbmd_md_steps(self.h, int(nsteps), int(half)))

    # ---- Comm
    @staticmethod
    def unique_id():
        L = load_library()
        buf = C.create_string_buffer(128)
        if L.cbmd_comm_unique_id(buf) != 0:
            raise CbmdError(L.cbmd_last_error().decode())
        return buf.raw

    def comm_init(self, nranks=1, rank=0, uid=None):
        """uid: the 128-byte NCCL id (one process per GPU), or a Hub (ranks = host threads of
        this process, in-process transport)."""
        if isinstance(uid, Hub):
            assert nranks == uid.nranks, (nranks, uid.nranks)
            self._ck(self.L.cbmd_comm_init_hub(self.h, uid.h, rank))
            self._hub = uid  # keep the hub alive as long as this context is attached
            return
        buf = None if uid is None else C.create_string_buffer(uid, 128)
        self._ck(self.L.cbmd_comm_init(self.h, nranks, rank, buf))

    def exchange(self):
        n = C.c_int()
        self._ck(self.L.cbmd_exchange(self.h, C.byref(n)))
        return n.value

    def exchange_halo(self, depth):
        self._ck(self.L.cbmd_exchange_halo(self.h, depth))

    def update_halo(self):
        self._ck(self.L.cbmd_update_halo(self.h))

    def update_force(self):
        self._ck(self.L.cbmd_update_force(self.h))

    def reduce_sum(self, val):
        a = np.array([val], dtype=np.float64)
        self._ck(self.L.cbmd_reduce_sum_double(self.h, _dp(a), 1))
        return float(a[0])

    def reduce_sum_int(self, val):
        a = np.array([val], dtype=np.int32)
        self._ck(self.L.cbmd_reduce_sum_int(self.h, _ip(a), 1))
        return int(a[0])

    def scan_sum_int(self, val):
        a = np.array([val], dtype=np.int32)
        self._ck(self.L.cbmd_scan_sum_int(self.h, _ip(a), 1))
        return int(a[0])

    # ---- thermo
    def sum_mv2(self):
        a = C.c_double()
        self._ck(self.L.cbmd_sum_mv2(self.h, C.byref(a)))
        return a.value

    # ---- misc
    def sync(self):
        self._ck(self.L.cbmd_sync(self.h))

    def stream(self):
        return self.L.cbmd_stream(self.h)

    def launch_count(self):
        return self.L.cbmd_launch_count(self.h)

    BUCKETS = ("force", "neigh", "comm", "integrate", "other", "force_kernel")

    def timing_enable(self, on=True):
        self._ck(self.L.cbmd_timing_enable(self.h, int(on)))

    def timing_reset(self):
        self._ck(self.L.cbmd_timing_reset(self.h))

    def timing(self):
        """{bucket: (milliseconds, regions)} accumulated since the last reset."""
        out = {}
        for b, name in enumerate(self.BUCKETS):
            ms, n = C.c_double(), C.c_int64()
            self._ck(self.L.cbmd_timing_get(self.h, b, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    def set_option(self, name, value):
        self._ck(self.L.cbmd_set_option(self.h, name.encode(), float(value)))
