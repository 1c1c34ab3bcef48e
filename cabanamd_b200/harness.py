"""Python re-statement of the reference's call ORDER (CbnMD::init / run,
cabanamd_impl.h:197-243 and :285-399) over the C ABI, for tests and bench.py.
The product driver is the C++ host layer (cabanamd_b200/host); this harness only
sequences the same C-ABI calls from Python."""
import numpy as np

from .capi import Context, make_domain


def lj_tables(ntypes=1, eps=1.0, sigma=1.0, cut=2.5):
    lj1 = np.full((ntypes, ntypes), 48.0 * eps * sigma ** 12.0)
    lj2 = np.full((ntypes, ntypes), 24.0 * eps * sigma ** 6.0)
    cutsq = np.full((ntypes, ntypes), cut * cut)
    return lj1, lj2, cutsq


def fcc_lattice(cells, density=0.8442):
    """Positions of create_lattice's fcc branch on one rank (inputFile_impl.h:716-792):
    loops iz, iy, ix then the 4 basis sites; x = a*(i+basis)."""
    a = (4.0 / density) ** (1.0 / 3.0)
    nx, ny, nz = cells
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=np.float64)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    cell = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(np.float64)
    x = a * (1.0 * cell[:, None, :] + basis[None, :, :])
    return x.reshape(-1, 3), a


class Simulation:
    """One rank of the MD loop on one GPU context."""

    def __init__(self, ctx=None, device=0, mass=(2.0,), cut=2.5, skin=0.3, half=False,
                 exchange_rate=20, ghost_cutoff=20.0, dt=0.005, mvv2e=1.0, boltz=1.0,
                 max_neigh_guess=50, layout=0, nranks=1, rank=0, uid=None, eps=1.0, sigma=1.0,
                 precision=64):
        self.ctx = ctx or Context(device)
        self.half, self.cut, self.skin = half, cut, skin
        self.rn = cut + skin
        self.exchange_rate = exchange_rate
        self.ghost_cutoff = ghost_cutoff
        self.boltz, self.mvv2e, self.dt = boltz, mvv2e, dt
        self.guess, self.layout = max_neigh_guess, layout
        self.nranks, self.rank = nranks, rank
        self.mass = np.asarray(mass, dtype=np.float64)
        c = self.ctx
        c.set_units(boltz, mvv2e, dt)
        c.set_mass(self.mass)
        c.set_lj(*lj_tables(len(self.mass), eps, sigma, cut))
        c.comm_init(nranks, rank, uid)
        if precision != 64:
            c.set_option("precision", precision)  # FP32 force sweep (full lists)
        self.step = 0
        self.N = 0
        self.thermo = []
        self.fuse_energy = True

    def set_box(self, glo, ghi):
        d = make_domain(glo, ghi, self.nranks, self.rank, self.ghost_cutoff)
        self.dom = d
        self.ctx.set_domain(d["glo"], d["ghi"], d["llo"], d["lhi"], d["ghost_lo"], d["ghost_hi"],
                            d["grid"], d["pos"])

    def set_atoms(self, x, v, type_=None, id_=None, n_global=None):
        """Upload this rank's owned atoms (callers filter by llo <= x < lhi)."""
        self.ctx.set_atoms(x, v, None, type_, id_)
        n = self.ctx.reduce_sum_int(len(x)) if self.nranks > 1 else len(x)
        self.N = n if n_global is None else n_global

    def owned_mask(self, x):
        d = self.dom
        m = np.ones(len(x), dtype=bool)
        for k in range(3):
            last = d["pos"][k] == d["grid"][k] - 1
            hi_ok = (x[:, k] <= d["lhi"][k]) if last else (x[:, k] < d["lhi"][k])
            m &= (x[:, k] >= d["llo"][k]) & hi_ok
        return m

    # cabanamd_impl.h:197-223
    def setup(self):
        c = self.ctx
        c.exchange()
        c.bin_sort(self.rn)
        c.exchange_halo(self.rn)
        self.guess = c.neigh_build(self.rn, self.half, self.layout, self.guess)
        c.zero_force()
        c.force(self.half)
        if self.half:
            c.update_force()
        self.step = 0

    # cabanamd_impl.h:285-399 (one step)
    def _special(self, step, thermo_rate):
        return step % self.exchange_rate == 0 or bool(thermo_rate and step % thermo_rate == 0)

    def run(self, nsteps, thermo_rate=0, batch=True):
        """batch: stretches of plain steps (no rebuild, no thermo) go through ONE C-ABI call
        (cbmd_md_steps: the same entry points in the same order inside the library)."""
        c = self.ctx
        end = self.step + nsteps
        while self.step < end:
            if batch and not self._special(self.step + 1, thermo_rate):
                k = 1
                while self.step + k < end and not self._special(self.step + k + 1, thermo_rate):
                    k += 1
                c.md_steps(k, self.half)
                self.step += k
                continue
            self.step += 1
            c.integrate_initial()
            if self.step % self.exchange_rate == 0:
                c.exchange()
                c.bin_sort(self.rn)
                c.exchange_halo(self.rn)
                self.guess = c.neigh_build(self.rn, self.half, self.layout, self.guess)
            else:
                c.update_halo()
            c.zero_force()
            if thermo_rate and self.step % thermo_rate == 0 and self.fuse_energy:
                c.request_energy()  # thermo step: PE in the same sweep as the force
            c.force(self.half)
            if self.half:
                c.update_force()
            c.integrate_final()
            if thermo_rate and self.step % thermo_rate == 0:
                self.record_thermo()

    # property_temperature_impl.h:55-77, property_kine_impl.h:55-76, property_pote_impl.h:55-63
    def temperature(self):
        s = self.ctx.reduce_sum(self.ctx.sum_mv2())
        return s * (self.mvv2e / ((3 * self.N - 3) * self.boltz))

    def kinetic(self):
        return self.ctx.reduce_sum(self.ctx.sum_mv2()) * 0.5 * self.mvv2e

    def potential(self, corrected=False):
        pe, pe_c = self.ctx.energy(self.half)
        return self.ctx.reduce_sum(pe_c if corrected else pe)

    def record_thermo(self):
        self.thermo.append((self.step, self.temperature(), self.potential() / self.N,
                            self.kinetic() / self.N))


def velocity_geom_uniform(x, seed):
    """Vectorised LAMMPS_RandomVelocityGeom (inputFile.h:73-148): per atom, seed =
    Jenkins one-at-a-time hash of the 4 bytes of `seed` and the 24 bytes of (x,y,z)
    (bytes added as SIGNED char), masked to 27 bits; 5 warm-up draws; returns the
    three uniform() draws per atom."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.shape[0]
    M = np.uint64(0xFFFFFFFF)
    h = np.zeros(n, dtype=np.uint64)

    def mix(h, byte_vals):
        # hash += (signed char) b  -> as unsigned 32-bit wraparound
        h = (h + (byte_vals.astype(np.int64).astype(np.uint64) & M)) & M
        h = (h + ((h << np.uint64(10)) & M)) & M
        h = h ^ (h >> np.uint64(6))
        return h

    sb = np.frombuffer(np.array([seed], dtype=np.int32).tobytes(), dtype=np.int8)
    for b in sb:
        h = mix(h, np.full(n, b, dtype=np.int8))
    xb = x.view(np.int8).reshape(n, 24)
    for k in range(24):
        h = mix(h, xb[:, k])
    h = (h + ((h << np.uint64(3)) & M)) & M
    h = h ^ (h >> np.uint64(11))
    h = (h + ((h << np.uint64(15)) & M)) & M
    s = (h & np.uint64(0x7FFFFFF)).astype(np.int64)
    s[s == 0] = 1
    IA, IM, IQ, IR = 16807, 2147483647, 127773, 2836

    def draw(s):
        k = s // IQ
        s = IA * (s - k * IQ) - IR * k
        s = np.where(s < 0, s + IM, s)
        return s

    for _ in range(5):
        s = draw(s)
    out = np.empty((n, 3))
    for c in range(3):
        s = draw(s)
        out[:, c] = (1.0 / IM) * s
    return out


def create_velocities(sim, x, type_, temp=1.4, seed=87287):
    """inputFile_impl.h:811-865: v = (u-0.5)/sqrt(m), subtract the global centre-of-mass
    velocity, rescale to `temp` with T computed on the device (as the reference does)."""
    m = sim.mass[type_]
    u = velocity_geom_uniform(x, seed)
    v = (u - 0.5) / np.sqrt(m)[:, None]
    tot = np.array([m.sum(), (m * v[:, 0]).sum(), (m * v[:, 1]).sum(), (m * v[:, 2]).sum()])
    if sim.nranks > 1:
        tot = np.array([sim.ctx.reduce_sum(t) for t in tot])
    v = v - tot[1:] / tot[0]
    return v
