"""The C++ host layer (cabanamd_b200/host: reference class surface + cbnMD driver).

CPU tests: the driver builds, parses decks with the reference's error convention, and
fails loudly without a GPU.  GPU tests: `cbnMD -il <deck>` reproduces the oracle's thermo
trace and the reference's output layout."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CBNMD = os.path.join(ROOT, "cabanamd_b200", "lib", "cbnMD")

DECK = """# 3d Lennard-Jones melt
units           lj
atom_style      atomic

newton          off
lattice         fcc 0.8442
region          box block 0 {c} 0 {c} 0 {c}
create_box      1 box
create_atoms    1 box
mass            1 2.0

velocity        all create 1.4 87287 loop geom

pair_style      lj/cut 2.5
pair_coeff      1 1 1.0 1.0 2.5

neighbor        0.3 bin
neigh_modify    every 20 one 50
comm_modify     cutoff * 20
fix             1 all nve
thermo          10

dump            dmpvtk all vtk 10 dump%_*.vtu

run             {steps}
"""


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cabanamd_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cabanamd_b200", "host")])
    assert os.path.exists(CBNMD)


def run_cbnmd(tmp_path, deck, *args, env=None, exe=None):
    f = tmp_path / "in.deck"
    f.write_text(deck)
    out, err = tmp_path / "md.out", tmp_path / "md.err"
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([exe or CBNMD, "-il", str(f), "-o", str(out), "-e", str(err), *args],
                       capture_output=True, text=True, cwd=tmp_path, env=e, timeout=600)
    return p, (out.read_text() if out.exists() else ""), (err.read_text() if err.exists() else "")


def test_unknown_keyword_is_an_input_error(tmp_path):
    p, out, err = run_cbnmd(tmp_path, "units lj\nfoo bar\n")
    assert p.returncode != 0
    assert "Unknown input file keyword: foo bar" in err
    assert "Aborting after error from input. See error file." in p.stderr
    # the deck is echoed up to (not including) the offending line
    assert "#InputFile:" in out and "units lj" in out and "foo bar" not in out


@pytest.mark.parametrize("line,msg", [
    ("variable x equal 1", "'variable' keyword is not supported"),
    ("units si", "'units' command only supports"),
    ("lattice bcc 1.0", "'lattice' command only supports 'sc' and 'fcc'"),
    ("region box sphere 0 0 0 1", "'region' command only supports 'block'"),
    ("pair_style eam", "'pair_style' command only supports"),
    ("neigh_modify delay 5", "'neigh_modify' only supports 'every' and 'one'"),
    ("comm_modify mode multi", "'comm_modify' command only supports single cutoff"),
    ("fix 1 all nvt", "'fix' command only supports 'nve'"),
    ("velocity all set 0 0 0", "'velocity' command can only be used with option 'create'"),
    ("newton maybe", "'newton' must be followed by 'on' or 'off'"),
    ("dump d all custom 10 f.txt", "'dump' command only supports 'vtk'"),
    ("dump d all vtk 10 f.vtu", "requires '*' in file name"),
])
def test_deck_errors_follow_the_reference(tmp_path, line, msg):
    p, out, err = run_cbnmd(tmp_path, line + "\n")
    assert p.returncode != 0 and msg in err


def test_unknown_cli_argument(tmp_path):
    p, _, _ = run_cbnmd(tmp_path, "units lj\n", "--bogus")
    assert p.returncode != 0 and "Unknown command line argument: --bogus" in p.stdout
    p, _, _ = run_cbnmd(tmp_path, "units lj\n", "--kokkos-threads=4", "--device-type", "OPENMP")
    assert p.returncode != 0 and "no CPU fallback" in p.stderr


def test_no_gpu_fails_loudly(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p, out, err = run_cbnmd(tmp_path, DECK.format(c=4, steps=1))
    assert p.returncode != 0
    assert "no CUDA device available" in p.stderr and "no CPU fallback" in p.stderr
    assert "Read input file." not in out  # stops at context creation


# ------------------------------------------------------------------ GPU
THERMO = re.compile(r"^(\d+)\t(-?\d+\.\d{6})\t(-?\d+\.\d{6})\t(-?\d+\.\d{6})\t(\d+\.\d{2})\t(\d\.\d{2}e[+-]\d{2})$")


def parse_thermo(out):
    rows = []
    for ln in out.splitlines():
        m = THERMO.match(ln)
        if m:
            rows.append((int(m.group(1)), float(m.group(2)), float(m.group(3)), float(m.group(4))))
    return rows


@pytest.mark.gpu
def test_cbnmd_fp32_build_variant(tmp_path):
    """cbnMD_f32 (-DCBMD_SINGLE_PRECISION, the reference's T_F_FLOAT/T_X_FLOAT = float build): same
    deck, thermo rows track the FP64 binary's to float accuracy and the run conserves energy."""
    deck = DECK.format(c=10, steps=100)
    p64, out64, err64 = run_cbnmd(tmp_path, deck)
    d32 = tmp_path / "f32"
    d32.mkdir()
    p32, out32, err32 = run_cbnmd(d32, deck, exe=CBNMD + "_f32")
    assert p64.returncode == 0 and p32.returncode == 0, p32.stderr + err32
    r64, r32 = np.array(parse_thermo(out64)), np.array(parse_thermo(out32))
    assert r64.shape == r32.shape and len(r32) == 11
    assert not np.array_equal(r64, r32)                      # it is a different arithmetic
    assert np.abs(r64[:3, 1:] - r32[:3, 1:]).max() < 5e-5    # same physics at the start
    e32 = r32[:, 3]                                          # the thermo columns are T, PE, ETot
    assert np.abs(e32 - e32[0]).max() < 1e-3                 # NVE drift stays small


@pytest.mark.gpu
@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("layout", ["VERLET_2D", "VERLET_CSR"])
def test_cbnmd_matches_oracle_thermo(tmp_path, half, layout):
    import oracle_lib as O

    cells, steps = 12, 60
    args = ["--neigh-type", layout] + (["--force-iteration", "NEIGH_HALF"] if half else [])
    p, out, err = run_cbnmd(tmp_path, DECK.format(c=cells, steps=steps), *args)
    assert p.returncode == 0, p.stderr + err
    rows = parse_thermo(out)
    assert [r[0] for r in rows] == list(range(0, steps + 1, 10))

    ref = O.Sim(mass=[2.0], half=half).create_lattice_fcc(cells=(cells,) * 3).setup()
    ref.record_thermo()
    ref.run(steps, 10)
    want = np.array(ref.thermo())  # step, T, PE/N, KE/N
    got = np.array(rows)
    assert np.array_equal(got[:, 0], want[:, 0])
    # six printed decimals
    assert np.abs(got[:, 1] - want[:, 1]).max() <= 1.1e-6
    assert np.abs(got[:, 2] - want[:, 2]).max() <= 1.1e-6
    assert np.abs(got[:, 3] - (want[:, 2] + want[:, 3])).max() <= 2.1e-6

    # layout of the output file
    n = 4 * cells ** 3
    assert out.startswith("\n#InputFile:\n#====")
    assert "Read input file." in out
    assert "Using: SystemVectorLength: 1 System:1AoSoA" in out
    nb = "Neighbor:CabanaVerletHalf" if half else "Neighbor:CabanaVerletFull"
    assert f"Using: Force:LJCabana {nb} Comm:CabanaMPI Binning:CabanaLinkedCell Integrator:NVE" in out
    assert f"Atoms: {n} {n}" in out and "Created atoms." in out
    assert "\n#Timestep Temperature PotE ETot Time Atomsteps/s \n" in out
    assert re.search(rf"\n1 {n} \| \d+\.\d\d( \d+\.\d\d){{6}} \| PERFORMANCE\n", out)
    assert re.search(rf"\n1 {n} \| 1\.00( \d+\.\d\d){{6}} \| FRACTION\n", out)
    assert re.search(r"#Steps/s Atomsteps/s Atomsteps/\(proc\*s\)\n\d\.\d\de[+-]\d\d \d\.\d\de[+-]\d\d \d\.\d\de[+-]\d\d", out)


@pytest.mark.gpu
def test_cbnmd_inlj_step0_known_answer(tmp_path):
    """in.lj geometry (40^3 cells, 256 000 atoms): T=1.400000 PotE=-6.332812 ETot=-4.232820
    on the perfect lattice (SURVEY.md 8c iii)."""
    p, out, err = run_cbnmd(tmp_path, DECK.format(c=40, steps=0))
    assert p.returncode == 0, p.stderr + err
    rows = parse_thermo(out)
    assert rows[0] == (0, 1.4, -6.332812, -4.232820)
    assert "Atoms: 256000 256000" in out


MULTI_REGION_DECK = """# several regions, several types, per-type masses and temperatures (in.lb-style deck)
units           lj
atom_style      atomic
newton          off
lattice         fcc 0.8442
region          base block 0 8 0 8 0 8
region          core block 2 6 2 6 2 6
region          slab block 0 8 0 8 6 8
create_box      3 box
mass            1 2.0
mass            2 3.5
mass            3 1.25
create_atoms    1 region base
create_atoms    2 region core
create_atoms    3 region slab
velocity        1 create 1.4 87287 loop geom
velocity        2 create 1.4 4711 loop geom
velocity        3 create 1.4 1234 loop geom
pair_style      lj/cut 2.5
pair_coeff      1 1 1.0 1.0 2.5
pair_coeff      1 2 0.9 1.05 2.5
pair_coeff      1 3 1.1 0.95 2.4
pair_coeff      2 2 0.8 1.1 2.5
pair_coeff      2 3 1.0 1.0 2.5
pair_coeff      3 3 1.2 0.9 2.3
neighbor        0.3 bin
neigh_modify    every 20 one 50
comm_modify     cutoff * 20
fix             1 all nve
thermo          10
run             60
"""


@pytest.mark.gpu
def test_multi_region_multi_type_deck(tmp_path):
    """Decks of the shape of input/in.lb: overlapping regions with their own types, masses,
    velocity seeds and a full pair_coeff matrix run through the same kernels
    (inputFile_impl.h:241-303,396-410; force_lj_cabana_neigh_impl.h:70-85,178-183)."""
    import struct

    import oracle_lib as O

    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    p, out, err = run_cbnmd(tmp_path, MULTI_REGION_DECK, "--dumpbinary", "60", str(ref_dir))
    assert p.returncode == 0, p.stderr + err
    assert "Atoms: 2048 2048" in out
    rows = np.array(parse_thermo(out))
    assert list(rows[:, 0]) == list(range(0, 61, 10))
    assert abs(rows[0, 1] - 1.4) < 1e-6              # rescaled to the (common) target temperature
    # NVE: total energy per atom conserved to the integrator's accuracy while T and PE move
    assert np.abs(rows[:, 3] - rows[0, 3]).max() < 5e-3
    assert np.abs(rows[:, 2] - rows[0, 2]).max() > 0.05

    # ---- against the oracle: the deck's initial state (binary dump of step 0) through the oracle's
    # step loop with the deck's masses and pair table; thermo rows and the final forces must agree
    def read_dump(path):
        b = path.read_bytes()
        n = struct.unpack_from("i", b)[0]
        off, d = 4, {}
        for name, dt, cnt in [("id", "i4", n), ("type", "i4", n), ("q", "f8", n), ("x", "f8", 3 * n),
                              ("v", "f8", 3 * n), ("f", "f8", 3 * n)]:
            d[name] = np.frombuffer(b, dtype=dt, count=cnt, offset=off)
            off += d[name].nbytes
        return n, d

    n0, d0 = read_dump(ref_dir / "output.0000000000.000")
    n1, d1 = read_dump(ref_dir / "output.0000000060.000")
    assert n0 == n1 == 2048 and set(np.unique(d0["type"])) == {0, 1, 2}
    eps = np.array([[1.0, 0.9, 1.1], [0.9, 0.8, 1.0], [1.1, 1.0, 1.2]])
    sig = np.array([[1.0, 1.05, 0.95], [1.05, 1.1, 1.0], [0.95, 1.0, 0.9]])
    cut = np.array([[2.5, 2.5, 2.4], [2.5, 2.5, 2.5], [2.4, 2.5, 2.3]])
    tables = (48.0 * eps * sig ** 12, 24.0 * eps * sig ** 6, cut * cut)
    a = (4.0 / 0.8442) ** (1.0 / 3.0)
    ref = O.Sim(ntypes=3, mass=[2.0, 3.5, 1.25], cut=2.5, skin=0.3, tables=tables)
    ref.set_atoms([0.0] * 3, [8 * a] * 3, 1, d0["x"].reshape(-1, 3), d0["v"].reshape(-1, 3), d0["type"], d0["id"])
    ref.setup()
    ref.record_thermo()
    ref.run(60, 10)
    to = np.array(ref.thermo())
    # printed columns: step, T, PE, ETot (6 decimals)
    mine = rows[:, 1:4]
    theirs = np.stack([to[:, 1], to[:, 2], to[:, 2] + to[:, 3]], axis=1)
    assert np.abs(mine - theirs).max() < 2e-6
    b = ref.get()
    fo = b["f"][: b["n_local"]][np.argsort(b["id"][: b["n_local"]])]
    fm = d1["f"].reshape(-1, 3)[np.argsort(d1["id"])]
    assert np.abs(fm - fo).max() <= 1e-9 * np.abs(fo).max()
    # the same deck twice gives the same trace (deterministic full-list path, same rand() stream)
    p2, out2, _ = run_cbnmd(tmp_path, MULTI_REGION_DECK)
    assert [r[:4] for r in parse_thermo(out2)] == [tuple(r) for r in rows.tolist()]
