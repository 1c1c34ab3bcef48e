"""File formats either side of the MD path: LAMMPS data files
(`read_data` / `write_data`, reference src/read_data.h), VTK particle dumps
(src/vtk_writer.h:205-294) and the binary dump / correctness check
(src/cabanamd_impl.h:434-642).

CPU tests drive the host-only `cbmd_io_tool` (same templates the cbnMD driver uses, over
a plain host particle store) and compare with independent Python restatements of the
formats.  GPU tests run the real driver: write_data -> read_data restart continues the
trajectory, dumps land next to the run, --dumpbinary / --correctness close the loop."""
import os
import re
import shutil
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cabanamd_b200", "lib")
TOOL = os.path.join(LIB, "cbmd_io_tool")
CBNMD = os.path.join(LIB, "cbnMD")


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cabanamd_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "cabanamd_b200", "host")])
    assert os.path.exists(TOOL)


def tool(*args, cwd=None):
    return subprocess.run([TOOL, *map(str, args)], capture_output=True, text=True, cwd=cwd, timeout=120)


def g(v):
    """what `stream << double` (precision 6) and printf("%g") print"""
    return "%g" % v


def make_state(n, seed=0, box=(0.0, 7.5), ntypes=2):
    rng = np.random.default_rng(seed)
    x = rng.uniform(box[0], box[1], (n, 3))
    v = rng.normal(0, 1.3, (n, 3))
    # values that stress %g: tiny, huge, negative zero-ish, integers
    v[0] = [1e-7, -123456789.0, 0.0]
    x[1] = [box[0], 1.0, 2.5]
    ids = rng.permutation(n) + 1
    typ = rng.integers(1, ntypes + 1, n)
    return ids, typ, x, v


def write_data_file(path, ids, typ, x, v, box, ntypes, masses=None, charge=None, shuffle_vel=False,
                    pair_coeffs=False, digits=17):
    f = lambda a: repr(float(a)) if digits == 17 else ("%.*g" % (digits, a))
    L = ["LAMMPS data file written by the test-suite", "", f"{len(ids)} atoms   # comment", f"{ntypes} atom types",
         "", f"{f(box[0])} {f(box[1])} xlo xhi", f"{f(box[0])} {f(box[1])} ylo yhi", f"{f(box[0])} {f(box[1])} zlo zhi",
         ""]
    if masses is not None:
        L += ["Masses", ""] + [f"{t + 1} {m}" for t, m in enumerate(masses)] + [""]
    if pair_coeffs:
        L += ["Pair Coeffs # lj/cut", ""] + [f"{t + 1} 1.0 1.0" for t in range(ntypes)] + [""]
    L += ["Atoms # atomic" if charge is None else "Atoms # charge", ""]
    for k in range(len(ids)):
        q = "" if charge is None else f" {f(charge[k])}"
        L.append(f"{ids[k]} {typ[k]}{q} {f(x[k, 0])} {f(x[k, 1])} {f(x[k, 2])}")
    L += ["", "Velocities", ""]
    order = np.random.default_rng(5).permutation(len(ids)) if shuffle_vel else range(len(ids))
    for k in order:
        L.append(f"{ids[k]} {f(v[k, 0])} {f(v[k, 1])} {f(v[k, 2])}")
    path.write_text("\n".join(L) + "\n")


def expected_write_data(ids, typ, x, v, box, ntypes, fmt=g):
    L = ["LAMMPS data file from CabanaMD", "", f"{len(ids)} atoms", f"{ntypes} atom types", ""]
    L += [f"{fmt(box[0])} {fmt(box[1])} {a}lo {a}hi" for a in "xyz"]
    L += ["", "Atoms # atomic", ""]
    L += [f"{ids[k]} {typ[k]} {fmt(x[k, 0])} {fmt(x[k, 1])} {fmt(x[k, 2])}" for k in range(len(ids))]
    L += ["", "Velocities", ""]
    L += [f"{ids[k]} {fmt(v[k, 0])} {fmt(v[k, 1])} {fmt(v[k, 2])}" for k in range(len(ids))]
    return "\n".join(L) + "\n"


def parse_data_file(text):
    """minimal reader for the files write_data produces"""
    lines = text.splitlines()
    n = int(lines[2].split()[0])
    a0 = lines.index("Atoms # atomic") + 2
    v0 = lines.index("Velocities") + 2
    A = np.array([ln.split() for ln in lines[a0:a0 + n]], dtype=float)
    V = np.array([ln.split() for ln in lines[v0:v0 + n]], dtype=float)
    return A[:, 0].astype(int), A[:, 1].astype(int), A[:, 2:5], V[:, 0].astype(int), V[:, 1:4]


# ------------------------------------------------------------------ data files (CPU)
def test_write_data_matches_reference_format(tmp_path):
    box, ntypes = (0.0, 7.5), 2
    ids, typ, x, v = make_state(200)
    write_data_file(tmp_path / "in.data", ids, typ, x, v, box, ntypes, masses=[2.0, 8.5], shuffle_vel=True,
                    pair_coeffs=True)
    p = tool("data", tmp_path / "in.data", tmp_path / "out.data")
    assert p.returncode == 0, p.stderr
    assert p.stdout.strip() == "200 200 2 | 2 8.5"
    assert "Ignoring potential parameters in data file" in p.stderr
    # one rank: the reference's text byte for byte (6 significant digits, file order kept,
    # velocities re-associated by id although the section was shuffled)
    assert (tmp_path / "out.data").read_text() == expected_write_data(ids, typ, x, v, box, ntypes)


def test_data_roundtrip_is_exact_at_precision_17(tmp_path):
    box, ntypes = (-3.25, 9.0), 3
    ids, typ, x, v = make_state(333, seed=3, box=box, ntypes=ntypes)
    write_data_file(tmp_path / "a.data", ids, typ, x, v, box, ntypes)
    assert tool("data", tmp_path / "a.data", tmp_path / "b.data", 17).returncode == 0
    assert tool("data", tmp_path / "b.data", tmp_path / "c.data", 17).returncode == 0
    assert (tmp_path / "b.data").read_text() == (tmp_path / "c.data").read_text()
    i2, t2, x2, iv, v2 = parse_data_file((tmp_path / "c.data").read_text())
    assert np.array_equal(i2, ids) and np.array_equal(t2, typ) and np.array_equal(iv, ids)
    assert np.array_equal(x2, x) and np.array_equal(v2, v)  # bit-exact through two text passes


def test_read_data_keeps_only_the_ranks_own_atoms(tmp_path):
    box = (0.0, 8.0)
    ids, typ, x, v = make_state(500, seed=7, box=box)
    x[3] = [4.0, 1.0, 1.0]  # exactly on the upper face of the sub-box: belongs to the neighbour
    x[4] = [0.0, 1.0, 1.0]  # exactly on the lower face: owned
    write_data_file(tmp_path / "in.data", ids, typ, x, v, box, 2)
    p = tool("data", tmp_path / "in.data", tmp_path / "out.data", 17, "atomic", 0, 4, 0, 8, 0, 8)
    assert p.returncode == 0, p.stderr
    mine = x[:, 0] < 4.0
    assert not mine[3] and mine[4]
    assert p.stdout.split("|")[0].split() == ["500", str(int(mine.sum())), "2"]
    i2, t2, x2, iv, v2 = parse_data_file_subset((tmp_path / "out.data").read_text(), int(mine.sum()))
    assert np.array_equal(i2, ids[mine]) and np.array_equal(x2, x[mine]) and np.array_equal(v2, v[mine])


def parse_data_file_subset(text, n):
    lines = text.splitlines()
    a0 = lines.index("Atoms # atomic") + 2
    v0 = lines.index("Velocities") + 2
    A = np.array([ln.split() for ln in lines[a0:a0 + n]], dtype=float)
    V = np.array([ln.split() for ln in lines[v0:v0 + n]], dtype=float)
    return A[:, 0].astype(int), A[:, 1].astype(int), A[:, 2:5], V[:, 0].astype(int), V[:, 1:4]


def test_read_data_charge_style_and_missing_velocities(tmp_path):
    box = (0.0, 5.0)
    ids, typ, x, v = make_state(40, seed=9, box=box, ntypes=1)
    q = np.linspace(-1, 1, 40)
    write_data_file(tmp_path / "in.data", ids, typ, x, v, box, 1, charge=q)
    text = (tmp_path / "in.data").read_text()
    (tmp_path / "nov.data").write_text(text[: text.index("Velocities")])
    p = tool("data", tmp_path / "nov.data", tmp_path / "out.data", 17, "charge")
    assert p.returncode == 0, p.stderr
    i2, t2, x2, iv, v2 = parse_data_file((tmp_path / "out.data").read_text())
    assert np.array_equal(x2, x) and np.array_equal(i2, ids)
    assert not v2.any()  # atoms without a velocity line start at rest


@pytest.mark.parametrize("mutate,msg", [
    (lambda t: t.replace("Atoms # atomic", "Bonds"), "Unknown data file keyword: Bonds"),
    (lambda t: t.replace("Velocities", "Xelocities").replace("Atoms # atomic", "Velocities", 1)
     .replace("Xelocities", "Atoms"), "Must read Atoms before Velocities"),
    (lambda t: t[: t.index("zlo zhi") - 20], "header ended before the 'zlo zhi' line"),
    (lambda t: "", "Could not read from data file"),
])
def test_read_data_errors(tmp_path, mutate, msg):
    ids, typ, x, v = make_state(20, seed=2)
    write_data_file(tmp_path / "ok.data", ids, typ, x, v, (0.0, 7.5), 2)
    (tmp_path / "bad.data").write_text(mutate((tmp_path / "ok.data").read_text()))
    p = tool("data", tmp_path / "bad.data", tmp_path / "out.data")
    assert p.returncode != 0 and msg in p.stderr


# ------------------------------------------------------------------ VTK (CPU)
def expected_vtu(ids, typ0, x, v):
    n = len(ids)
    arr = lambda ty, name, nc: f'\t\t<DataArray type="{ty}" Name="{name}" NumberOfComponents="{nc}" format="ascii">\n'
    end = "\n\t\t</DataArray>\n"
    s = '<?xml version="1.0"?>\n'
    s += '<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian" header_type="UInt32">\n'
    s += "<UnstructuredGrid>\n"
    s += f'<Piece NumberOfPoints="{n}" NumberOfCells="0">\n'
    s += "\t<PointData>\n"
    s += arr("Float64", "Velocity", 3) + "".join(f"{g(a)} {g(b)} {g(c)} " for a, b, c in v) + end
    s += arr("Int32", "Id", 1) + "".join(f"{i} " for i in ids) + end
    s += arr("Int32", "Type", 1) + "".join(f"{t} " for t in typ0) + end
    s += "\t</PointData>\n\t<CellData>\n\t</CellData>\n\t<Points>\n"
    s += arr("Float64", "Points", 3) + "".join(f"{g(a)} {g(b)} {g(c)} " for a, b, c in x) + end
    s += "\t</Points>\n\t<Cells>\n"
    s += arr("Int32", "connectivity", 1) + end + arr("Int32", "offsets", 1) + end + arr("UInt8", "types", 1) + end
    s += "\t</Cells>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n"
    return s


@pytest.mark.parametrize("n,workers", [(0, 1), (1, 3), (9000, 1), (50_000, 8)])
def test_vtk_particle_dump_matches_reference_format(tmp_path, n, workers):
    box = (0.0, 30.0)
    ids, typ, x, v = make_state(max(n, 2), seed=n)
    ids, typ, x, v = ids[:n], typ[:n], x[:n], v[:n]
    if n > 10:
        # %g corner cases: exponent forms, rounding to 6 digits, negative zero
        v[2] = [9.9999995, 999999.5, -0.0]
        v[3] = [1e-5, 1.00000049e-4, 123456.5]
        v[4] = [1e15, -1e-300, 0.1 + 0.2]
    write_data_file(tmp_path / "in.data", ids, typ, x, v, box, 2)
    p = tool("vtk", tmp_path / "in.data", "dump%_*.vtu", 40, 0, 1, workers, cwd=tmp_path)
    assert p.returncode == 0, p.stderr
    want = expected_vtu(ids, typ - 1, x, v)
    assert (tmp_path / "dump_0_0040.vtu").read_text() == want
    assert (tmp_path / "dump_0_0041.vtu").read_text() == want
    assert not list(tmp_path.glob("*.pvtu"))  # the index is rank 1's job (vtk_writer.h:264)


def test_vtk_parallel_index_written_by_rank_1(tmp_path):
    ids, typ, x, v = make_state(10)
    write_data_file(tmp_path / "in.data", ids, typ, x, v, (0.0, 7.5), 2)
    p = tool("vtk", tmp_path / "in.data", "out/d%_*.vtu".replace("out/", ""), 7, 1, 4, cwd=tmp_path)
    assert p.returncode == 0, p.stderr
    idx = (tmp_path / "d_0007.pvtu").read_text()
    assert idx.startswith('<?xml version="1.0"?>\n<VTKFile type="PUnstructuredGrid" version="0.1" '
                          'byte_order="LittleEndian" header_type="UInt32">\n<PUnstructuredGrid>\n\t<PPointData>\n')
    assert '\t\t<PDataArray type="Float64" Name="Velocity"/>\n\t\t<PDataArray type="Int32" Name="Id"/>\n' in idx
    assert [m for m in re.findall(r'<Piece Source="([^"]+)"/>', idx)] == [f"d_{r}_0007.vtu" for r in range(4)]
    assert idx.endswith("</PUnstructuredGrid>\n</VTKFile>\n")
    assert (tmp_path / "d_1_0007.vtu").exists()


def test_vtk_pattern_errors(tmp_path):
    ids, typ, x, v = make_state(4)
    write_data_file(tmp_path / "in.data", ids, typ, x, v, (0.0, 7.5), 2)
    p = tool("vtk", tmp_path / "in.data", "dump_%.vtu", 1, 0, 1, cwd=tmp_path)
    assert p.returncode != 0 and "does not contain required '*'" in p.stderr
    p = tool("vtk", tmp_path / "in.data", "dump_*.vtu", 1, 0, 1, cwd=tmp_path)
    assert p.returncode != 0 and "does not contain required '%'" in p.stderr


# ------------------------------------------------------------------ binary dump (CPU)
def read_dump(path):
    b = path.read_bytes()
    n = struct.unpack_from("i", b)[0]
    off = 4
    out = {}
    for name, dt, cnt in [("id", "i4", n), ("type", "i4", n), ("q", "f8", n), ("x", "f8", 3 * n), ("v", "f8", 3 * n),
                          ("f", "f8", 3 * n)]:
        out[name] = np.frombuffer(b, dtype=dt, count=cnt, offset=off)
        off += out[name].nbytes
    assert off == len(b)
    return n, out


def test_binary_dump_layout_and_correctness_report(tmp_path):
    box = (0.0, 7.5)
    ids, typ, x, v = make_state(300, seed=11)
    write_data_file(tmp_path / "a.data", ids, typ, x, v, box, 2)
    (tmp_path / "ref").mkdir()
    assert tool("dump", tmp_path / "a.data", tmp_path / "ref", 20, 3).returncode == 0
    n, d = read_dump(tmp_path / "ref" / "output.0000000020.003")
    assert n == 300 and np.array_equal(d["id"], ids) and np.array_equal(d["type"], typ - 1)
    assert np.array_equal(d["x"].reshape(-1, 3), x) and np.array_equal(d["v"].reshape(-1, 3), v)
    assert np.array_equal(d["f"].reshape(-1, 3), -x) and not d["q"].any()

    # same atoms in another order with known perturbations: matched by id
    perm = np.random.default_rng(1).permutation(300)
    x2, v2 = x[perm].copy(), v[perm].copy()
    x2[5, 1] += 0.25
    x2[9, 2] -= 0.5
    v2[7, 0] += 2.0
    write_data_file(tmp_path / "b.data", ids[perm], typ[perm], x2, v2, box, 2)
    rep = tmp_path / "corr.txt"
    shutil.copy(tmp_path / "ref" / "output.0000000020.003", tmp_path / "ref" / "output.0000000000.003")
    for step in (0, 20):
        assert tool("check", tmp_path / "b.data", tmp_path / "ref", step, 3, rep).returncode == 0
    dr = np.sqrt(0.25 ** 2 + 0.5 ** 2)
    line = f"{g(dr)} 0.5 2 2 {g(dr)} 0.5"  # f = -x in the tool, so |df| = |dr|
    assert rep.read_text() == ("# timestep deltarnorm maxdelr deltavnorm maxdelv deltafnorm maxdelf\n"
                               f"0 {line}\n20 {line}\n")

    # atom-count mismatch and missing files are reported, not ignored
    write_data_file(tmp_path / "c.data", ids[:299], typ[:299], x[:299], v[:299], box, 2)
    assert tool("check", tmp_path / "c.data", tmp_path / "ref", 20, 3, rep).returncode == 10 + 2
    assert tool("check", tmp_path / "b.data", tmp_path / "ref", 40, 3, rep).returncode == 10 + 1


# ------------------------------------------------------------------ the real driver (GPU)
DECK = """units lj
atom_style atomic
newton off
lattice fcc 0.8442
region box block 0 {c} 0 {c} 0 {c}
create_box 1 box
create_atoms 1 box
mass 1 2.0
velocity all create 1.4 87287 loop geom
pair_style lj/cut 2.5
pair_coeff 1 1 1.0 1.0 2.5
neighbor 0.3 bin
neigh_modify every 20 one 50
comm_modify cutoff * 20
fix 1 all nve
thermo 10
{extra}
run {steps}
"""

RESTART = """units lj
atom_style atomic
newton off
read_data {data}
pair_style lj/cut 2.5
pair_coeff 1 1 1.0 1.0 2.5
neighbor 0.3 bin
neigh_modify every 20 one 50
comm_modify cutoff * 20
fix 1 all nve
thermo 10
run {steps}
"""

THERMO = re.compile(r"^(\d+)\t(-?\d+\.\d{6})\t(-?\d+\.\d{6})\t(-?\d+\.\d{6})\t")


def cbnmd(tmp_path, deck, *args, name="md"):
    f = tmp_path / f"{name}.deck"
    f.write_text(deck)
    out, err = tmp_path / f"{name}.out", tmp_path / f"{name}.err"
    p = subprocess.run([CBNMD, "-il", str(f), "-o", str(out), "-e", str(err), *map(str, args)],
                       capture_output=True, text=True, cwd=tmp_path, timeout=600)
    rows = [tuple(float(v) for v in m.groups()) for m in map(THERMO.match, out.read_text().splitlines()) if m] \
        if out.exists() else []
    return p, rows, (err.read_text() if err.exists() else "")


@pytest.mark.gpu
def test_write_data_then_read_data_continues_the_run(tmp_path):
    """40 steps + write_data (17 digits) + read_data + 40 steps == 80 steps in one go.
    Step 40 is a neighbour-rebuild step, so both runs rebuild their lists from the same
    positions.  write_data stores no masses (as in the reference), so the restart deck
    keeps its `mass` line."""
    c = 8
    p, full, err = cbnmd(tmp_path, DECK.format(c=c, steps=80, extra=""), name="full")
    assert p.returncode == 0, p.stderr + err
    p, first, err = cbnmd(tmp_path, DECK.format(c=c, steps=40, extra="write_data mid.data precision 17"),
                          name="first")
    assert p.returncode == 0, p.stderr + err
    text = (tmp_path / "mid.data").read_text()
    assert text.startswith("LAMMPS data file from CabanaMD\n\n2048 atoms\n1 atom types\n\n0 ")
    deck = RESTART.format(data="mid.data", steps=40).replace("read_data mid.data", "read_data mid.data\nmass 1 2.0")
    p, second, err = cbnmd(tmp_path, deck, name="second")
    assert p.returncode == 0, p.stderr + err
    assert "Atoms: 2048 2048" in (tmp_path / "second.out").read_text()
    assert first == full[:5]
    # restart rows 0..40 are the original rows 40..80 (T, PE, ETot to the printed digits;
    # summation order differs after the re-sort, hence one unit in the last place)
    a, b = np.array(second), np.array(full[4:])
    assert a.shape == b.shape == (5, 4)
    assert np.abs(a[:, 1:] - b[:, 1:]).max() <= 2.1e-6


@pytest.mark.gpu
def test_default_write_data_is_the_reference_text(tmp_path):
    p, rows, err = cbnmd(tmp_path, DECK.format(c=4, steps=0, extra="write_data lattice.data"))
    assert p.returncode == 0, p.stderr + err
    ids, typ, x, iv, v = parse_data_file((tmp_path / "lattice.data").read_text())
    assert len(ids) == 256 and set(typ) == {1} and np.array_equal(np.sort(ids), np.arange(1, 257))
    a = (4 / 0.8442) ** (1 / 3)
    # six significant digits of the fcc sites
    frac = x / a * 2
    assert np.abs(frac - np.round(frac)).max() < 2e-5
    assert (tmp_path / "lattice.data").read_text().splitlines()[5] == f"0 {g(4 * a)} xlo xhi"


@pytest.mark.gpu
def test_vtk_dump_from_the_step_loop(tmp_path):
    p, rows, err = cbnmd(tmp_path, DECK.format(c=6, steps=20, extra="dump dmpvtk all vtk 10 dump%_*.vtu"))
    assert p.returncode == 0, p.stderr + err
    files = sorted(f.name for f in tmp_path.glob("dump_*.vtu"))
    assert files == ["dump_0_0010.vtu", "dump_0_0020.vtu"]
    t = (tmp_path / "dump_0_0020.vtu").read_text()
    assert '<Piece NumberOfPoints="864" NumberOfCells="0">' in t
    body = t.split('Name="Id" NumberOfComponents="1" format="ascii">\n')[1].split("\n")[0]
    assert sorted(map(int, body.split())) == list(range(1, 865))
    pts = t.split('Name="Points" NumberOfComponents="3" format="ascii">\n')[1].split("\n")[0]
    assert len(pts.split()) == 3 * 864


@pytest.mark.gpu
def test_dumpbinary_then_correctness(tmp_path):
    (tmp_path / "ref").mkdir()
    deck = DECK.format(c=6, steps=40, extra="")
    p, rows, err = cbnmd(tmp_path, deck, "--dumpbinary", 20, tmp_path / "ref", name="a")
    assert p.returncode == 0, p.stderr + err
    assert sorted(f.name for f in (tmp_path / "ref").iterdir()) == [f"output.{s:010d}.000" for s in (0, 20, 40)]
    n, d = read_dump(tmp_path / "ref" / "output.0000000040.000")
    assert n == 864 and np.array_equal(np.sort(d["id"]), np.arange(1, 865))
    assert np.abs(d["f"].reshape(-1, 3).sum(0)).max() < 1e-9  # Newton 3
    # the same run checked against its own dumps: deterministic kernels -> all zeros
    p, rows2, err = cbnmd(tmp_path, deck, "--correctness", 20, tmp_path / "ref", tmp_path / "corr.txt", name="b")
    assert p.returncode == 0, p.stderr + err
    lines = (tmp_path / "corr.txt").read_text().splitlines()
    assert lines[0] == "# timestep deltarnorm maxdelr deltavnorm maxdelv deltafnorm maxdelf"
    assert lines[1:] == ["0 0 0 0 0 0 0", "20 0 0 0 0 0 0", "40 0 0 0 0 0 0"]
    # a half-list run differs from the full-list reference by round-off only
    p, rows3, err = cbnmd(tmp_path, deck, "--force-iteration", "NEIGH_HALF", "--correctness", 20, tmp_path / "ref",
                          tmp_path / "corr_half.txt", name="c")
    assert p.returncode == 0, p.stderr + err
    vals = np.array([[float(v) for v in ln.split()] for ln in (tmp_path / "corr_half.txt").read_text().splitlines()[1:]])
    assert vals.shape == (3, 7) and vals[:, 1:].max() < 1e-8 and vals[2, 5] > 0


def test_write_data_from_several_ranks_is_one_complete_file(tmp_path):
    """Multi-rank write_data: every rank formats its own atoms, ranks > 0 hand theirs over
    through part files, rank 0 writes ONE file with the global box and all atoms (the
    reference writes rank 0's sub-box and atoms only, read_data.h:385-421).  Played here as
    three sequential single-rank calls of the same code."""
    box = (0.0, 9.0)
    ids, typ, x, v = make_state(600, seed=21, box=box)
    write_data_file(tmp_path / "in.data", ids, typ, x, v, box, 2)
    slabs = [(0.0, 3.0), (3.0, 6.0), (6.0, 9.0)]          # 3 ranks along x
    for rank in (2, 1, 0):                                 # rank 0 last: it assembles
        lo, hi = slabs[rank]
        p = tool("data", tmp_path / "in.data", tmp_path / "out.data", 17, "atomic", lo, hi, 0, 9, 0, 9, rank, 3)
        assert p.returncode == 0, p.stderr
    assert not list(tmp_path.glob("out.data.part*"))       # part files are consumed
    text = (tmp_path / "out.data").read_text()
    assert text.splitlines()[2] == "600 atoms"
    assert text.splitlines()[5] == "0 9 xlo xhi"            # the GLOBAL box
    i2, t2, x2, iv, v2 = parse_data_file(text)
    order = np.concatenate([np.flatnonzero((x[:, 0] >= lo) & (x[:, 0] < hi)) for lo, hi in slabs])
    assert len(order) == 600
    assert np.array_equal(i2, ids[order]) and np.array_equal(iv, ids[order])   # rank order, file order inside
    assert np.array_equal(x2, x[order]) and np.array_equal(v2, v[order]) and np.array_equal(t2, typ[order])
