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


@pytest.mark.parametrize("group", [1, 8])
@pytest.mark.parametrize("rows", [1, 7, 8, 50, 97])
def test_verlet_table_addressing(group, rows):
    """The table layouts the build, force, energy and CSR-export kernels share (nb_entry in
    cbmd_internal.cuh): every (atom, entry) has its own slot inside the table, a 32-atom tile
    is one contiguous block, and a warp reads its index stream as whole 128-byte lines."""
    import ctypes as C

    import numpy as np

    L = cb.load_library()
    L.cbmd_table_offset.restype = C.c_int64
    L.cbmd_table_offset.argtypes = [C.c_int] * 4
    L.cbmd_table_size.restype = C.c_int64
    L.cbmd_table_size.argtypes = [C.c_int] * 3
    n_atoms = 75  # rounds up to 96 = 3 tiles
    size = L.cbmd_table_size(group, n_atoms, rows)
    off = np.array([[L.cbmd_table_offset(group, i, n, rows) for n in range(rows)] for i in range(96)])
    assert off.min() >= 0 and off.max() < size
    assert len(np.unique(off)) == off.size                       # no two entries share a slot
    block = size // 3
    for t in range(3):                                           # tile = contiguous block
        o = off[32 * t:32 * t + 32]
        assert o.min() >= t * block and o.max() < (t + 1) * block
    if group == 1:
        # row n of a tile is one 128-byte line: lane = atom & 31
        for n in range(rows):
            assert np.array_equal(off[:32, n] - off[0, n], np.arange(32))
            assert off[0, n] % 32 == 0
    else:
        # chunk r of a quad (4 atoms x 8 entries) is one 128-byte line: lane = (atom&3)*8 + (n&7)
        for q in range(0, 96, 4):
            for r in range((rows + 7) // 8):
                base = off[q, 8 * r] if 8 * r < rows else None
                for a in range(4):
                    for g in range(8):
                        n = 8 * r + g
                        if n < rows:
                            assert off[q + a, n] - base == a * 8 + g
                assert base % 32 == 0
