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
