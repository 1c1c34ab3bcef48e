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
