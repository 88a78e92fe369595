"""CPU suite: the C-ABI shared library loads and exports every symbol include/espic.h declares
(no compute calls: there is no GPU here), and refuses to run without a device instead of falling back."""
import ctypes as C
import os
import re
import numpy as np
import pytest

import statefile as sf
from engines import _espic

HEADER = os.path.join(sf.ROOT, "include", "espic.h")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(espic_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    es = _espic()
    lib = es.load()
    names = declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libespic_cuda.so does not export %s" % n
    assert sorted(es.EXPORTS) == names, "espic.py binding list and include/espic.h disagree"


def test_no_cpu_fallback():
    """Without a CUDA device espic_create must fail with an error text, never compute on the host."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("CUDA device present")
    except ImportError:
        pass
    es = _espic()
    with pytest.raises(es.EspicError) as ei:
        es.Engine(5, 5, 5, (0, 0, 0), (1, 1, 1))
    assert "no CUDA device" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (SURVEY 8c rule: oracle is test infrastructure only)."""
    pkg = os.path.join(sf.ROOT, "plasma-simulations-by-example_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in txt and "espic_oracle" not in txt and "from oracle" not in txt, os.path.join(root, f)


def test_multigrid_plan_host_logic(monkeypatch):
    """espic_mg_plan (host-only planning of the multigrid preconditioner): semi-coarsening on the reference's dz = 2 dx meshes,
    the coarsest-level size bound, and which levels the slab-decomposed solver runs redundantly on every rank."""
    es = _espic()
    monkeypatch.delenv("ESPIC_MG_SLAB_REDUNDANT_NODES", raising=False)

    def dh(n):
        return (0.2 / (n[0] - 1), 0.2 / (n[1] - 1), 0.4 / (n[2] - 1))
    # bench mesh (BASELINE configs[2..3]): z is not coarsened on the first level
    dims, fr, unit = es.mg_plan(128, 128, 128, dh((128, 128, 128)))
    assert dims == [(128, 128, 128), (64, 64, 128), (32, 32, 64), (16, 16, 32), (8, 8, 16)]
    assert fr == 4 and unit == 8                      # one rank: nothing is redundant but the coarsest level by construction
    # configs[4] on 8 ranks: levels of <= 65536 nodes are solved by every rank in full, 256 planes split 8 x 32 (unit 16)
    dims, fr, unit = es.mg_plan(256, 256, 256, dh((256, 256, 256)), nranks=8)
    assert dims[:4] == [(256, 256, 256), (128, 128, 256), (64, 64, 128), (32, 32, 64)] and len(dims) == 6
    assert fr == 3 and unit == 16 and 256 % (8 * unit) == 0
    monkeypatch.setenv("ESPIC_MG_SLAB_REDUNDANT_NODES", "0")
    assert es.mg_plan(256, 256, 256, dh((256, 256, 256)), nranks=8)[1] == 5
    monkeypatch.setenv("ESPIC_MG_SLAB_REDUNDANT_NODES", "100000000")
    assert es.mg_plan(256, 256, 256, dh((256, 256, 256)), nranks=8)[1] == 1        # level 0 is never redundant
    monkeypatch.delenv("ESPIC_MG_SLAB_REDUNDANT_NODES")
    # the shipped ch3 mesh (21 x 21 x 41, isotropic spacing): full coarsening, two levels
    dims, fr, unit = es.mg_plan(21, 21, 41, (0.01, 0.01, 0.01), nranks=2)
    assert dims == [(21, 21, 41), (11, 11, 21)] and fr == 1 and unit == 2
    # a mesh below the coarsest-level bound has no hierarchy (the solver falls back to the Jacobi preconditioner)
    assert es.mg_plan(9, 9, 9, (1.0, 1.0, 1.0))[0] == [(9, 9, 9)]
    with pytest.raises(es.EspicError):
        es.mg_plan(1, 9, 9, (1.0, 1.0, 1.0))
