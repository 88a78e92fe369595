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
