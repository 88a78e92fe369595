"""CPU suite: the oracle restatement replays every committed golden fixture (tests/golden/*.npz, produced by
the unmodified reference via tests/golden/make_golden.py) and must reproduce the reference's output bit for bit:
particles (order included), densities, rho, phi, ef and the diagnostics."""
import glob
import os
import numpy as np
import pytest

import statefile as sf
from engines import OracleEngine

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def load(path):
    d = np.load(path)
    return d, [str(c) for c in d["cmds"]], str(d["which"]), sf.state_from_dict(d, "in_"), sf.state_from_dict(d, "out_")


def test_fixtures_present():
    assert len(GOLDEN) >= 12


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_reference(path):
    _, cmds, which, st_in, ref = load(path)
    got = OracleEngine(st_in, box=(which == "ref_ch2")).run(cmds)
    for name in ("phi", "rho", "ef", "node_vol"):
        assert np.array_equal(bits(getattr(got, name)), bits(getattr(ref, name))), name
    assert np.array_equal(got.object_id, ref.object_id)
    assert len(got.species) == len(ref.species)
    for a, b in zip(got.species, ref.species):
        assert a["part"].shape == b["part"].shape
        for k in ("part", "den", "den_ave"):
            assert np.array_equal(bits(a[k]), bits(b[k])), k
    assert np.array_equal(bits(got.diag), bits(ref.diag)), (got.diag, ref.diag)
