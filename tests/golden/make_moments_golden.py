"""Generates tests/golden/moments.npz with the compiled, unmodified ch4 reference (oracle/_ref/ref_ch4_moments)."""
import os
import sys
import tempfile
import pathlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_moments as tm   # noqa: E402

w, sp = tm.make_case(seed=11, dims=(10, 9, 12), n=8000)
with tempfile.TemporaryDirectory() as t:
    ref = tm.run_reference(w, sp, 2, pathlib.Path(t))
np.savez_compressed(os.path.join(HERE, "ch4", "moments.npz"), ni=w.ni, nj=w.nj, nk=w.nk, x0=w.x0, xm=w.xm, mass=sp.mass, reps=2,
                    part=sp.particles(), **ref)
print("wrote moments.npz", {k: float(np.abs(v).max()) for k, v in ref.items()})
