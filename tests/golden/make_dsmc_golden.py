"""Generates tests/golden/ch4/dsmc.npz with the compiled, unmodified ch4 reference (oracle/_ref/ref_ch4_dsmc): DSMC_MEX::apply
three times on the case of tests/test_dsmc.py, then Species::computeMPC."""
import os
import sys
import tempfile
import pathlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_dsmc as td   # noqa: E402

w, part = td.make_case()
with tempfile.TemporaryDirectory() as t:
    out, sig, mpc = td.run_reference(w, part, 4242, 3, pathlib.Path(t))
np.savez_compressed(os.path.join(HERE, "ch4", "dsmc.npz"), seed=4242, reps=3, part=out, sigma_cr_max=sig, mpc=mpc)
print("wrote dsmc.npz", out.shape, sig)
