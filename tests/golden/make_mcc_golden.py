"""Generates tests/golden/ch4/mcc.npz with the compiled, unmodified ch4 reference (oracle/_ref/ref_ch4_mcc): MCC_CEX::apply twice
on the case of tests/test_mcc.py."""
import os
import sys
import tempfile
import pathlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_mcc as tm   # noqa: E402

w, part, den, vel, T = tm.make_case()
with tempfile.TemporaryDirectory() as t:
    out = tm.run_reference(w, part, den, vel, T, 777, 2, pathlib.Path(t))
np.savez_compressed(os.path.join(HERE, "ch4", "mcc.npz"), seed=777, reps=2, part=out)
print("wrote mcc.npz", out.shape, int(np.all(out[3:6] == 0, axis=0).sum()), "stopped")
