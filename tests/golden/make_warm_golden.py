"""Generates tests/golden/ch4/warm_source.npz with the compiled, unmodified ch4 reference (oracle/_ref/ref_ch4_warm)."""
import os
import sys
import tempfile
import pathlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_warm_source as tw   # noqa: E402

w = tw.make_world(5)
with tempfile.TemporaryDirectory() as t:
    part = tw.run_reference(w, 2024, 2, pathlib.Path(t))
np.savez_compressed(os.path.join(HERE, "ch4", "warm_source.npz"), world_seed=5, seed=2024, reps=2, ef=w.ef, part=part)
print("wrote warm_source.npz", part.shape)
