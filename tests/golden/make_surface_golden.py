"""Generates tests/golden/ch4/surface.npz with the compiled, unmodified ch4 reference (oracle/_ref/ref_ch4_surface):
Species::advance(neutrals, spherium) on the three cases of tests/test_surface.py."""
import os
import sys
import tempfile
import pathlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_surface as ts   # noqa: E402
import surface_cases as sc  # noqa: E402

w = sc.make_world()
part, pdt = sc.make_particles(w, 17, 4000, mpw=5.0)
out = dict(seed=2024, reps=2, ef=w.ef)
with tempfile.TemporaryDirectory() as t:
    for name, charge, same in ts.CASES:
        adv, neut, sput = sc.species_triplet(w, charge)
        ref = ts.run_reference(w, adv, adv if charge == 0 else neut, sput, part, pdt, 2024, 2, same, pathlib.Path(t))
        for s in range(3):
            out[f"{name}_{s}"] = ref[s]
        print(name, [r.shape for r in ref])
np.savez_compressed(os.path.join(HERE, "ch4", "surface.npz"), **out)
