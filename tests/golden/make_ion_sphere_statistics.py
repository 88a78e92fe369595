"""Generates tests/golden/ch4_ion_sphere_statistics.json from TWO runs of oracle/_ref/ref_ch4_ion_sphere (host/ion_sphere.cpp
compiled against the unmodified ch4 reference sources, shipped GS solver): the first run is the pin, the largest relative
difference between the two runs per observable calibrates the statistical tolerance (the reference seeds from random_device).

    oracle/_ref/ref_ch4_ion_sphere 400 GS > a.json ; oracle/_ref/ref_ch4_ion_sphere 400 GS > b.json
    python tests/golden/make_ion_sphere_statistics.py a.json b.json
"""
import json
import os
import sys

import numpy as np

def load(p):
    txt = open(p).read()          # the reference prints "Steady state reached at time step N" before the driver's JSON
    return json.loads(txt[txt.index("{"):])


a, b = (load(p) for p in sys.argv[1:3])
spread = {}
for k, v in a.items():
    if isinstance(v, list):
        x, y = np.array(v), np.array(b[k])
        spread[k] = float(np.abs(x - y).max() / np.abs(x).max())
out = {"run": a, "second_run_scalars": {k: v for k, v in b.items() if not isinstance(v, list)}, "spread_between_two_reference_runs": spread}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ch4_ion_sphere_statistics.json")
json.dump(out, open(dst, "w"), indent=1)
print("wrote", dst, spread)
