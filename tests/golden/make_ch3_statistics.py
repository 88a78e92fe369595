"""Generates tests/golden/ch3_sphere_statistics.json from a run of the UNMODIFIED reference ch3/ver2 program
(g++ -O2 /root/reference/ch3/ver2/*.cpp; mkdir results; ./a.out): the steady-state observables the north-star's second
parity check names.  The reference seeds its RNG from std::random_device, so these are statistical pins.

    python tests/golden/make_ch3_statistics.py <run directory with runtime_diags.csv and results/fields_00400.vti>
"""
import json
import os
import re
import sys

import numpy as np


def vti_array(path, name, ni, nj, nk):
    txt = open(path).read()
    m = re.search(r'<DataArray Name="%s"[^>]*>\n(.*?)</DataArray>' % re.escape(name), txt, re.S)
    a = np.array(m.group(1).split(), dtype=np.float64)
    return a.reshape(nk, nj, ni)          # VTK order: i fastest


def summarise(run_dir, ni=21, nj=21, nk=41, last=400):
    rows = [l.split(",") for l in open(os.path.join(run_dir, "runtime_diags.csv")).read().splitlines()[1:]]
    row = [r for r in rows if int(r[0]) == last][0]
    vti = os.path.join(run_dir, "results", "fields_%05d.vti" % last)
    nd = vti_array(vti, "nd-ave.O+", ni, nj, nk)
    phi = vti_array(vti, "phi", ni, nj, nk)
    log = os.path.join(run_dir, "run.log")
    steady = None
    if os.path.exists(log):
        m = re.search(r"Steady state reached at time step (\d+)", open(log).read())
        steady = int(m.group(1)) if m else None
    return {
        "ts": last, "mp_count": float(row[3]), "real_count": float(row[4]), "pz": float(row[7]), "KE": float(row[8]),
        "PE": float(row[9]), "steady_state_ts": steady,
        "nd_ave_k_profile": nd.mean(axis=(1, 2)).tolist(),            # mean over each z plane
        "nd_ave_axis_profile": nd[:, nj // 2, ni // 2].tolist(),       # along the axis through the sphere
        "nd_ave_wake_plane": nd[30].mean(axis=0).tolist(),             # y-averaged x profile behind the sphere (k=30)
        "phi_k_profile": phi.mean(axis=(1, 2)).tolist(),
    }


if __name__ == "__main__":
    out = summarise(sys.argv[1])
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ch3_sphere_statistics.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, {k: out[k] for k in ("mp_count", "KE", "PE", "steady_state_ts")})
