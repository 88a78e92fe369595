"""Generates tests/golden/ch4_neutral_flow_statistics.json from a run of the UNMODIFIED reference ch4 program (oracle/_ref/
ref_ch4_main = g++ -O2 of /root/reference/ch4/*.cpp; mkdir results; run): warm neutral beam past the sphere with diffuse
re-emission from its surface and DSMC collisions, 2000 steps.  The reference seeds its RNG from std::random_device, so these are
statistical pins.

    python tests/golden/make_ch4_statistics.py <run directory with runtime_diags.csv, run.log and results/fields_01999.vti>
"""
import json
import os
import re
import sys

import numpy as np

NI, NJ, NK = 41, 21, 41
TS = (100, 300, 600, 1000, 1500, 1999)


def vti_array(txt, name, shape):
    m = re.search(r'<DataArray Name="%s"[^>]*>\n(.*?)</DataArray>' % re.escape(name), txt, re.S)
    return np.array(m.group(1).split(), dtype=np.float64).reshape(shape)


def summarise(run_dir, last=1999):
    rows = [l.split(",") for l in open(os.path.join(run_dir, "runtime_diags.csv")).read().splitlines()[1:]]
    by_ts = {int(r[0]): r for r in rows}
    txt = open(os.path.join(run_dir, "results", "fields_%05d.vti" % last)).read()
    nd = vti_array(txt, "nd-ave.O", (NK, NJ, NI))              # VTK order: i fastest
    T = vti_array(txt, "T.O", (NK, NJ, NI))
    vel = vti_array(txt, "vel.O", (NK, NJ, NI, 3))
    mpc = vti_array(txt, "mpc.O", (NK - 1, NJ - 1, NI - 1))
    log = os.path.join(run_dir, "run.log")
    steady = None
    if os.path.exists(log):
        m = re.search(r"Steady state reached at time step (\d+)", open(log).read())
        steady = int(m.group(1)) if m else None
    jm, im = NJ // 2, NI // 2
    return {
        "diag": {str(t): {"mp_count": float(by_ts[t][3]), "real_count": float(by_ts[t][4]), "pz": float(by_ts[t][7]),
                          "KE": float(by_ts[t][8])} for t in TS},
        "steady_state_ts": steady,
        "nd_ave_k_profile": nd.mean(axis=(1, 2)).tolist(),              # mean over each z plane
        "nd_ave_axis_profile": nd[:, jm - 1:jm + 2, im - 1:im + 2].mean(axis=(1, 2)).tolist(),   # 3x3 column through the sphere
        "T_k_profile": T.mean(axis=(1, 2)).tolist(),
        "T_axis_profile": T[:, jm - 1:jm + 2, im - 1:im + 2].mean(axis=(1, 2)).tolist(),
        "w_k_profile": vel[..., 2].mean(axis=(1, 2)).tolist(),          # stream velocity along the beam
        "mpc_total": float(mpc.sum()),
        "mpc_k_profile": mpc.sum(axis=(1, 2)).tolist(),
    }


if __name__ == "__main__":
    out = summarise(sys.argv[1])
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ch4_neutral_flow_statistics.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, out["diag"]["1999"], out["steady_state_ts"])
