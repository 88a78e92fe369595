#!/usr/bin/env python
"""Compares the observables of the book's unmodified ch4/Main.cpp run on the engine (profiles/r1_ch4_main_gpu_runs/gpu_<seed>.json,
written on a B200 by scripts/gpu_ch4_summary.sh) with the run of the compiled reference (tests/golden/ch4_neutral_flow_statistics.json),
using the quantities and tolerances of tests/test_host_shim.py::test_reference_ch4_main_neutral_flow_statistics."""
import glob, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ch4_neutral_flow_statistics.json")))
runs = sorted(glob.glob(os.path.join(ROOT, "profiles", "r1_ch4_main_gpu_runs", "gpu_*.json")))
print("reference: steady state at ts %s, mpc_total %.0f, ts=1999 %s" % (ref["steady_state_ts"], ref["mpc_total"], ref["diag"]["1999"]))
worst = {}
for path in runs:
    got = json.load(open(path))
    name = os.path.basename(path)
    d = max(abs(got["diag"][ts][k] / row[k] - 1) for ts, row in ref["diag"].items() for k in ("mp_count", "real_count", "pz", "KE"))
    line = ["diag max rel %.2e" % d, "steady %d" % got["steady_state_ts"], "mpc_total rel %.2e" % abs(got["mpc_total"] / ref["mpc_total"] - 1)]
    worst["diag"] = max(worst.get("diag", 0), d)
    for key in ("nd_ave_k_profile", "w_k_profile", "T_k_profile", "mpc_k_profile", "nd_ave_axis_profile", "T_axis_profile"):
        a, b = np.array(got[key]), np.array(ref[key])
        e = np.abs(a - b).max() / np.abs(b).max()
        worst[key] = max(worst.get(key, 0), e)
        line.append("%s %.2e" % (key.replace("_profile", ""), e))
    print(name + ": " + "  ".join(line))
print("worst over runs:", {k: float("%.3g" % v) for k, v in worst.items()})
