"""One-time, offline: run the compiled, unmodified reference (oracle/_ref/ref_ch3 = /root/reference/ch3/ver2 + the I/O harness) to
CONVERGENCE on the warm-started Poisson problem of one full-size bench step, and record what bench.py's reference arm needs:
how many solveGS sweeps (PotentialSolver.cpp:334-430, tolerance 1e-4, the solver ch3/ver2/Main.cpp ships) one step takes.

    python bench.py --dump-warm gpurun_out/warm_128.npz ...      (on a B200: phi of step n and rho of step n+1, 2e8 ions, 128^3)
    python scripts/ref_gs_convergence.py gpurun_out/warm_128.npz  (here, CPU only; writes profiles/r2_reference_gs_convergence.json)

The exact sweep count comes from the oracle's restatement of solveGS (bit-pinned to the reference by tests/test_oracle_vs_ref.py),
the wall time from the reference binary itself; bench.py --impl reference multiplies ITS OWN live per-sweep time by the count."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import statefile as sf          # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main(path):
    d = np.load(path)
    n = int(d["mesh"])
    st = sf.State()
    st.ni = st.nj = st.nk = n
    st.flags = 3
    st.x0, st.xm, st.dt = np.array((-0.1, -0.1, 0.0)), np.array((0.1, 0.1, 0.4)), 1e-7
    st.sphere_c, st.sphere_r, st.sphere_phi = np.array((0.0, 0.0, 0.15)), 0.05, -100.0
    st.phi0, st.Te0, st.n0 = 0.0, 1.5, 1e12
    st.phi, st.rho = d["phi"], d["rho_next"]
    out = {"mesh": n, "particles": int(d["particles"]), "tolerance": 1e-4, "max_it": 20000, "host": os.uname().nodename,
           "cpu": [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]}
    with tempfile.TemporaryDirectory() as tmp:
        fin = os.path.join(tmp, "in.state")
        sf.write_state(fin, st)
        env = dict(os.environ, ESPIC_REF_TIMING="1", ESPIC_REF_NODUMP="1")
        exe = os.path.join(ROOT, "oracle", "_ref", "ref_ch3")
        r = subprocess.run([exe, fin, "/dev/null", "solve_gs:50:0"], capture_output=True, text=True, env=env, check=True)
        t_sweep = float([l for l in r.stdout.splitlines() if l.startswith("T ")][0].split()[2]) / 50
        t0 = time.time()
        r = subprocess.run([exe, fin, os.path.join(tmp, "out.state"), "solve_gs:20000:1e-4"], capture_output=True, text=True,
                           env=dict(os.environ, ESPIC_REF_TIMING="1"), check=True)
        t_solve = float([l for l in r.stdout.splitlines() if l.startswith("T ")][0].split()[2])
        res = sf.read_state(os.path.join(tmp, "out.state"))
        out.update({"reference_per_sweep_s": t_sweep, "reference_solve_s": t_solve, "reference_converged": bool(res.diag[0] == 1.0),
                    "reference_stderr": r.stderr[-200:], "wall_s": time.time() - t0})
    # exact sweep count from the bit-pinned restatement
    w = orc.World(n, n, n, tuple(st.x0), tuple(st.xm))
    w.add_sphere(tuple(st.sphere_c), st.sphere_r, st.sphere_phi)
    w.add_inlet()
    w.set_reference_values(st.phi0, st.Te0, st.n0)
    w.phi[:] = st.phi
    w.rho[:] = st.rho
    info = w.solve_gs(20000, 1e-4)
    out.update({"sweeps_to_converge": int(info["gs_iters"]), "oracle_converged": int(info["converged"]),
                "phi_oracle_vs_reference_max_abs": float(np.abs(w.phi - res.phi).max()),
                "sweeps_from_time_ratio": t_solve / t_sweep,
                "note": "warm-started solveGS(20000, 1e-4) on the Poisson problem of one bench step (phi of step n, rho of step n+1, "
                        "%d ions on a %d^3 mesh, state dumped by bench.py --dump-warm on a B200); solve time measured with the compiled "
                        "reference on %s, sweep count from the bit-pinned oracle" % (out["particles"], n, out["cpu"])})
    dst = os.path.join(ROOT, "profiles", "r2_reference_gs_convergence.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "warm_128.npz"))
