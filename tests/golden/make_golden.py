"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/ref_ch3, ref_ch2:
reference sources compiled where they lie by oracle/Makefile) on seeded inputs.

    python tests/golden/make_golden.py          (needs /root/reference; run in the build container)

Each file holds: cmds, which harness, the input state ("in_*") and the reference's output state ("out_*").
tests/test_golden.py replays the commands with the oracle (CPU suite) and tests/test_gpu_parity.py with
the CUDA library (-m gpu); neither needs /root/reference at run time.
"""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases            # noqa: E402
import statefile as sf  # noqa: E402
from cases import orc, QE, AMU, ME  # noqa: E402


def with_rho(seed, n=4000, dims=(9, 9, 13), drop_particles=False, warm=True, **kw):
    """Near-equilibrium field state: ion density ~ n0, phi pre-solved with the (robust) nonlinear GS, then the
    particles move three steps so the solver under test has real work to do from a warm start.  (The reference's
    Newton-PCG diverges from a cold phi=0 start; its own Main starts from the ctor's solveQN guess instead.)"""
    w, sp = cases.sphere_case(seed=seed, ni=dims[0], nj=dims[1], nk=dims[2], n=n, amp=0.0, mpw=1e10 * 0.016 / n, **kw)
    sp.compute_number_density()
    w.compute_charge_density([sp])
    if warm:
        assert w.solve_gs(20000, 1e-6)["converged"]
        w.compute_ef()
        for _ in range(3):
            sp.advance(1e-7)
        sp.compute_number_density()
        w.compute_charge_density([sp])
    if drop_particles:
        sp.set_particles(np.zeros((7, 0)))
        sp.den[:] = 0
    return w, sp


def scenarios():
    out = {}
    w, sp = cases.sphere_case(seed=101, n=1500)
    out["sphere_push_quiet"] = ("ref_ch3", sf.state_from_oracle(w, [sp], 1e-7),
                                ["advance", "deposit", "rho", "advance", "deposit", "rho"])
    w, sp = cases.sphere_case(seed=102, n=1500, near_walls=0.3)
    out["sphere_push_kill"] = ("ref_ch3", sf.state_from_oracle(w, [sp], 2e-6),
                               ["advance", "advance", "advance", "deposit", "rho"])
    w, sp = with_rho(103, warm=False)
    w.set_reference_values(0.5, 2.0, 3e9)
    out["sphere_qn_ef"] = ("ref_ch3", sf.state_from_oracle(w, [sp], 1e-7), ["solve_qn", "ef"])
    w, sp = with_rho(104, drop_particles=True)
    out["sphere_gs"] = ("ref_ch3", sf.state_from_oracle(w, [sp], 1e-7), ["solve_gs:5000:1e-4", "ef"])
    w, sp = with_rho(105, drop_particles=True)
    out["sphere_pcg"] = ("ref_ch3", sf.state_from_oracle(w, [sp], 1e-7), ["solve_pcg:2000:1e-4", "ef"])
    w, sp = with_rho(106, n=100000, dims=(21, 21, 41), drop_particles=True)
    out["sphere_pcg_shipped_mesh"] = ("ref_ch3", sf.state_from_oracle(w, [sp], 1e-7), ["solve_pcg:2000:1e-4"])
    w, sp = with_rho(107, drop_particles=True)
    out["sphere_pcg_fallback"] = ("ref_ch3", sf.state_from_oracle(w, [sp], 1e-7), ["solve_pcg:12:1e-4"])
    w, sp = cases.sphere_case(seed=108, n=20)
    out["sphere_sample_mt"] = ("ref_ch3", sf.state_from_oracle(w, [sp], 1e-7), ["sample:0:7000:2e9:4242:2"])
    w, sp = with_rho(109, n=3000)
    out["sphere_full_steps"] = ("ref_ch3", sf.state_from_oracle(w, [sp], 1e-7),
                                ["advance", "deposit", "rho", "solve_pcg:3000:1e-6", "ef"] * 3 + ["average:0"])
    w, sp = with_rho(111, n=3000)
    out["sphere_full_steps_gs"] = ("ref_ch3", sf.state_from_oracle(w, [sp], 1e-7),
                                   ["advance", "deposit", "rho", "solve_gs:5000:1e-6", "ef"] * 3)
    w, sps = cases.box_case(seed=110, npart=1200)
    out["box_step"] = ("ref_ch2", sf.state_from_oracle(w, sps, 2e-9),
                       ["advance", "deposit", "rho", "solve:3000:1e-4", "ef", "advance", "deposit", "rho"])
    w = cases.box_world(9)
    out["box_quiet_start"] = ("ref_ch2", sf.state_from_oracle(w, [orc.Species(w, 16 * AMU, QE), orc.Species(w, ME, -QE)], 2e-10),
                              ["loadqs:0:1e11:11:11:11:0", "loadqs:1:1e11:6:6:6:1", "deposit", "rho", "solve:3000:1e-4", "ef"])
    return out


def main():
    for name, (which, st, cmds) in scenarios().items():
        with tempfile.TemporaryDirectory() as tmp:
            res = sf.run_ref(which, st, cmds, tmp)
        d = dict(cmds=np.array(cmds), which=np.array(which), stderr=np.array(res.stderr))
        d.update(sf.state_to_dict(st, "in_"))
        d.update(sf.state_to_dict(res, "out_"))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print("%-28s %8.1f KB  converged=%g" % (name, os.path.getsize(path) / 1024, res.diag[0]))


if __name__ == "__main__":
    main()
