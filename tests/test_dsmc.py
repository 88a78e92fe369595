"""DSMC collisions of ch4 (cell-indexed collisions): DSMC_MEX::apply / collide / evalSigma
(ch4/Collisions.cpp:84-182, ch4/Collisions.h:59-83: Bird's no-time-counter scheme with VHS cross-sections, momentum exchange
between particles of one species in the same cell) and Species::computeMPC (ch4/Species.cpp:228-235).

CPU: the oracle restatement in its mt19937 mode reproduces the compiled, unmodified ch4 reference bit-for-bit (velocities of
every particle, sigma_cr_max carried between calls, macroparticles per cell), live and against the committed fixture; momentum
and energy are conserved by every collision.
GPU: espic_dsmc_mex / espic_compute_mpc against the oracle's Philox mode (same counters).
"""
import os
import struct
import subprocess

import numpy as np
import pytest

import cases
import statefile as sf
from cases import orc, AMU

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ch4", "dsmc.npz")
DT, MASS, MPW0 = 2e-6, 16 * AMU, 5e13


def make_case(seed=9, n=20000, dims=(9, 9, 13)):
    rng = np.random.default_rng(seed)
    w = cases.sphere_world(*dims)
    part = cases.random_particles(w, rng, n, v_drift=7000.0, v_th=900.0, mpw=MPW0)
    return w, part


def run_reference(w, part, seed, reps, tmp_path):
    exe = os.path.join(sf.REF_DIR, "ref_ch4_dsmc")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_ch4_dsmc is built only where the reference tree is present")
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    n = part.shape[1]
    with open(fin, "wb") as f:
        f.write(struct.pack("<4i2I", w.ni, w.nj, w.nk, reps, seed, 0))
        f.write(np.asarray(w.x0, dtype="<f8").tobytes() + np.asarray(w.xm, dtype="<f8").tobytes())
        f.write(struct.pack("<3d", DT, MASS, MPW0))
        f.write(struct.pack("<q", n))
        f.write(np.ascontiguousarray(part, dtype="<f8").tobytes())
    subprocess.run([exe, fin, fout], check=True)
    raw = open(fout, "rb").read()
    m = struct.unpack_from("<q", raw, 0)[0]
    out = np.frombuffer(raw, dtype="<f8", count=7 * m, offset=8).reshape(7, m).copy()
    sig = struct.unpack_from("<d", raw, 8 + 56 * m)[0]
    mpc = np.frombuffer(raw, dtype="<f8", offset=16 + 56 * m).copy()
    return out, sig, mpc


def run_oracle(w, part, reps, rng_for_rep):
    sp = orc.Species(w, MASS, 0.0, MPW0, cap=part.shape[1])
    sp.set_particles(part)
    sig, cols = 1e-14, []                 # DSMC_MEX's initial sigma_cr_max (ch4/Collisions.h:77)
    for r in range(reps):
        c, sig = sp.dsmc_mex(DT, sig, rng_for_rep(r))
        cols.append(c)
    return sp, sig, cols


def test_oracle_dsmc_matches_reference_bits(tmp_path):
    w, part = make_case()
    ref, sig_ref, mpc_ref = run_reference(w, part, 4242, 3, tmp_path)
    mt = orc.mt19937(4242)
    sp, sig, cols = run_oracle(w, part, 3, lambda r: ("mt", mt))
    got = sp.particles()
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64))
    assert sig == sig_ref and sig != 1e-14
    assert np.array_equal(sp.compute_mpc(), mpc_ref) and mpc_ref.sum() == part.shape[1]
    changed = np.any(got[3:6] != part[3:6], axis=0).sum()
    assert min(cols) > 200 and changed > 1000, (cols, changed)


def test_oracle_dsmc_matches_golden():
    d = np.load(GOLD)
    w, part = make_case()
    mt = orc.mt19937(int(d["seed"]))
    sp, sig, _ = run_oracle(w, part, int(d["reps"]), lambda r: ("mt", mt))
    assert np.array_equal(sp.particles().view(np.uint64), d["part"].view(np.uint64))
    assert sig == float(d["sigma_cr_max"])
    assert np.array_equal(sp.compute_mpc(), d["mpc"])


def test_dsmc_conserves_momentum_and_energy():
    w, part = make_case(seed=12)
    sp, _, cols = run_oracle(w, part, 2, lambda r: ("philox", 31337, 5, r))
    got = sp.particles()
    assert sum(cols) > 500
    for c in range(3, 6):
        assert abs(got[c].sum() - part[c].sum()) <= 1e-9 * np.abs(part[c]).sum()
    ke0, ke1 = (part[3:6] ** 2).sum(), (got[3:6] ** 2).sum()
    assert abs(ke1 / ke0 - 1) < 1e-12
    assert np.array_equal(got[:3], part[:3]) and np.array_equal(got[6], part[6])


# ---- GPU -------------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_gpu_dsmc_matches_oracle_philox():
    from engines import GpuEngine
    w, part = make_case(seed=21, n=40000)
    sp0 = orc.Species(w, MASS, 0.0, MPW0, cap=part.shape[1])
    sp0.set_particles(part)
    g = GpuEngine(sf.state_from_oracle(w, [sp0], DT))
    sp, sig_o = sp0, 1e-14
    sig_g = 1e-14
    for step in range(3):
        cols_g, sig_g = g.e.dsmc_mex(g.species[0], DT, sig_g, 0xD5C0FFEE, 3, step)
        cols_o, sig_o = sp.dsmc_mex(DT, sig_o, ("philox", 0xD5C0FFEE, 3, step))
        assert cols_g == cols_o and cols_o > 300, (step, cols_g, cols_o)
        assert abs(sig_g / sig_o - 1) < 1e-13            # pow() of the CUDA math library vs glibc
    a, b = g.e.download(g.species[0]), sp.particles()
    assert np.array_equal(a[:3], b[:3]) and np.array_equal(a[6], b[6])
    assert np.abs(a[3:6] - b[3:6]).max() <= 1e-12 * np.abs(b[3:6]).max()
    untouched = np.all(b[3:6] == part[3:6], axis=0)
    assert np.array_equal(a[3:6, untouched], part[3:6, untouched]) and 0.2 < untouched.mean() < 0.98
    assert np.array_equal(g.e.compute_mpc(g.species[0]), sp.compute_mpc())
