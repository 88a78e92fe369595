"""MCC charge-exchange operator of ch4: MCC_CEX::apply (ch4/Collisions.cpp:43-82) -- every source particle collides with the
mesh-averaged target gas with probability 1 - exp(-n*sigma*|v - u|*dt).

CPU: the oracle in its mt19937 mode reproduces the compiled, unmodified reference bit-for-bit (live and against the fixture).
GPU: espic_mcc_cex against the oracle's Philox mode.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

import cases
import statefile as sf
from cases import orc, AMU, QE

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ch4", "mcc.npz")
DT, MASS = 2e-6, 16 * AMU


def make_case(seed=14, n=20000, dims=(9, 9, 13)):
    rng = np.random.default_rng(seed)
    w = cases.sphere_world(*dims)
    part = cases.random_particles(w, rng, n, v_drift=7000.0, v_th=900.0, mpw=50.0)
    i, j, k = np.meshgrid(np.arange(w.ni), np.arange(w.nj), np.arange(w.nk), indexing="ij")
    u = (k * w.ni * w.nj + j * w.ni + i).ravel()
    den, T, vel = np.zeros(w.nn), np.zeros(w.nn), np.zeros((w.nn, 3))
    den[u] = (4e17 * (0.2 + np.sin(0.5 * i) ** 2 + 0.3 * np.cos(0.7 * k) ** 2 + 0.1 * j)).ravel()
    T[u] = (800.0 + 50.0 * i + 20.0 * k).ravel()
    vel[u, 0] = (300.0 * np.sin(0.9 * j)).ravel()
    vel[u, 1] = (-200.0 + 40.0 * i).ravel()
    vel[u, 2] = (5000.0 + 400.0 * np.cos(0.4 * k)).ravel()
    return w, part, den, np.ascontiguousarray(vel.ravel()), T


def run_reference(w, part, den, vel, T, seed, reps, tmp_path):
    exe = os.path.join(sf.REF_DIR, "ref_ch4_mcc")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_ch4_mcc is built only where the reference tree is present")
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<4i2I", w.ni, w.nj, w.nk, reps, seed, 0))
        f.write(np.asarray(w.x0, dtype="<f8").tobytes() + np.asarray(w.xm, dtype="<f8").tobytes())
        f.write(struct.pack("<3d", DT, MASS, MASS))
        f.write(struct.pack("<q", part.shape[1]))
        for a in (part, den, vel, T):
            f.write(np.ascontiguousarray(a, dtype="<f8").tobytes())
    subprocess.run([exe, fin, fout], check=True)
    raw = open(fout, "rb").read()
    m = struct.unpack_from("<q", raw, 0)[0]
    return np.frombuffer(raw, dtype="<f8", count=7 * m, offset=8).reshape(7, m).copy()


def run_oracle(w, part, den, vel, T, reps, rng_for_rep):
    sp = orc.Species(w, MASS, QE, 50.0, cap=part.shape[1])
    sp.set_particles(part)
    cols = [sp.mcc_cex(den, vel, T, MASS, DT, rng_for_rep(r)) for r in range(reps)]
    return sp, cols


def test_oracle_mcc_matches_reference_bits(tmp_path):
    w, part, den, vel, T = make_case()
    ref = run_reference(w, part, den, vel, T, 777, 2, tmp_path)
    mt = orc.mt19937(777)
    sp, cols = run_oracle(w, part, den, vel, T, 2, lambda r: ("mt", mt))
    assert np.array_equal(sp.particles().view(np.uint64), ref.view(np.uint64))
    stopped = np.all(ref[3:6] == 0, axis=0).sum()
    assert 0.1 * part.shape[1] < cols[0] < 0.6 * part.shape[1] and stopped >= cols[0]


def test_oracle_mcc_matches_golden():
    d = np.load(GOLD)
    w, part, den, vel, T = make_case()
    mt = orc.mt19937(int(d["seed"]))
    sp, _ = run_oracle(w, part, den, vel, T, int(d["reps"]), lambda r: ("mt", mt))
    assert np.array_equal(sp.particles().view(np.uint64), d["part"].view(np.uint64))


@pytest.mark.gpu
def test_gpu_mcc_matches_oracle_philox():
    from engines import GpuEngine, _espic
    es = _espic()
    w, part, den, vel, T = make_case(seed=15, n=50000)
    src = orc.Species(w, MASS, QE, 50.0, cap=part.shape[1])
    src.set_particles(part)
    tgt = orc.Species(w, MASS, 0.0, 1e12, cap=16)
    g = GpuEngine(sf.state_from_oracle(w, [src, tgt], DT))
    g.e.set_field(es.DEN, den, g.species[1])
    g.e.set_field(es.VEL, vel, g.species[1])
    for step in range(2):
        cols_g = g.e.mcc_cex(g.species[0], g.species[1], DT, 0xC0111DE, 9, step)
        cols_o = src.mcc_cex(den, vel, T, MASS, DT, ("philox", 0xC0111DE, 9, step))
        assert cols_g == cols_o and cols_o > 0.1 * part.shape[1] * (0.5 ** step), (step, cols_g, cols_o)
    a, b = g.e.download(g.species[0]), src.particles()
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
