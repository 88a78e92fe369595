"""Velocity moments: Species::sampleMoments / computeGasProperties / clearSamples of ch4
(ch4/Species.cpp:190-241) -- the mesh-averaged velocity and temperature the north-star's second parity check names.

CPU: the oracle restatement is pinned bit-for-bit against the compiled, unmodified ch4 reference (oracle/_ref/ref_ch4_moments)
and against the committed golden fixture generated from it (tests/golden/moments.npz).
GPU: espic_sample_moments / espic_compute_gas_properties through the C ABI against the oracle and the fixture.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

import cases
import statefile as sf
from cases import orc, AMU, QE

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ch4", "moments.npz")
NAMES = ("n_sum", "nv_sum", "nuu_sum", "nvv_sum", "nww_sum", "vel", "T")


def make_case(seed=7, dims=(9, 8, 11), n=5000):
    w, sp = cases.sphere_case(seed=seed, ni=dims[0], nj=dims[1], nk=dims[2], n=n, v_th=2000.0)
    part = sp.particles()
    part[6] *= np.random.default_rng(seed).uniform(0.5, 1.5, size=part.shape[1])     # unequal weights
    part = np.ascontiguousarray(part[:, part[0] < 0.02])                             # leaves the nodes at large x empty
    sp.set_particles(part)
    return w, sp


def run_reference(w, sp, reps, tmp_path):
    exe = os.path.join(sf.REF_DIR, "ref_ch4_moments")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_ch4_moments is built only where the reference tree is present")
    part = sp.particles()
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<4i", w.ni, w.nj, w.nk, reps))
        f.write(np.asarray(w.x0, dtype="<f8").tobytes() + np.asarray(w.xm, dtype="<f8").tobytes())
        f.write(struct.pack("<dq", sp.mass, part.shape[1]))
        f.write(np.ascontiguousarray(part, dtype="<f8").tobytes())
    subprocess.run([exe, fin, fout], check=True)
    raw = np.fromfile(fout, dtype="<f8")
    nn = w.nn
    sizes = (nn, 3 * nn, nn, nn, nn, 3 * nn, nn)
    out, o = {}, 0
    for name, sz in zip(NAMES, sizes):
        out[name] = raw[o:o + sz].copy()
        o += sz
    assert o == raw.size
    return out


def oracle_moments(sp, reps):
    sp.clear_samples()
    for _ in range(reps):
        sp.sample_moments()
    sp.compute_gas_properties()
    return {k: getattr(sp, k).copy() for k in NAMES}


def test_oracle_moments_match_reference_bits(tmp_path):
    w, sp = make_case()
    ref = run_reference(w, sp, 3, tmp_path)
    got = oracle_moments(sp, 3)
    for k in NAMES:
        assert np.array_equal(got[k].view(np.uint64), ref[k].view(np.uint64)), k
    assert (ref["T"] > 0).any() and (ref["n_sum"] == 0).any(), "the case must cover empty and populated nodes"


def test_oracle_moments_match_golden():
    d = np.load(GOLD)
    w = orc.World(int(d["ni"]), int(d["nj"]), int(d["nk"]), tuple(d["x0"]), tuple(d["xm"]))
    sp = orc.Species(w, float(d["mass"]), 0.0, 1.0, cap=2 * d["part"].shape[1])
    sp.set_particles(d["part"])
    got = oracle_moments(sp, int(d["reps"]))
    for k in NAMES:
        assert np.array_equal(got[k].view(np.uint64), d[k].view(np.uint64)), k


@pytest.mark.gpu
@pytest.mark.parametrize("sort", [False, True])
def test_gpu_moments(sort):
    """Sums differ from the reference only by the order of the atomic additions (1e-12 of the field maximum); velocity and
    temperature are pointwise functions of the sums: T subtracts nearly equal numbers (<u2> - <u>^2 with a 7 km/s drift), so
    its tolerance is 1e-12 of <u2>, i.e. relative to mass/(2K) * max(nuu/n)."""
    from engines import GpuEngine, _espic
    es = _espic()
    d = np.load(GOLD)
    w = orc.World(int(d["ni"]), int(d["nj"]), int(d["nk"]), tuple(d["x0"]), tuple(d["xm"]))
    sp = orc.Species(w, float(d["mass"]), 0.0, 1.0, cap=2 * d["part"].shape[1])
    sp.set_particles(d["part"])
    st = sf.state_from_oracle(w, [sp], 1e-7)
    g = GpuEngine(st)
    e, s0 = g.e, g.species[0]
    if sort:
        e.sort_by_cell(s0)
    e.clear_samples(s0)
    for _ in range(int(d["reps"])):
        e.sample_moments(s0)
    e.compute_gas_properties(s0)
    got = {"n_sum": e.field(es.N_SUM, s0), "nv_sum": e.field(es.NV_SUM, s0), "nuu_sum": e.field(es.NUU_SUM, s0),
           "nvv_sum": e.field(es.NVV_SUM, s0), "nww_sum": e.field(es.NWW_SUM, s0), "vel": e.field(es.VEL, s0), "T": e.field(es.T, s0)}
    for k in ("n_sum", "nv_sum", "nuu_sum", "nvv_sum", "nww_sum"):
        assert np.abs(got[k] - d[k]).max() <= 1e-12 * np.abs(d[k]).max(), k
    assert np.abs(got["vel"] - d["vel"]).max() <= 1e-11 * np.abs(d["vel"]).max()
    K = 1.380648e-23
    t_scale = float(d["mass"]) / (2 * K) * (d["nww_sum"][d["n_sum"] > 0] / d["n_sum"][d["n_sum"] > 0]).max()
    assert np.abs(got["T"] - d["T"]).max() <= 1e-11 * t_scale
    assert np.array_equal(got["T"] == 0, d["T"] == 0), "empty nodes stay exactly zero"
    # clearSamples
    e.clear_samples(s0)
    assert not e.field(es.N_SUM, s0).any() and not e.field(es.NV_SUM, s0).any()
