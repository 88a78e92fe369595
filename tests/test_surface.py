"""Surface interactions of ch4: Species::advance(neutrals, spherium) (ch4/Species.cpp:8-100) -- per-particle
sub-step loop, World::lineSphereIntersect, diffuse re-emission of neutrals from the sphere (sampleReflectedVelocity /
sphereDiffuseVector), ions that hit the sphere die and emit neutrals + sputtered material.

CPU: the oracle restatement in its mt19937 mode reproduces the compiled, unmodified ch4 reference bit-for-bit (all three
species, Particle::dt included, particle ORDER included), live and against the committed fixture.
GPU: espic_push_surface against the oracle's Philox mode (same counters).
"""
import os
import struct
import subprocess

import numpy as np
import pytest

import statefile as sf
import surface_cases as sc
from cases import orc, QE

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ch4", "surface.npz")


def run_reference(w, adv, neut, sput, part, pdt, seed, reps, same_target, tmp_path):
    exe = os.path.join(sf.REF_DIR, "ref_ch4_surface")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_ch4_surface is built only where the reference tree is present")
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    n = part.shape[1]
    with open(fin, "wb") as f:
        f.write(struct.pack("<6i2I", w.ni, w.nj, w.nk, reps, int(same_target), 0, seed, 0))
        f.write(np.asarray(w.x0, dtype="<f8").tobytes() + np.asarray(w.xm, dtype="<f8").tobytes())
        f.write(struct.pack("<5d", sc.DT, *sc.SPH_C, sc.SPH_R))
        for s in (adv, neut, sput):
            f.write(struct.pack("<3d", s.mass, s.charge, s.mpw0))
        f.write(struct.pack("<q", n))
        f.write(np.ascontiguousarray(np.vstack([part[:6], pdt[None], part[6:7]]), dtype="<f8").tobytes())
        f.write(np.asarray(w.ef, dtype="<f8").tobytes())
    subprocess.run([exe, fin, fout], check=True)
    raw = open(fout, "rb").read()
    out, off = [], 0
    for _ in range(3):
        m = struct.unpack_from("<q", raw, off)[0]
        off += 8
        out.append(np.frombuffer(raw, dtype="<f8", count=8 * m, offset=off).reshape(8, m).copy())
        off += 64 * m
    return out


def run_oracle(w, charge, part, pdt, rng_for_rep, reps, same_target):
    adv, neut, sput = sc.species_triplet(w, charge)
    adv.set_particles(part)
    adv.pdt[:adv.np] = pdt
    for r in range(reps):
        if charge == 0:
            adv.advance_surface(sc.DT, adv, adv, rng_for_rep(r))
        else:
            adv.advance_surface(sc.DT, neut, neut if same_target else sput, rng_for_rep(r))
    return [np.vstack([s.arr[:6, :s.np], s.pdt[None, :s.np], s.arr[6:7, :s.np]]) for s in (adv, neut, sput)]


CASES = [("neutral_bounce", 0.0, True), ("ion_same_target", QE, True), ("ion_two_targets", QE, False)]


@pytest.mark.parametrize("name,charge,same", CASES)
def test_oracle_surface_advance_matches_reference_bits(name, charge, same, tmp_path):
    w = sc.make_world()
    part, pdt = sc.make_particles(w, 17, 4000, mpw=5.0)
    adv, neut, sput = sc.species_triplet(w, charge)
    ref = run_reference(w, adv, adv if charge == 0 else neut, sput, part, pdt, 2024, 2, same, tmp_path)
    mt = orc.mt19937(2024)
    got = run_oracle(w, charge, part, pdt, lambda r: ("mt", mt), 2, same)
    for s in range(3):
        assert got[s].shape == ref[s].shape, (name, s, got[s].shape, ref[s].shape)
        assert np.array_equal(got[s].view(np.uint64), ref[s].view(np.uint64)), (name, s)
    # the case must exercise what it claims to
    n0 = part.shape[1]
    if charge == 0:
        moved = ref[0].shape[1]
        assert 0.5 * n0 < moved < n0
    else:
        assert ref[1].shape[1] > 200 and (same or ref[2].shape[1] > 5)


def test_oracle_surface_advance_matches_golden():
    d = np.load(GOLD)
    w = sc.make_world()
    assert np.array_equal(w.ef, d["ef"])
    part, pdt = sc.make_particles(w, 17, 4000, mpw=5.0)
    for name, charge, same in CASES:
        mt = orc.mt19937(int(d["seed"]))
        got = run_oracle(w, charge, part, pdt, lambda r: ("mt", mt), int(d["reps"]), same)
        for s in range(3):
            assert np.array_equal(got[s].view(np.uint64), d[f"{name}_{s}"].view(np.uint64)), (name, s)


def test_bounced_neutrals_leave_the_sphere_with_wall_temperature():
    """Physics of the restatement: after the advance no live particle is inside the sphere, and re-emitted neutrals carry the
    Birdsall thermal speed of a 1000 K wall (full accommodation, ch4/Species.cpp:93-100), not their 7 km/s impact speed."""
    w = sc.make_world()
    part, pdt = sc.make_particles(w, 23, 6000, mpw=5.0)
    got = run_oracle(w, 0.0, part, pdt, lambda r: ("philox", 99, 0, r), 1, True)[0]
    c = np.array(sc.SPH_C)[:, None]
    r2 = ((got[:3] - c) ** 2).sum(0)
    assert np.all(r2 > sc.SPH_R ** 2)
    speed = np.linalg.norm(got[3:6], axis=0)
    slow = speed < 2500.0                  # wall-temperature population: v_th(1000 K, 16 amu) ~ 1019 m/s
    assert slow.sum() > 500
    v_th = np.sqrt(2 * 1.380648e-23 * 1000 / (16 * orc.AMU))
    # sampleVth: three components v_th*(u1+u2+u3-1.5), each of variance v_th^2/4, magnitude scaled by 3/sqrt(6)
    assert abs(np.sqrt((speed[slow] ** 2).mean()) / (v_th * np.sqrt(0.75 * 1.5)) - 1) < 0.1
    assert np.all(got[6] == 0)             # every survivor finished its step: Particle::dt == 0


# ---- GPU: espic_push_surface against the oracle's Philox mode -----------------------------------------------------------

def _stage(w, charge, n, seed):
    """Oracle species triplet and the matching GPU engine: the settled particles (Particle::dt = 0) are uploaded, the fresh ones
    (dt = world dt) go through addParticle on both sides."""
    from engines import GpuEngine
    part, pdt = sc.make_particles(w, seed, n, mpw=5.0)
    settled, fresh = part[:, pdt == 0], part[:, pdt != 0]
    adv, neut, sput = sc.species_triplet(w, charge, cap=2 * n)
    adv.set_particles(settled)
    st = sf.state_from_oracle(w, [adv, neut, sput], sc.DT)
    g = GpuEngine(st)
    for q in range(fresh.shape[1]):
        adv.add_particle(fresh[:3, q], fresh[3:6, q], fresh[6, q], sc.DT)
    adv.pdt[settled.shape[1]:adv.np] = sc.DT
    added = g.e.add_particles(g.species[0], np.ascontiguousarray(fresh), sc.DT)
    assert added == fresh.shape[1] == adv.np - settled.shape[1]
    return g, (adv, neut, sput)


def _compare(g, osp, idx, exact_positions):
    a, b = g.e.download(g.species[idx]), osp.particles()
    assert a.shape == b.shape, (idx, a.shape, b.shape)
    if a.shape[1] == 0:
        return 0
    if exact_positions:
        assert np.array_equal(a[:3].view(np.uint64), b[:3].view(np.uint64)), "positions"
    else:
        # a re-emitted particle moves on with a velocity that went through sin/cos: CUDA's differ from glibc's by <= 1 ulp
        assert np.abs(a[:3] - b[:3]).max() <= 1e-13, "positions"
    assert np.array_equal(a[6], b[6]), "weights"
    assert np.abs(a[3:6] - b[3:6]).max() <= 1e-12 * np.abs(b[3:6]).max(), "velocities (libm differences only)"
    return a.shape[1]


@pytest.mark.gpu
def test_gpu_neutral_surface_bounce_matches_oracle_philox():
    w = sc.make_world()
    g, (adv, _, _) = _stage(w, 0.0, 6000, 41)
    n0 = adv.np
    for step in range(3):
        g.e.push_surface(g.species[0], sc.DT, 0, 0, 0x5EED5EED77, 4, step)
        adv.advance_surface(sc.DT, adv, adv, ("philox", 0x5EED5EED77, 4, step))
        n = _compare(g, adv, 0, exact_positions=False)
    assert 0.3 * n0 < n < n0
    # bit-exact where no libm call is involved: particles that never touched the sphere
    a, b = g.e.download(g.species[0]), adv.particles()
    same = np.all(a[3:6].view(np.uint64) == b[3:6].view(np.uint64), axis=0)
    assert same.mean() > 0.3 and np.array_equal(a[:3, same].view(np.uint64), b[:3, same].view(np.uint64))


@pytest.mark.gpu
@pytest.mark.parametrize("same_target", [True, False])
def test_gpu_ion_impact_emission_matches_oracle_philox(same_target):
    w = sc.make_world()
    g, (adv, neut, sput) = _stage(w, QE, 6000, 43)
    tgt = neut if same_target else sput
    gi = 1 if same_target else 2
    for step in range(2):
        em_g = g.e.push_surface(g.species[0], sc.DT, g.species[1], g.species[gi], 0xABCDEF, 1, step)
        em_o = adv.advance_surface(sc.DT, neut, tgt, ("philox", 0xABCDEF, 1, step))
        if same_target:
            assert em_g[0] == sum(em_o) and em_g[1] == 0
        else:
            assert em_g == em_o
        _compare(g, adv, 0, exact_positions=True)          # ions: no libm on their own path
        assert _compare(g, neut, 1, exact_positions=True) > (300 if step == 0 else 0)
        ns = _compare(g, sput, 2, exact_positions=True)
        assert same_target or ns > 5 or step > 0
    # the emitted neutrals are new particles of their species: their first advance moves them for 2*dt and bounces them
    g.e.push_surface(g.species[1], sc.DT, 0, 0, 0xABCDEF, 2, 7)
    neut.advance_surface(sc.DT, neut, neut, ("philox", 0xABCDEF, 2, 7))
    assert _compare(g, neut, 1, exact_positions=False) > 300


@pytest.mark.gpu
def test_gpu_sort_refused_between_injection_and_surface_advance():
    w = sc.make_world()
    g, _ = _stage(w, 0.0, 500, 47)
    es = g.es
    g.e.push_surface(g.species[0], sc.DT, 0, 0, 1, 0, 0)
    g.e.sort_by_cell(g.species[0])                                   # fine: everything settled
    g.e.inject_warm_beam(g.species[0], 7000.0, 2e7, 1000.0, sc.DT, 5, 0, 1)
    with pytest.raises(Exception, match="sort after the advance"):
        g.e.sort_by_cell(g.species[0])
    g.e.push_surface(g.species[0], sc.DT, 0, 0, 1, 0, 1)
    g.e.sort_by_cell(g.species[0])
