"""Pins the C restatement (oracle/espic_oracle.c) against the UNMODIFIED reference compiled into
oracle/_ref (ref_ch3 = ch3/ver2 sources, ref_ch2 = ch2 sources, ref_mt = ch9/MT sources).  Bit-exact unless stated.
Skipped where oracle/_ref was never built (it is built by __graft_entry__.build() / oracle/Makefile
whenever /root/reference is present, and travels to the GPU box)."""
import numpy as np
import pytest

import cases
import statefile as sf
from cases import orc, QE, AMU, ME

ref3 = pytest.mark.skipif(not sf.have_ref("ref_ch3"), reason="oracle/_ref/ref_ch3 not built")
ref2 = pytest.mark.skipif(not sf.have_ref("ref_ch2"), reason="oracle/_ref/ref_ch2 not built")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bits(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    bad = np.nonzero(bits(a) != bits(b))[0]
    assert bad.size == 0, "%s: %d/%d differ, first %s: %r vs %r" % (
        what, bad.size, a.size, bad[:3], a.ravel()[bad[:3]], b.ravel()[bad[:3]])


def warm_case(seed, dims, n):
    """ion density ~ n0 with phi pre-solved by the nonlinear GS, then three pushes (see make_golden.with_rho)."""
    w, sp = cases.sphere_case(seed=seed, ni=dims[0], nj=dims[1], nk=dims[2], n=n, amp=0.0, mpw=1e10 * 0.016 / n)
    sp.compute_number_density()
    w.compute_charge_density([sp])
    assert w.solve_gs(20000, 1e-6)["converged"]
    w.compute_ef()
    for _ in range(3):
        sp.advance(1e-7)
    sp.compute_number_density()
    w.compute_charge_density([sp])
    return w, sp


@ref3
def test_geometry_setup(tmp_path):
    """setExtents / computeNodeVolumes / addSphere / addInlet (World.cpp:22-36,58-69,87-115)."""
    for dims in ((9, 9, 13), (21, 21, 41), (12, 7, 10)):
        w = cases.sphere_world(*dims)
        st = sf.state_from_oracle(w, [], 1e-7, flags=1 | 2 | 4)
        st.phi = np.zeros(w.nn)
        r = sf.run_ref("ref_ch3", st, [], tmp_path)
        assert np.array_equal(r.object_id, w.object_id)
        assert_bits(r.node_vol, w.node_vol, "node_vol")
        assert_bits(r.phi, w.phi, "phi fixed")


@ref3
def test_compute_ef(tmp_path):
    w, sp = cases.sphere_case(seed=3)
    st = sf.state_from_oracle(w, [sp], 1e-7)
    r = sf.run_ref("ref_ch3", st, ["ef"], tmp_path)
    assert_bits(r.ef, w.ef, "ef")
    assert_bits(np.array([r.diag[1]]), np.array([w.pe()]), "PE")


@ref3
@pytest.mark.parametrize("seed,near", [(4, 0.0), (5, 0.3)])
def test_advance_deposit_rho(tmp_path, seed, near):
    """Species::advance incl. kill + swap compaction order, computeNumberDensity, computeChargeDensity, diagnostics."""
    w, sp = cases.sphere_case(seed=seed, n=3000, near_walls=near)
    dt = 2e-6 if near else 1e-7   # large dt: many particles reach the sphere / leave
    st = sf.state_from_oracle(w, [sp], dt)
    r = sf.run_ref("ref_ch3", st, ["advance", "deposit", "rho", "advance", "advance", "deposit", "rho"], tmp_path)
    n0 = sp.np
    for _ in range(3):
        sp.advance(dt)
    sp.compute_number_density()
    w.compute_charge_density([sp])
    assert r.species[0]["part"].shape[1] == sp.np
    if near:
        assert sp.np < n0 - 50      # the case really exercises removal
    assert_bits(r.species[0]["part"], sp.particles(), "particles (order included)")
    assert_bits(r.species[0]["den"], sp.den, "den")
    assert_bits(r.rho, w.rho, "rho")
    assert_bits(r.diag[2:7], np.concatenate([[sp.real_count()], sp.momentum(), [sp.ke()]]), "diag")


refmt = pytest.mark.skipif(not sf.have_ref("ref_mt"), reason="oracle/_ref/ref_mt not built")


@refmt
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_threaded_reference_build_gives_the_serial_particles(tmp_path, threads):
    """ch9/MT (std::thread advance and density scatter, ch9/MT/Species.cpp:9-123): whatever the thread count, the particles after
    advance + removal equal the serial algorithm's bit for bit -- values AND order (the removal sweep stays serial) -- and the density
    differs only by the order in which the per-thread buffers are added (1e-13 of the maximum).  This is the build the CPU baseline
    of bench.py times on all host cores."""
    w, sp = cases.sphere_case(seed=5, n=5000, near_walls=0.3)
    dt = 2e-6
    st = sf.state_from_oracle(w, [sp], dt)
    r = sf.run_ref("ref_mt", st, ["threads:%d" % threads, "advance", "deposit", "rho", "advance", "advance", "deposit", "rho"], tmp_path)
    n0 = sp.np
    for _ in range(3):
        sp.advance(dt)
    sp.compute_number_density()
    w.compute_charge_density([sp])
    assert sp.np < n0 - 50 and r.species[0]["part"].shape[1] == sp.np
    assert_bits(r.species[0]["part"], sp.particles(), "particles (order included)")
    scale = np.abs(sp.den).max()
    assert np.abs(r.species[0]["den"] - sp.den).max() <= 1e-13 * scale
    assert np.abs(r.rho - w.rho).max() <= 1e-13 * np.abs(w.rho).max()
    if threads == 1:
        assert_bits(r.species[0]["den"], sp.den, "den, one thread")


@ref3
def test_solve_qn_and_ctor(tmp_path):
    w, sp = cases.sphere_case(seed=6, n=4000)
    sp.compute_number_density()
    w.compute_charge_density([sp])
    w.set_reference_values(0.5, 2.0, 3e9)
    st = sf.state_from_oracle(w, [sp], 1e-7)
    r = sf.run_ref("ref_ch3", st, ["solve_qn", "ef"], tmp_path)
    w.solve_qn()
    w.compute_ef()
    assert_bits(r.phi, w.phi, "phi QN")
    assert_bits(r.ef, w.ef, "ef")


@ref3
def test_solve_gs_nonlinear(tmp_path):
    """solveGS: identical sweep order => bit-identical iterates and iteration count."""
    w, sp = cases.sphere_case(seed=7, n=4000)
    sp.compute_number_density()
    w.compute_charge_density([sp])
    st = sf.state_from_oracle(w, [sp], 1e-7)
    r = sf.run_ref("ref_ch3", st, ["solve_gs:5000:1e-4"], tmp_path)
    info = w.solve_gs(5000, 1e-4)
    assert r.diag[0] == 1.0 and info["converged"] == 1
    assert_bits(r.phi, w.phi, "phi GS")
    # non-converged branch (max_it hit)
    w2, _ = cases.sphere_case(seed=7, n=10)
    w2.rho[:] = st.rho
    st2 = sf.state_from_oracle(w2, [], 1e-7)
    r2 = sf.run_ref("ref_ch3", st2, ["solve_gs:30:1e-12"], tmp_path)
    info2 = w2.solve_gs(30, 1e-12)
    assert r2.diag[0] == 0.0 and info2["converged"] == 0
    assert_bits(r2.phi, w2.phi, "phi GS 30 sweeps")


@ref3
@pytest.mark.parametrize("dims", [(9, 9, 13), (21, 21, 41)])
def test_solve_nrpcg(tmp_path, dims):
    """solveNRPCG + solvePCGLinear (+ solveGSLinear fallback): same operation order => bit-identical."""
    w, sp = warm_case(8, dims, 8000 if dims[0] < 20 else 100000)
    st = sf.state_from_oracle(w, [sp], 1e-7)
    r = sf.run_ref("ref_ch3", st, ["solve_pcg:2000:1e-4"], tmp_path)
    info = w.solve_nrpcg(2000, 1e-4)
    assert r.diag[0] == float(info["converged"]) == 1.0
    assert info["lin_iters"] > 50
    assert_bits(r.phi, w.phi, "phi NR-PCG")


@ref3
def test_solve_nrpcg_fallback(tmp_path):
    """max_it too small for PCG => 'PCG failed to converge' and the GS fallback path is taken."""
    w, sp = warm_case(9, (9, 9, 13), 4000)
    st = sf.state_from_oracle(w, [sp], 1e-7)
    r = sf.run_ref("ref_ch3", st, ["solve_pcg:12:1e-4"], tmp_path)
    info = w.solve_nrpcg(12, 1e-4)
    assert info["gs_fallbacks"] > 0
    assert "PCG failed to converge" in r.stderr
    assert r.diag[0] == float(info["converged"])
    assert_bits(r.phi, w.phi, "phi NR-PCG with GS fallback")


@ref3
def test_cold_beam_sample_mt19937(tmp_path):
    """ColdBeamSource::sample + addParticle with the reference's own generator, reseeded (mt19937 +
    libstdc++ uniform_real_distribution)."""
    w, sp = cases.sphere_case(seed=10, n=50)
    st = sf.state_from_oracle(w, [sp], 1e-7)
    r = sf.run_ref("ref_ch3", st, ["sample:0:7000:1e10:12345:3"], tmp_path)
    g = orc.mt19937(12345)
    added = [sp.sample_cold_beam_mt(7000.0, 1e10, 1e-7, g) for _ in range(3)]
    assert all(a in (5600, 5601) for a in added)          # first injection = 5600 (+Bernoulli)
    assert_bits(r.species[0]["part"], sp.particles(), "injected particles")


@ref3
def test_update_average(tmp_path):
    w, sp = cases.sphere_case(seed=11, n=500)
    sp.compute_number_density()
    st = sf.state_from_oracle(w, [sp], 1e-7)
    r = sf.run_ref("ref_ch3", st, ["average:0", "advance", "deposit", "average:0", "average:0"], tmp_path)
    sp.update_averages()
    sp.advance(1e-7)
    sp.compute_number_density()
    sp.update_averages()
    sp.update_averages()
    assert_bits(r.species[0]["den_ave"], sp.den_ave, "den_ave")


@ref2
def test_ch2_step(tmp_path):
    """Grounded box: reflecting advance (ch2/Species.cpp:7-38), two species, linear GS (ch2/PotentialSolver.cpp:11-67)."""
    w, (ions, eles) = cases.box_case(seed=12)
    dt = 2e-9
    st = sf.state_from_oracle(w, [ions, eles], dt)
    r = sf.run_ref("ref_ch2", st, ["advance", "deposit", "rho", "solve:3000:1e-4", "ef", "advance"], tmp_path)
    before = eles.particles()
    for s in (ions, eles):
        s.advance_box(dt)
        s.compute_number_density()
    w.compute_charge_density([ions, eles])
    info = w.solve_gs_box(3000, 1e-4)
    w.compute_ef()
    for s in (ions, eles):
        s.advance_box(dt)
    assert (np.sign(before[3:6]) != np.sign(eles.particles()[3:6])).sum() > 20   # reflections happened
    assert r.diag[0] == float(info["converged"]) == 1.0
    assert_bits(r.phi, w.phi, "phi")
    assert_bits(r.ef, w.ef, "ef")
    assert_bits(r.rho, w.rho, "rho")
    for q, s in enumerate((ions, eles)):
        assert_bits(r.species[q]["part"], s.particles(), "particles sp%d" % q)
        assert_bits(r.species[q]["den"], s.den, "den sp%d" % q)


@ref2
def test_ch2_quiet_start(tmp_path):
    """loadParticlesBoxQS (ch2/Species.cpp:101-141): ions over the box, electrons over half of it."""
    w = cases.box_world(9)
    ions = orc.Species(w, 16 * AMU, QE)
    eles = orc.Species(w, ME, -QE)
    st = sf.state_from_oracle(w, [ions, eles], 2e-10)
    r = sf.run_ref("ref_ch2", st, ["loadqs:0:1e11:13:13:13:0", "loadqs:1:1e11:7:7:7:1", "deposit", "rho"], tmp_path)
    ions.load_box_qs(w.x0, w.xm, 1e11, (13, 13, 13), 2e-10)
    eles.load_box_qs(w.x0, w.xc, 1e11, (7, 7, 7), 2e-10)
    for s in (ions, eles):
        s.compute_number_density()
    w.compute_charge_density([ions, eles])
    assert ions.np == 13 ** 3 and eles.np == 7 ** 3
    assert_bits(r.species[0]["part"], ions.particles(), "ions")
    assert_bits(r.species[1]["part"], eles.particles(), "electrons")
    assert_bits(r.rho, w.rho, "rho")
