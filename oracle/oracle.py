"""ctypes front end of the CPU oracle (oracle/espic_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from the product package.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

EPS_0 = 8.85418782e-12
QE = 1.602176565e-19
AMU = 1.660538921e-27
ME = 9.10938215e-31


def build(force=False):
    src = os.path.join(HERE, "espic_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, src, "-lm"])
    return LIB


class CMesh(C.Structure):
    _fields_ = [("ni", C.c_int), ("nj", C.c_int), ("nk", C.c_int),
                ("x0", C.c_double * 3), ("xm", C.c_double * 3), ("dh", C.c_double * 3), ("xc", C.c_double * 3),
                ("sphere_c", C.c_double * 3), ("sphere_r2", C.c_double)]


class CParticles(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in ("x", "y", "z", "vx", "vy", "vz", "mpw")] + \
               [("np", C.c_int64), ("cap", C.c_int64)]


class CSolveInfo(C.Structure):
    _fields_ = [("converged", C.c_int), ("nr_iters", C.c_int), ("lin_calls", C.c_int), ("lin_iters", C.c_int64),
                ("gs_fallbacks", C.c_int), ("gs_iters", C.c_int64), ("residual", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class CSurfaceRng(C.Structure):
    _fields_ = [("mode", C.c_int), ("mt", C.c_void_p), ("seed", C.c_uint64), ("stream", C.c_uint32), ("step", C.c_uint32)]


class CSurfaceTarget(C.Structure):
    _fields_ = [("p", C.POINTER(CParticles)), ("pdt", C.POINTER(C.c_double)),
                ("charge", C.c_double), ("mass", C.c_double), ("mpw0", C.c_double)]


class CMT(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        L.orc_mt_uniform.restype = C.c_double
        L.orc_mt_next.restype = C.c_uint32
        L.orc_real_count.restype = C.c_double
        L.orc_ke.restype = C.c_double
        L.orc_pe.restype = C.c_double
        L.orc_advance_sphere.restype = C.c_int64
        L.orc_cold_beam_num_sim.restype = C.c_int64
        L.orc_cold_beam_sample_mt.restype = C.c_int64
        L.orc_warm_beam_sample_mt.restype = C.c_int64
        L.orc_warm_beam_sample_philox.restype = C.c_int64
        L.orc_cold_beam_sample_philox.restype = C.c_int64
        L.orc_load_box_qs.restype = C.c_int64
        L.orc_cold_beam_num_sim.argtypes = [C.POINTER(CMesh)] + [C.c_double] * 5
        L.orc_advance_sphere.argtypes = [C.POINTER(CMesh), dp, C.POINTER(CParticles), C.c_double, C.c_double, C.c_double]
        L.orc_advance_box.argtypes = L.orc_advance_sphere.argtypes
        L.orc_push_sphere_nocompact.argtypes = L.orc_advance_sphere.argtypes
        L.orc_number_density.argtypes = [C.POINTER(CMesh), dp, C.POINTER(CParticles), dp]
        L.orc_sample_moments.argtypes = [C.POINTER(CMesh), C.POINTER(CParticles), dp, dp, dp, dp, dp]
        L.orc_gas_properties.argtypes = [C.POINTER(CMesh), C.c_double, dp, dp, dp, dp, dp, dp, dp]
        L.orc_rho_add.argtypes = [C.POINTER(CMesh), dp, dp, C.c_double]
        L.orc_momentum.argtypes = [C.POINTER(CParticles), C.c_double, dp]
        L.orc_ke.argtypes = [C.POINTER(CParticles), C.c_double]
        L.orc_add_sphere.argtypes = [C.POINTER(CMesh), dp, C.c_double, C.c_double, ip, dp]
        L.orc_solve_qn.argtypes = [C.POINTER(CMesh), ip, dp, dp, C.c_double, C.c_double, C.c_double]
        L.orc_solve_gs.argtypes = [C.POINTER(CMesh), ip, dp, dp, C.c_double, C.c_double, C.c_double,
                                   C.c_uint, C.c_double, C.POINTER(CSolveInfo)]
        L.orc_solve_gs_box.argtypes = [C.POINTER(CMesh), dp, dp, C.c_uint, C.c_double, C.POINTER(CSolveInfo)]
        L.orc_solve_nrpcg.argtypes = [C.POINTER(CMesh), ip, dp, dp, C.c_double, C.c_double, C.c_double,
                                      C.c_uint, C.c_double, C.c_int, C.c_double, C.POINTER(CSolveInfo)]
        L.orc_cold_beam_sample_mt.argtypes = [C.POINTER(CMesh), dp, C.POINTER(CParticles)] + [C.c_double] * 6 + [C.POINTER(CMT)]
        L.orc_warm_beam_sample_mt.argtypes = [C.POINTER(CMesh), dp, C.POINTER(CParticles)] + [C.c_double] * 7 + [C.POINTER(CMT)]
        L.orc_warm_beam_sample_philox.argtypes = [C.POINTER(CMesh), dp, C.POINTER(CParticles)] + [C.c_double] * 7 + \
            [C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_cold_beam_sample_philox.argtypes = [C.POINTER(CMesh), dp, C.POINTER(CParticles)] + [C.c_double] * 6 + \
                                                 [C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_load_box_qs.argtypes = [C.POINTER(CMesh), dp, C.POINTER(CParticles), dp, dp, C.c_double,
                                      C.POINTER(C.c_int), C.c_double, C.c_double, C.c_double]
        L.orc_philox_uniform2.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, dp]
        L.orc_add_particle.argtypes = [C.POINTER(CMesh), dp, C.POINTER(CParticles), dp, dp] + [C.c_double] * 4
        L.orc_advance_surface.restype = C.c_int64
        L.orc_advance_surface.argtypes = [C.POINTER(CMesh), dp, C.POINTER(CParticles), dp] + [C.c_double] * 4 + \
            [C.POINTER(CSurfaceTarget), C.POINTER(CSurfaceTarget), C.POINTER(CSurfaceRng), C.POINTER(C.c_int64)]
        L.orc_dsmc_mex.restype = C.c_int64
        L.orc_dsmc_mex.argtypes = [C.POINTER(CMesh), C.POINTER(CParticles), C.c_double, C.c_double, C.c_double, dp,
                                   C.POINTER(CSurfaceRng)]
        L.orc_compute_mpc.argtypes = [C.POINTER(CMesh), C.POINTER(CParticles), dp]
        L.orc_vhs_sigma.restype = C.c_double
        L.orc_vhs_sigma.argtypes = [C.c_double, C.c_double]
        L.orc_mcc_cex.restype = C.c_int64
        L.orc_mcc_cex.argtypes = [C.POINTER(CMesh), C.POINTER(CParticles), dp, dp, dp, C.c_double, C.c_double,
                                  C.POINTER(CSurfaceRng)]
        L.orc_line_sphere_intersect.restype = C.c_double
        L.orc_line_sphere_intersect.argtypes = [C.POINTER(CMesh), dp, dp]
        _lib = L
    return _lib


def _dp(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class Species:
    """SoA particle store + density fields of one species."""

    def __init__(self, world, mass, charge, mpw0=1.0, cap=1024):
        self.world = world
        self.mass, self.charge, self.mpw0 = float(mass), float(charge), float(mpw0)
        self.cap = 0
        self.np = 0
        self.arr = np.zeros((7, 0))
        self.pdt = np.zeros(0)              # ch4 Particle::dt (Species.h:15), used by advance_surface only
        self.reserve(cap)
        self.den = np.zeros(world.nn)
        self.den_ave = np.zeros(world.nn)
        self.ave_samples = C.c_int(0)

    def reserve(self, cap):
        if cap <= self.cap:
            return
        new = np.zeros((7, cap))
        new[:, :self.np] = self.arr[:, :self.np]
        pdt = np.zeros(cap)
        pdt[:self.np] = self.pdt[:self.np]
        self.arr, self.pdt, self.cap = new, pdt, cap

    def set_particles(self, soa):
        soa = np.ascontiguousarray(soa, dtype=np.float64)
        n = soa.shape[1]
        self.reserve(max(n, 1))
        self.arr[:, :n] = soa
        self.np = n

    def particles(self):
        return self.arr[:, :self.np].copy()

    def _c(self):
        p = CParticles()
        for c, name in enumerate(("x", "y", "z", "vx", "vy", "vz", "mpw")):
            setattr(p, name, _dp(self.arr[c]))
        p.np, p.cap = self.np, self.cap
        return p

    # --- reference Species API
    def advance(self, dt):
        p = self._c()
        self.np = lib().orc_advance_sphere(C.byref(self.world.m), _dp(self.world.ef), C.byref(p), self.charge, self.mass, dt)

    def advance_surface(self, dt, neutrals, sput, rng, headroom=None):
        """ch4 Species::advance(neutrals, spherium).  rng = ("mt", CMT) or ("philox", seed, stream, step).
        Returns (emitted into neutrals, emitted into sput)."""
        targets = [neutrals] if sput is neutrals else [neutrals, sput]
        for t in targets:
            if t is not self:
                t.reserve(t.np + (headroom if headroom is not None else 4 * self.np + 64))
        p = self._c()
        cps = {id(self): p}
        for t in targets:
            if id(t) not in cps:
                cps[id(t)] = t._c()
        def target(t):
            return CSurfaceTarget(C.pointer(cps[id(t)]), _dp(t.pdt), t.charge, t.mass, t.mpw0)
        tn, ts = target(neutrals), target(sput)
        r = CSurfaceRng()
        if rng[0] == "mt":
            r.mode, r.mt = 0, C.cast(C.pointer(rng[1]), C.c_void_p)
        else:
            r.mode, r.seed, r.stream, r.step = 1, rng[1], rng[2], rng[3]
        em = (C.c_int64 * 2)()
        self.np = lib().orc_advance_surface(C.byref(self.world.m), _dp(self.world.ef), C.byref(p), _dp(self.pdt),
                                            self.charge, self.mass, self.mpw0, dt, C.byref(tn), C.byref(ts), C.byref(r), em)
        for t in targets:
            t.np = cps[id(t)].np
        return int(em[0]), int(em[1])

    @staticmethod
    def _rng(rng):
        r = CSurfaceRng()
        if rng[0] == "mt":
            r.mode, r.mt = 0, C.cast(C.pointer(rng[1]), C.c_void_p)
        else:
            r.mode, r.seed, r.stream, r.step = 1, rng[1], rng[2], rng[3]
        return r

    def dsmc_mex(self, dt, sigma_cr_max, rng):
        """ch4 DSMC_MEX::apply on this species; returns (collisions, new sigma_cr_max)"""
        p = self._c()
        s = np.array([sigma_cr_max], dtype=np.float64)
        r = self._rng(rng)
        cols = lib().orc_dsmc_mex(C.byref(self.world.m), C.byref(p), self.mass, self.mpw0, dt, _dp(s), C.byref(r))
        return int(cols), float(s[0])

    def mcc_cex(self, target_den, target_vel, target_T, target_mass, dt, rng):
        """ch4 MCC_CEX::apply with this species as the source; returns the number of collisions"""
        p = self._c()
        r = self._rng(rng)
        return int(lib().orc_mcc_cex(C.byref(self.world.m), C.byref(p), _dp(target_den), _dp(target_vel), _dp(target_T),
                                     target_mass, dt, C.byref(r)))

    def compute_mpc(self):
        w = self.world
        mpc = np.zeros((w.ni - 1) * (w.nj - 1) * (w.nk - 1))
        p = self._c()
        lib().orc_compute_mpc(C.byref(w.m), C.byref(p), _dp(mpc))
        return mpc

    def push_nocompact(self, dt):
        p = self._c()
        lib().orc_push_sphere_nocompact(C.byref(self.world.m), _dp(self.world.ef), C.byref(p), self.charge, self.mass, dt)

    def advance_box(self, dt):
        p = self._c()
        lib().orc_advance_box(C.byref(self.world.m), _dp(self.world.ef), C.byref(p), self.charge, self.mass, dt)

    def compute_number_density(self):
        p = self._c()
        lib().orc_number_density(C.byref(self.world.m), _dp(self.world.node_vol), C.byref(p), _dp(self.den))

    # ---- velocity moments (ch4/Species.cpp:190-241)
    def _moment_arrays(self):
        if not hasattr(self, "n_sum"):
            nn = self.world.nn
            self.n_sum, self.nv_sum = np.zeros(nn), np.zeros(3 * nn)
            self.nuu_sum, self.nvv_sum, self.nww_sum = np.zeros(nn), np.zeros(nn), np.zeros(nn)
            self.vel, self.T = np.zeros(3 * nn), np.zeros(nn)

    def clear_samples(self):
        self._moment_arrays()
        for a in (self.n_sum, self.nv_sum, self.nuu_sum, self.nvv_sum, self.nww_sum):
            a[:] = 0

    def sample_moments(self):
        self._moment_arrays()
        p = self._c()
        lib().orc_sample_moments(C.byref(self.world.m), C.byref(p), _dp(self.n_sum), _dp(self.nv_sum), _dp(self.nuu_sum),
                                 _dp(self.nvv_sum), _dp(self.nww_sum))

    def compute_gas_properties(self):
        self._moment_arrays()
        lib().orc_gas_properties(C.byref(self.world.m), C.c_double(self.mass), _dp(self.n_sum), _dp(self.nv_sum), _dp(self.nuu_sum),
                                 _dp(self.nvv_sum), _dp(self.nww_sum), _dp(self.vel), _dp(self.T))

    def update_averages(self):
        lib().orc_update_average(C.byref(self.world.m), _dp(self.den_ave), _dp(self.den), C.byref(self.ave_samples))

    def real_count(self):
        p = self._c()
        return lib().orc_real_count(C.byref(p))

    def momentum(self):
        p = self._c()
        out = np.zeros(3)
        lib().orc_momentum(C.byref(p), self.mass, _dp(out))
        return out

    def ke(self):
        p = self._c()
        return lib().orc_ke(C.byref(p), self.mass)

    def add_particle(self, pos, vel, mpw, dt):
        self.reserve(max(self.np + 1, 2 * self.cap if self.np + 1 > self.cap else self.cap))
        p = self._c()
        pos = np.asarray(pos, dtype=np.float64)
        vel = np.asarray(vel, dtype=np.float64)
        r = lib().orc_add_particle(C.byref(self.world.m), _dp(self.world.ef), C.byref(p), _dp(pos), _dp(vel),
                                   float(mpw), self.charge, self.mass, dt)
        self.np = p.np
        return r

    def sample_cold_beam_mt(self, v_drift, den, dt, mt):
        n_max = int(lib().orc_cold_beam_num_sim(C.byref(self.world.m), den, v_drift, dt, self.mpw0, 1.0)) + 1
        self.reserve(self.np + n_max)
        p = self._c()
        added = lib().orc_cold_beam_sample_mt(C.byref(self.world.m), _dp(self.world.ef), C.byref(p), self.charge, self.mass,
                                              self.mpw0, v_drift, den, dt, C.byref(mt))
        self.np = p.np
        return added

    def sample_cold_beam_philox(self, v_drift, den, dt, seed, stream, step):
        n_max = int(lib().orc_cold_beam_num_sim(C.byref(self.world.m), den, v_drift, dt, self.mpw0, 1.0)) + 1
        self.reserve(self.np + n_max)
        p = self._c()
        added = lib().orc_cold_beam_sample_philox(C.byref(self.world.m), _dp(self.world.ef), C.byref(p), self.charge,
                                                  self.mass, self.mpw0, v_drift, den, dt, seed, stream, step)
        self.np = p.np
        return added

    def sample_warm_beam_mt(self, v_drift, den, T, dt, mt):
        n_max = int(lib().orc_cold_beam_num_sim(C.byref(self.world.m), den, v_drift, dt, self.mpw0, 1.0)) + 1
        self.reserve(self.np + n_max)
        p = self._c()
        added = lib().orc_warm_beam_sample_mt(C.byref(self.world.m), _dp(self.world.ef), C.byref(p), self.charge, self.mass,
                                              self.mpw0, v_drift, den, T, dt, C.byref(mt))
        self.np = p.np
        return added

    def sample_warm_beam_philox(self, v_drift, den, T, dt, seed, stream, step):
        n_max = int(lib().orc_cold_beam_num_sim(C.byref(self.world.m), den, v_drift, dt, self.mpw0, 1.0)) + 1
        self.reserve(self.np + n_max)
        p = self._c()
        added = lib().orc_warm_beam_sample_philox(C.byref(self.world.m), _dp(self.world.ef), C.byref(p), self.charge,
                                                  self.mass, self.mpw0, v_drift, den, T, dt, seed, stream, step)
        self.np = p.np
        return added

    def load_box_qs(self, x1, x2, num_den, grid, dt):
        n = int(grid[0]) * int(grid[1]) * int(grid[2])
        self.reserve(self.np + n)
        p = self._c()
        x1 = np.asarray(x1, dtype=np.float64)
        x2 = np.asarray(x2, dtype=np.float64)
        g = (C.c_int * 3)(*[int(v) for v in grid])
        added = lib().orc_load_box_qs(C.byref(self.world.m), _dp(self.world.ef), C.byref(p), _dp(x1), _dp(x2),
                                      num_den, g, self.charge, self.mass, dt)
        self.np = p.np
        return added


class World:
    """Flat-array counterpart of the reference World (+ PotentialSolver entry points)."""

    def __init__(self, ni, nj, nk, x0, xm):
        self.m = CMesh()
        x0a = (C.c_double * 3)(*x0)
        xma = (C.c_double * 3)(*xm)
        lib().orc_mesh_init(C.byref(self.m), ni, nj, nk, x0a, xma)
        self.ni, self.nj, self.nk = ni, nj, nk
        self.nn = ni * nj * nk
        self.phi = np.zeros(self.nn)
        self.rho = np.zeros(self.nn)
        self.ef = np.zeros(3 * self.nn)
        self.node_vol = np.zeros(self.nn)
        self.object_id = np.zeros(self.nn, dtype=np.int32)
        lib().orc_node_volumes(C.byref(self.m), _dp(self.node_vol))
        self.phi0, self.Te0, self.n0 = 0.0, 1.5, 1e12
        self.sphere = None
        self.inlet = False

    @property
    def dh(self):
        return np.array(self.m.dh[:])

    @property
    def x0(self):
        return np.array(self.m.x0[:])

    @property
    def xm(self):
        return np.array(self.m.xm[:])

    @property
    def xc(self):
        return np.array(self.m.xc[:])

    def add_sphere(self, c, radius, phi_sphere):
        ca = np.asarray(c, dtype=np.float64)
        lib().orc_add_sphere(C.byref(self.m), _dp(ca), radius, phi_sphere, _ip(self.object_id), _dp(self.phi))
        self.sphere = (tuple(float(v) for v in c), float(radius), float(phi_sphere))

    def add_inlet(self):
        lib().orc_add_inlet(C.byref(self.m), _ip(self.object_id), _dp(self.phi))
        self.inlet = True

    def set_reference_values(self, phi0, Te0, n0):
        self.phi0, self.Te0, self.n0 = float(phi0), float(Te0), float(n0)

    def compute_charge_density(self, species):
        lib().orc_rho_clear(C.byref(self.m), _dp(self.rho))
        for sp in species:
            lib().orc_rho_add(C.byref(self.m), _dp(self.rho), _dp(sp.den), sp.charge)

    def solve_qn(self):
        lib().orc_solve_qn(C.byref(self.m), _ip(self.object_id), _dp(self.rho), _dp(self.phi), self.phi0, self.Te0, self.n0)

    def solve_gs(self, max_it, tol):
        info = CSolveInfo()
        lib().orc_solve_gs(C.byref(self.m), _ip(self.object_id), _dp(self.rho), _dp(self.phi), self.phi0, self.Te0, self.n0,
                           max_it, tol, C.byref(info))
        return info.as_dict()

    def solve_gs_box(self, max_it, tol):
        info = CSolveInfo()
        lib().orc_solve_gs_box(C.byref(self.m), _dp(self.rho), _dp(self.phi), max_it, tol, C.byref(info))
        return info.as_dict()

    def solve_nrpcg(self, max_it, tol, nr_max_it=20, nr_tol=1e-3):
        info = CSolveInfo()
        lib().orc_solve_nrpcg(C.byref(self.m), _ip(self.object_id), _dp(self.rho), _dp(self.phi), self.phi0, self.Te0, self.n0,
                              max_it, tol, nr_max_it, nr_tol, C.byref(info))
        return info.as_dict()

    def compute_ef(self):
        lib().orc_compute_ef(C.byref(self.m), _dp(self.phi), _dp(self.ef))

    def pe(self):
        return lib().orc_pe(C.byref(self.m), _dp(self.ef), _dp(self.node_vol))


def mt19937(seed):
    g = CMT()
    lib().orc_mt_seed(C.byref(g), C.c_uint32(seed))
    return g


def philox_uniform2(seed, stream, step, idx):
    out = np.zeros(2)
    lib().orc_philox_uniform2(seed, stream, step, idx, _dp(out))
    return out
