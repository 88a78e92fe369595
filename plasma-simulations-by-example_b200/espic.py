"""ctypes binding of libespic_cuda.so (include/espic.h) -- the Python-side caller of the C ABI.

Used by the GPU parity tests, bench.py and __graft_entry__.smoke().  There is no CPU fallback: loading fails
loudly if the CUDA library has not been built (python __graft_entry__.py build), and Engine() raises if the
library reports an error (e.g. no CUDA device).
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libespic_cuda.so")

PHI, RHO, EF, NODE_VOL, OBJECT_ID, DEN, DEN_AVE, VEL, T, N_SUM, NV_SUM, NUU_SUM, NVV_SUM, NWW_SUM, MPC = range(15)
WALL_ABSORB, WALL_REFLECT = 0, 1
PUSH_FUSE_DEPOSIT, PUSH_NO_COMPACT, PUSH_MIGRATE, PUSH_DIAG, PUSH_FIXED_POINT = 1, 2, 4, 8, 256
DEPOSIT_FP64, DEPOSIT_FIXED = 0, 1
SORT_XTOC, SORT_DRIFT_Z = 0, 1
SOLVE_GS, SOLVE_PCG, SOLVE_QN, SOLVE_GS_BOX, SOLVE_PCG_REF, SOLVE_PCG_MG, SOLVE_PCG_MG_SLAB = 0, 1, 2, 3, 4, 5, 6

EXPORTS = [
    "espic_create", "espic_destroy", "espic_last_error", "espic_set_stream", "espic_sync", "espic_kernel_launches",
    "espic_get_mesh", "espic_add_sphere", "espic_add_inlet", "espic_field_download", "espic_field_download_async", "espic_copy_sync", "espic_field_upload",
    "espic_field_devptr", "espic_species_create", "espic_species_reserve", "espic_species_count",
    "espic_species_upload", "espic_species_download", "espic_species_upload_device", "espic_species_add", "espic_species_prefetch", "espic_push", "espic_last_push_ms", "espic_deposit",
    "espic_sort_by_cell", "espic_sort_particles", "espic_inject_cold_beam", "espic_inject_warm_beam", "espic_push_surface", "espic_dsmc_mex", "espic_mcc_cex", "espic_compute_mpc", "espic_species_diag", "espic_update_average", "espic_sample_moments", "espic_compute_gas_properties", "espic_clear_samples",
    "espic_charge_density", "espic_solve", "espic_mg_plan", "espic_compute_ef", "espic_field_pe", "espic_comm_unique_id",
    "espic_comm_init",
    "espic_domain_set", "espic_domain_get", "espic_migrate", "espic_migrate_pack", "espic_migrate_segment", "espic_migrate_finish",
]


class SolveParams(C.Structure):
    _fields_ = [("type", C.c_int), ("max_it", C.c_int), ("tol", C.c_double), ("phi0", C.c_double), ("Te0", C.c_double),
                ("n0", C.c_double), ("nr_max_it", C.c_int), ("nr_tol", C.c_double)]


class SolveInfo(C.Structure):
    _fields_ = [("converged", C.c_int), ("nr_iters", C.c_int), ("lin_iters", C.c_longlong), ("gs_fallbacks", C.c_int),
                ("gs_iters", C.c_longlong), ("residual", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class EspicError(RuntimeError):
    pass


_lib = None


def load():
    """Load libespic_cuda.so; raises (never falls back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EspicError("CUDA extension %s is not built: run `python __graft_entry__.py build`" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    dp, vp = C.POINTER(C.c_double), C.c_void_p
    comp = C.POINTER(dp)
    L.espic_last_error.restype = C.c_char_p
    L.espic_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, dp, dp, C.c_int]
    L.espic_destroy.argtypes = [vp]
    L.espic_destroy.restype = None
    L.espic_set_stream.argtypes = [vp, vp]
    L.espic_sync.argtypes = [vp]
    L.espic_kernel_launches.argtypes = [vp]
    L.espic_kernel_launches.restype = C.c_longlong
    L.espic_get_mesh.argtypes = [vp, dp, dp]
    L.espic_add_sphere.argtypes = [vp, dp, C.c_double, C.c_double]
    L.espic_add_inlet.argtypes = [vp]
    L.espic_field_download.argtypes = [vp, C.c_int, C.c_int, vp]
    L.espic_field_download_async.argtypes = [vp, C.c_int, C.c_int, vp]
    L.espic_copy_sync.argtypes = [vp]
    L.espic_species_prefetch.argtypes = [vp, C.c_int, comp, C.c_longlong]
    L.espic_field_upload.argtypes = [vp, C.c_int, C.c_int, vp]
    L.espic_field_devptr.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.espic_species_create.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_longlong]
    L.espic_species_reserve.argtypes = [vp, C.c_int, C.c_longlong]
    L.espic_species_count.argtypes = [vp, C.c_int]
    L.espic_species_count.restype = C.c_longlong
    L.espic_species_upload.argtypes = [vp, C.c_int, comp, C.c_longlong, C.c_int]
    L.espic_species_upload_device.argtypes = [vp, C.c_int, comp, C.c_longlong, C.c_double, C.c_int]
    L.espic_species_download.argtypes = [vp, C.c_int, comp, C.c_longlong]
    L.espic_species_download.restype = C.c_longlong
    L.espic_species_add.argtypes = [vp, C.c_int, comp, C.c_longlong, C.c_double, C.POINTER(C.c_longlong)]
    L.espic_push.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_int]
    L.espic_last_push_ms.argtypes = [vp, C.POINTER(C.c_double)]
    L.espic_deposit.argtypes = [vp, C.c_int, C.c_int]
    L.espic_sort_by_cell.argtypes = [vp, C.c_int]
    L.espic_sort_particles.argtypes = [vp, C.c_int, C.c_int]
    L.espic_inject_warm_beam.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint32,
                                         C.c_uint32, C.POINTER(C.c_longlong)]
    L.espic_inject_cold_beam.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint32,
                                         C.c_uint32, C.POINTER(C.c_longlong)]
    L.espic_push_surface.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32,
                                     C.POINTER(C.c_longlong)]
    L.espic_dsmc_mex.argtypes = [vp, C.c_int, C.c_double, dp, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_longlong)]
    L.espic_compute_mpc.argtypes = [vp, C.c_int]
    L.espic_mcc_cex.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_longlong)]
    L.espic_species_diag.argtypes = [vp, C.c_int, dp]
    L.espic_update_average.argtypes = [vp, C.c_int]
    L.espic_sample_moments.argtypes = [vp, C.c_int]
    L.espic_compute_gas_properties.argtypes = [vp, C.c_int]
    L.espic_clear_samples.argtypes = [vp, C.c_int]
    L.espic_charge_density.argtypes = [vp]
    L.espic_solve.argtypes = [vp, C.POINTER(SolveParams), C.POINTER(SolveInfo)]
    L.espic_mg_plan.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.espic_compute_ef.argtypes = [vp]
    L.espic_field_pe.argtypes = [vp, dp]
    L.espic_comm_unique_id.argtypes = [vp]
    L.espic_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    ip, llp = C.POINTER(C.c_int), C.POINTER(C.c_longlong)
    L.espic_domain_set.argtypes = [vp, C.c_int, C.c_int, ip]
    L.espic_domain_get.argtypes = [vp, ip, ip, ip]
    L.espic_migrate.argtypes = [vp, C.c_int, llp, llp]
    L.espic_migrate_pack.argtypes = [vp, C.c_int, llp]
    L.espic_migrate_segment.argtypes = [vp, C.c_int, C.POINTER(vp), llp]
    L.espic_migrate_finish.argtypes = [vp, C.c_int]
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def mg_plan(ni, nj, nk, dh, nranks=1):
    """espic_mg_plan: (level dimensions, first level solved redundantly by every rank, fine k-planes per coarsest plane).
    Host-side planning only: works without a GPU."""
    L = load()
    dims = (C.c_longlong * 24)()
    fr, unit = C.c_int(0), C.c_int(0)
    d = (C.c_double * 3)(*[float(x) for x in dh])
    nlev = L.espic_mg_plan(int(ni), int(nj), int(nk), d, int(nranks), dims, C.byref(fr), C.byref(unit))
    if nlev < 0:
        raise EspicError(L.espic_last_error().decode())
    return [tuple(dims[3 * l + a] for a in range(3)) for l in range(nlev)], fr.value, unit.value


def _comp(arr7):
    """(7,n) C-contiguous float64 -> double*[7]"""
    ptrs = (C.POINTER(C.c_double) * 7)()
    for q in range(7):
        ptrs[q] = _dp(arr7[q])
    return ptrs


class Engine:
    """One World on one GPU (reference World + PotentialSolver entry points); species are integer ids."""

    def __init__(self, ni, nj, nk, x0, xm, device=0):
        self.L = load()
        self.h = C.c_void_p()
        x0a = np.asarray(x0, dtype=np.float64)
        xma = np.asarray(xm, dtype=np.float64)
        r = self.L.espic_create(C.byref(self.h), ni, nj, nk, _dp(x0a), _dp(xma), device)
        if r:
            raise EspicError(self.L.espic_last_error().decode())
        self.ni, self.nj, self.nk = ni, nj, nk
        self.nn = ni * nj * nk
        self.phi0, self.Te0, self.n0 = 0.0, 1.5, 1e12
        self.nr_max_it, self.nr_tol = 20, 1e-3

    def close(self):
        if self.h:
            self.L.espic_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, r):
        if r < 0:
            raise EspicError(self.L.espic_last_error().decode())
        return r

    # ---- plumbing
    def set_stream(self, cuda_stream):
        self._ck(self.L.espic_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sync(self):
        self._ck(self.L.espic_sync(self.h))

    def kernel_launches(self):
        return int(self.L.espic_kernel_launches(self.h))

    def mesh(self):
        dh, xc = np.zeros(3), np.zeros(3)
        self._ck(self.L.espic_get_mesh(self.h, _dp(dh), _dp(xc)))
        return dh, xc

    # ---- World
    def add_sphere(self, c, radius, phi_sphere):
        ca = np.asarray(c, dtype=np.float64)
        self._ck(self.L.espic_add_sphere(self.h, _dp(ca), radius, phi_sphere))

    def add_inlet(self):
        self._ck(self.L.espic_add_inlet(self.h))

    def field(self, which, sp=0, out=None):
        """Download a node field.  `out`: optional preallocated (e.g. pinned) array to receive it."""
        n, dt = self.nn * (3 if which in (EF, VEL, NV_SUM) else 1), (np.int32 if which == OBJECT_ID else np.float64)
        if out is None:
            out = np.zeros(n, dtype=dt)
        assert out.size == n and out.dtype == dt and out.flags["C_CONTIGUOUS"]
        self._ck(self.L.espic_field_download(self.h, which, sp, out.ctypes.data_as(C.c_void_p)))
        return out

    def field_async(self, which, out, sp=0):
        """Start the download of a node field into `out` (pinned memory) on the copy stream; copy_sync() waits for it."""
        n, dt = self.nn * (3 if which in (EF, VEL, NV_SUM) else 1), (np.int32 if which == OBJECT_ID else np.float64)
        assert out.size == n and out.dtype == dt and out.flags["C_CONTIGUOUS"]
        self._ck(self.L.espic_field_download_async(self.h, which, sp, out.ctypes.data_as(C.c_void_p)))
        return out

    def copy_sync(self):
        self._ck(self.L.espic_copy_sync(self.h))

    def set_field(self, which, arr, sp=0):
        arr = np.ascontiguousarray(arr, dtype=np.int32 if which == OBJECT_ID else np.float64)
        assert arr.size == self.nn * (3 if which in (EF, VEL, NV_SUM) else 1)
        self._ck(self.L.espic_field_upload(self.h, which, sp, arr.ctypes.data_as(C.c_void_p)))

    def field_devptr(self, which, sp=0):
        p = C.c_void_p()
        self._ck(self.L.espic_field_devptr(self.h, which, sp, C.byref(p)))
        return p.value

    def set_reference_values(self, phi0, Te0, n0):
        self.phi0, self.Te0, self.n0 = float(phi0), float(Te0), float(n0)

    def compute_charge_density(self):
        self._ck(self.L.espic_charge_density(self.h))

    def solve(self, solver, max_it, tol):
        p = SolveParams(solver, int(max_it), float(tol), self.phi0, self.Te0, self.n0, self.nr_max_it, self.nr_tol)
        info = SolveInfo()
        self._ck(self.L.espic_solve(self.h, C.byref(p), C.byref(info)))
        return info.as_dict()

    def compute_ef(self):
        self._ck(self.L.espic_compute_ef(self.h))

    def pe(self):
        out = np.zeros(1)
        self._ck(self.L.espic_field_pe(self.h, _dp(out)))
        return float(out[0])

    # ---- Species
    def add_species(self, mass, charge, mpw0=1.0, capacity=1024):
        return self._ck(self.L.espic_species_create(self.h, mass, charge, mpw0, int(capacity)))

    def reserve(self, sp, capacity):
        self._ck(self.L.espic_species_reserve(self.h, sp, int(capacity)))

    def count(self, sp):
        return int(self._ck(self.L.espic_species_count(self.h, sp)))

    def upload(self, sp, soa, append=False):
        soa = np.ascontiguousarray(soa, dtype=np.float64)
        assert soa.shape[0] == 7
        self._ck(self.L.espic_species_upload(self.h, sp, _comp(soa), soa.shape[1], int(append)))

    def upload_device(self, sp, dev_ptrs, n, mpw_max, append=False):
        """dev_ptrs: 7 device addresses (e.g. torch tensor.data_ptr()) of float64 arrays with n elements"""
        ptrs = (C.POINTER(C.c_double) * 7)()
        for q in range(7):
            ptrs[q] = C.cast(C.c_void_p(int(dev_ptrs[q])), C.POINTER(C.c_double))
        self._ck(self.L.espic_species_upload_device(self.h, sp, ptrs, int(n), float(mpw_max), int(append)))

    def download(self, sp):
        n = self.count(sp)
        out = np.zeros((7, n))
        got = self._ck(self.L.espic_species_download(self.h, sp, _comp(out), n))
        assert got == n
        return out

    def add_particles(self, sp, soa, dt):
        soa = np.ascontiguousarray(soa, dtype=np.float64)
        added = C.c_longlong(0)
        self._ck(self.L.espic_species_add(self.h, sp, _comp(soa), soa.shape[1], dt, C.byref(added)))
        return added.value

    def prefetch_particles(self, sp, soa):
        """Start the host -> device copy of the candidates a later add_particles(sp, soa, dt) will admit (soa: the same pinned array)."""
        assert soa.dtype == np.float64 and soa.flags["C_CONTIGUOUS"]
        self._ck(self.L.espic_species_prefetch(self.h, sp, _comp(soa), soa.shape[1]))

    def push(self, sp, dt, wall=WALL_ABSORB, flags=0):
        self._ck(self.L.espic_push(self.h, sp, dt, wall, flags))

    def last_push_ms(self):
        ms = C.c_double(0)
        self._ck(self.L.espic_last_push_ms(self.h, C.byref(ms)))
        return ms.value

    def deposit(self, sp, mode=DEPOSIT_FP64):
        self._ck(self.L.espic_deposit(self.h, sp, mode))

    def sort_particles(self, sp, order):
        self._ck(self.L.espic_sort_particles(self.h, sp, order))

    def sort_by_cell(self, sp):
        self._ck(self.L.espic_sort_by_cell(self.h, sp))

    def inject_cold_beam(self, sp, v_drift, den, dt, seed, stream, step):
        added = C.c_longlong(0)
        self._ck(self.L.espic_inject_cold_beam(self.h, sp, v_drift, den, dt, seed, stream, step, C.byref(added)))
        return added.value

    def inject_warm_beam(self, sp, v_drift, den, T, dt, seed, stream, step):
        added = C.c_longlong(0)
        self._ck(self.L.espic_inject_warm_beam(self.h, sp, v_drift, den, T, dt, seed, stream, step, C.byref(added)))
        return added.value

    def push_surface(self, sp, dt, neutrals_sp, sput_sp, seed, stream, step):
        """ch4 Species::advance(neutrals, spherium); returns (emitted into neutrals_sp, emitted into sput_sp)"""
        em = (C.c_longlong * 2)()
        self._ck(self.L.espic_push_surface(self.h, sp, dt, neutrals_sp, sput_sp, seed, stream, step, em))
        return int(em[0]), int(em[1])

    def dsmc_mex(self, sp, dt, sigma_cr_max, seed, stream, step):
        """ch4 DSMC_MEX::apply; returns (collisions, new sigma_cr_max)"""
        s = np.array([sigma_cr_max], dtype=np.float64)
        cols = C.c_longlong(0)
        self._ck(self.L.espic_dsmc_mex(self.h, sp, dt, _dp(s), seed, stream, step, C.byref(cols)))
        return cols.value, float(s[0])

    def mcc_cex(self, source_sp, target_sp, dt, seed, stream, step):
        cols = C.c_longlong(0)
        self._ck(self.L.espic_mcc_cex(self.h, source_sp, target_sp, dt, seed, stream, step, C.byref(cols)))
        return cols.value

    def compute_mpc(self, sp):
        self._ck(self.L.espic_compute_mpc(self.h, sp))
        out = np.zeros((self.ni - 1) * (self.nj - 1) * (self.nk - 1))
        self._ck(self.L.espic_field_download(self.h, MPC, sp, out.ctypes.data_as(C.c_void_p)))
        return out

    def diag(self, sp):
        out = np.zeros(5)
        self._ck(self.L.espic_species_diag(self.h, sp, _dp(out)))
        return out

    def sample_moments(self, sp):
        self._ck(self.L.espic_sample_moments(self.h, sp))

    def compute_gas_properties(self, sp):
        self._ck(self.L.espic_compute_gas_properties(self.h, sp))

    def clear_samples(self, sp):
        self._ck(self.L.espic_clear_samples(self.h, sp))

    def update_average(self, sp):
        self._ck(self.L.espic_update_average(self.h, sp))

    # ---- multi GPU
    def unique_id(self):
        buf = C.create_string_buffer(128)
        self._ck(self.L.espic_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank, nranks, uid):
        self._ck(self.L.espic_comm_init(self.h, rank, nranks, C.c_char_p(uid)))


    # ---- spatial decomposition with particle migration (ch9/MPI initMPIDomain / transferParticles)
    def set_domain(self, parts, part, k_bounds):
        kb = (C.c_int * (parts + 1))(*[int(k) for k in k_bounds])
        self._ck(self.L.espic_domain_set(self.h, parts, part, kb))
        self.parts, self.part = parts, part

    def migrate(self, sp):
        """collective over the communicator: returns (sent, received)"""
        a, b = C.c_longlong(0), C.c_longlong(0)
        self._ck(self.L.espic_migrate(self.h, sp, C.byref(a), C.byref(b)))
        return a.value, b.value

    def migrate_pack(self, sp):
        counts = (C.c_longlong * self.parts)()
        self._ck(self.L.espic_migrate_pack(self.h, sp, counts))
        return [int(v) for v in counts]

    def migrate_finish(self, sp):
        self._ck(self.L.espic_migrate_finish(self.h, sp))

    def migrate_segment(self, dest):
        """(device address, count) of the packed SoA [7][count] segment bound for part `dest`"""
        p, n = C.c_void_p(), C.c_longlong(0)
        self._ck(self.L.espic_migrate_segment(self.h, dest, C.byref(p), C.byref(n)))
        return (p.value or 0), n.value


def slab_bounds(nk, parts):
    """k_bounds of `parts` equal slabs of cells (the last one takes the remainder), the default cut of bench.py"""
    cells = nk - 1
    return [cells * r // parts for r in range(parts)] + [cells]


def balanced_bounds(weights, parts):
    """k_bounds that give every part about the same share of `weights` (one non-negative number per cell plane, e.g. the free
    volume or a particle histogram): the load balance of the decomposition (ch9/MPI cuts into equal node counts,
    World.h:96-99, whatever the particles do).  Every part keeps at least one cell plane."""
    w = np.asarray(weights, dtype=np.float64)
    cells = w.size
    assert 1 <= parts <= cells
    cum = np.concatenate([[0.0], np.cumsum(w)])
    kb = [0]
    for r in range(1, parts):
        k = int(np.searchsorted(cum, cum[-1] * r / parts, side="left"))
        if k > 0 and abs(cum[k - 1] - cum[-1] * r / parts) <= abs(cum[k] - cum[-1] * r / parts):
            k -= 1
        k = min(max(k, kb[-1] + 1), cells - (parts - r))
        kb.append(k)
    return kb + [cells]
