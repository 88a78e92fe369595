"""Two interchangeable executors of the harness command language (oracle/ref_harness.cpp):
OracleEngine (CPU restatement, oracle/) and GpuEngine (the product, through the C ABI).
Both start from a statefile.State and return one, so every parity test is
    run_ref(...) / OracleEngine(...).run(cmds)  ==  GpuEngine(...).run(cmds)
"""
import numpy as np

import statefile as sf
from cases import orc


class OracleEngine:
    def __init__(self, st, box=False):
        self.box = box
        self.dt = st.dt
        w = orc.World(st.ni, st.nj, st.nk, tuple(st.x0), tuple(st.xm))
        if not box:
            if st.flags & 1:
                w.add_sphere(tuple(st.sphere_c), st.sphere_r, st.sphere_phi)
            if st.flags & 2:
                w.add_inlet()
        if not (st.flags & 4):
            w.phi[:] = st.phi
        w.rho[:] = st.rho
        w.ef[:] = st.ef
        w.set_reference_values(st.phi0, st.Te0, st.n0)
        self.w = w
        self.species = []
        for r in st.species:
            sp = orc.Species(w, r["mass"], r["charge"], r.get("mpw0", 1.0), cap=max(16, 2 * r["part"].shape[1]))
            sp.set_particles(r["part"])
            sp.den[:] = r["den"]
            sp.den_ave[:] = r["den_ave"]
            self.species.append(sp)
        self.converged = -1.0
        self.flags = st.flags

    def run(self, cmds):
        w = self.w
        for cmd in cmds:
            c = cmd.split(":")
            op = c[0]
            if op == "advance":
                for sp in self.species:
                    sp.advance_box(self.dt) if self.box else sp.advance(self.dt)
            elif op == "deposit":
                for sp in self.species:
                    sp.compute_number_density()
            elif op == "rho":
                w.compute_charge_density(self.species)
            elif op == "ef":
                w.compute_ef()
            elif op == "solve":
                self.converged = float(w.solve_gs_box(int(c[1]), float(c[2]))["converged"])
            elif op == "solve_gs":
                self.converged = float(w.solve_gs(int(c[1]), float(c[2]))["converged"])
            elif op == "solve_pcg":
                self.converged = float(w.solve_nrpcg(int(c[1]), float(c[2]))["converged"])
            elif op == "solve_qn":
                w.solve_qn()
                self.converged = 1.0
            elif op == "sample":
                g = orc.mt19937(int(c[4]))
                for _ in range(int(c[5]) if len(c) > 5 else 1):
                    self.species[int(c[1])].sample_cold_beam_mt(float(c[2]), float(c[3]), self.dt, g)
            elif op == "loadqs":
                s = self.species[int(c[1])]
                s.load_box_qs(w.x0, w.xc if int(c[6]) else w.xm, float(c[2]), (int(c[3]), int(c[4]), int(c[5])), self.dt)
            elif op == "average":
                self.species[int(c[1])].update_averages()
            else:
                raise ValueError(cmd)
        return self.state()

    def state(self):
        st = sf.state_from_oracle(self.w, self.species, self.dt, flags=self.flags)
        st.diag[0] = self.converged
        st.diag[1] = self.w.pe()
        for q, sp in enumerate(self.species[:2]):
            st.diag[2 + 5 * q: 7 + 5 * q] = np.concatenate([[sp.real_count()], sp.momentum(), [sp.ke()]])
        return st


# ---------------------------------------------------------------------------------------------------
# the product: libespic_cuda.so through the C ABI
# ---------------------------------------------------------------------------------------------------

def _espic():
    import importlib.util
    import os
    path = os.path.join(sf.ROOT, "plasma-simulations-by-example_b200", "espic.py")
    spec = importlib.util.spec_from_file_location("espic", path)
    import sys
    if "espic" in sys.modules:
        return sys.modules["espic"]
    mod = importlib.util.module_from_spec(spec)
    sys.modules["espic"] = mod
    spec.loader.exec_module(mod)
    return mod


class GpuEngine:
    """Same command language, executed by the CUDA library.  Options:
    pcg_ref  -- solve_pcg runs the reference-exact (fragile) PCG instead of the SPD one
    fuse     -- advance uses ESPIC_PUSH_FUSE_DEPOSIT (deposit only finalises)
    fixed    -- fixed-point deposition
    sort     -- sort particles by cell before every advance (changes particle order, nothing else)"""

    def __init__(self, st, box=False, fuse=False, fixed=False, sort=False, device=0, pcg_ref=False):
        es = _espic()
        self.es = es
        self.box, self.fuse, self.fixed, self.sort = box, fuse, fixed, sort
        self.pcg_ref = pcg_ref      # solve_pcg -> ESPIC_SOLVE_PCG_REF (the reference's algorithm operation for operation)
        self.dt = st.dt
        self.flags = st.flags
        self.st0 = st
        e = es.Engine(st.ni, st.nj, st.nk, st.x0, st.xm, device=device)
        if not box:
            if st.flags & 1:
                e.add_sphere(st.sphere_c, st.sphere_r, st.sphere_phi)
            if st.flags & 2:
                e.add_inlet()
        if not (st.flags & 4):
            e.set_field(es.PHI, st.phi)
        e.set_field(es.RHO, st.rho)
        e.set_field(es.EF, st.ef)
        e.set_reference_values(st.phi0, st.Te0, st.n0)
        self.e = e
        self.species = []
        for r in st.species:
            sp = e.add_species(r["mass"], r["charge"], r.get("mpw0", 1.0), capacity=max(16, r["part"].shape[1]))
            e.upload(sp, r["part"])
            e.set_field(es.DEN, r["den"], sp)
            e.set_field(es.DEN_AVE, r["den_ave"], sp)
            self.species.append(sp)
        self.converged = -1.0
        self.info = None

    def run(self, cmds):
        e, es = self.e, self.es
        dmode = es.DEPOSIT_FIXED if self.fixed else es.DEPOSIT_FP64
        for cmd in cmds:
            c = cmd.split(":")
            op = c[0]
            if op == "advance":
                flags = (es.PUSH_FUSE_DEPOSIT if self.fuse else 0) | (es.PUSH_FIXED_POINT if (self.fuse and self.fixed) else 0)
                for sp in self.species:
                    if self.sort:
                        e.sort_by_cell(sp)
                    e.push(sp, self.dt, es.WALL_REFLECT if self.box else es.WALL_ABSORB, flags)
            elif op == "deposit":
                for sp in self.species:
                    e.deposit(sp, dmode)
            elif op == "rho":
                e.compute_charge_density()
            elif op == "ef":
                e.compute_ef()
            elif op in ("solve", "solve_gs", "solve_pcg", "solve_qn", "solve_mg", "solve_mgslab"):
                kind = {"solve": es.SOLVE_GS_BOX, "solve_gs": es.SOLVE_GS, "solve_qn": es.SOLVE_QN, "solve_mg": es.SOLVE_PCG_MG,
                        "solve_mgslab": es.SOLVE_PCG_MG_SLAB,
                        "solve_pcg": es.SOLVE_PCG_REF if self.pcg_ref else es.SOLVE_PCG}[op]
                self.info = e.solve(kind, int(c[1]) if len(c) > 1 else 1, float(c[2]) if len(c) > 2 else 1.0)
                self.converged = float(self.info["converged"])
            elif op == "loadqs":
                self._loadqs(self.species[int(c[1])], float(c[2]), (int(c[3]), int(c[4]), int(c[5])), int(c[6]))
            elif op == "average":
                e.update_average(self.species[int(c[1])])
            else:
                raise ValueError(cmd)
        return self.state()

    def _loadqs(self, sp, num_den, grid, half):
        """Host side of Species::loadParticlesBoxQS (ch2/Species.cpp:101-141): candidate generation in numpy with the
        reference's arithmetic; admission + half-step rewind happen on the GPU (espic_species_add)."""
        st = self.st0
        x1 = np.asarray(st.x0, dtype=np.float64)
        x2 = (np.asarray(st.x0) + np.asarray(st.xm)) * 0.5 if half else np.asarray(st.xm, dtype=np.float64)
        box_vol = (x2[0] - x1[0]) * (x2[1] - x1[1]) * (x2[2] - x1[2])
        tot = (grid[0] - 1) * (grid[1] - 1) * (grid[2] - 1)
        mpw = num_den * box_vol / tot
        d = [(x2[a] - x1[a]) / (grid[a] - 1) for a in range(3)]
        i, j, k = np.meshgrid(np.arange(grid[0]), np.arange(grid[1]), np.arange(grid[2]), indexing="ij")
        idx = [i.ravel(), j.ravel(), k.ravel()]
        soa = np.zeros((7, idx[0].size))
        w = np.ones(idx[0].size)
        for a in range(3):
            p = x1[a] + idx[a] * d[a]
            p = np.where(p == x2[a], p - 1e-4 * d[a], p)
            soa[a] = p
            w = np.where((idx[a] == 0) | (idx[a] == grid[a] - 1), w * 0.5, w)
        soa[6] = mpw * w
        self.e.add_particles(sp, soa, self.dt)

    def state(self):
        e, es = self.e, self.es
        st = sf.State()
        s0 = self.st0
        st.ni, st.nj, st.nk, st.flags = s0.ni, s0.nj, s0.nk, s0.flags
        st.x0, st.xm, st.dt = s0.x0, s0.xm, s0.dt
        st.sphere_c, st.sphere_r, st.sphere_phi = s0.sphere_c, s0.sphere_r, s0.sphere_phi
        st.phi0, st.Te0, st.n0 = s0.phi0, s0.Te0, s0.n0
        st.phi, st.rho, st.ef = e.field(es.PHI), e.field(es.RHO), e.field(es.EF)
        st.node_vol, st.object_id = e.field(es.NODE_VOL), e.field(es.OBJECT_ID)
        st.diag[0] = self.converged
        st.diag[1] = e.pe()
        for q, sp in enumerate(self.species):
            r0 = s0.species[q]
            st.species.append(dict(mass=r0["mass"], charge=r0["charge"], mpw0=r0.get("mpw0", 1.0),
                                   den=e.field(es.DEN, sp), den_ave=e.field(es.DEN_AVE, sp), part=e.download(sp)))
            if q < 2:
                st.diag[2 + 5 * q: 7 + 5 * q] = e.diag(sp)
        return st
