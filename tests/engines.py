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
