import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
es = bench.load_espic()
n_mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20000000
dev = torch.device("cuda", 0)
e = es.Engine(n_mesh, n_mesh, n_mesh, bench.X0, bench.XM)
e.add_sphere(*bench.SPHERE); e.add_inlet(); e.set_reference_values(0.0, 1.5, 1e12)
mpw = 1e12 * 0.016 / n
sp = e.add_species(16 * bench.AMU, bench.QE, mpw, capacity=n + 1024)
t = bench.make_particles_device(torch, n, 1, mpw, dev)
torch.cuda.synchronize()
e.upload_device(sp, [t[c].data_ptr() for c in range(7)], n, mpw); e.sync()
def stat(name):
    phi = e.field(es.PHI); rho = e.field(es.RHO)
    print(name, "phi[min,max,nan]", phi.min(), phi.max(), np.isnan(phi).sum(), "rho/qe/n0 mean", rho.mean() / bench.QE / 1e12, flush=True)
e.sort_by_cell(sp); e.deposit(sp); e.compute_charge_density(); stat("deposit")
e.solve(es.SOLVE_QN, 1, 1.0); stat("qn")
for its in (100, 1000, 5000):
    t0 = time.time(); info = e.solve(es.SOLVE_GS, its, 1e-2); e.sync(); print("GS", its, info, time.time() - t0); stat("gs")
t0 = time.time(); info = e.solve(es.SOLVE_PCG, 2000, 1e-4); e.sync(); print("PCG", info, time.time() - t0); stat("pcg")
e.compute_ef()
for i in range(3):
    t0 = time.time()
    e.push(sp, 1e-7, es.WALL_ABSORB, es.PUSH_FUSE_DEPOSIT); e.deposit(sp); e.compute_charge_density()
    info = e.solve(es.SOLVE_PCG, 2000, 1e-4); e.compute_ef(); e.sync()
    print("step", i, e.count(sp), info, time.time() - t0, flush=True)
