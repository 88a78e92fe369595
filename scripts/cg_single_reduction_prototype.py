"""numpy prototype (design aid, not part of the product): Chronopoulos-Gear CG -- one global reduction per iteration instead of
three -- with the aggregation-multigrid preconditioner of espic_mg.cuh.  On the 64^3 sphere case it takes the same 13 iterations to the
same residual as the standard recurrence (the stopping test sees |r| one V-cycle later).  Candidate for the slab solver, where a
reduction is an inter-GPU barrier."""
import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
import mg_prototype as M, mg_semi_prototype as S
M.JW = 0.9
def vcycle(levels, l, b):
    L = levels[l]; x = np.zeros_like(b)
    if l == len(levels) - 1: return M.jacobi(L, x, b, 0.9, 6)
    x = M.jacobi(L, x, b, 0.9, 1); r = b - M.apply(L, x); f = levels[l + 1]["f"]
    x = x + S.prolong(vcycle(levels, l + 1, S.restrict(r, f)), b.shape, f) * L["mask"]
    return M.jacobi(L, x, b, 0.9, 1)
def pcg_cg(L, b, Minv, tol, maxit):
    """Chronopoulos-Gear: one reduction per iteration (gamma = r.z, delta = w.z, |r|^2 of the previous update)"""
    x = np.zeros_like(b); r = b.copy(); nn = b.size
    d = np.zeros_like(b); q = np.zeros_like(b); gamma_old = alpha_old = None
    for it in range(1, maxit + 1):
        z = Minv(r); w = M.apply(L, z)
        gamma = (r * z).sum(); delta = (w * z).sum(); rr = (r * r).sum()      # ONE reduction
        if np.sqrt(rr / nn) < tol: return x, it - 1, np.sqrt(rr / nn)
        if it == 1: beta = 0.0; alpha = gamma / delta
        else: beta = gamma / gamma_old; alpha = gamma / (delta - beta * gamma / alpha_old)
        d = z + beta * d; q = w + beta * q
        x += alpha * d; r -= alpha * q
        gamma_old, alpha_old = gamma, alpha
    return x, maxit, np.sqrt((r * r).sum() / nn)
n = 64
p = M.build(n); phi = p["phi"].copy()
rho = np.where(p["reg"], M.QE * p["n0"], 0.0) * (1 + 0.05 * np.random.default_rng(0).standard_normal(phi.shape))
L0 = M.fine_level(p, phi)
lv = [L0]
while lv[-1]["diag"].size > 4096 or len(lv) == 1:
    lv.append(S.coarsen_shape(lv[-1], (2,2,1) if len(lv)==1 else (2,2,2)))
R = np.where(p["reg"], rho / M.EPS0, 0.0)
y1, it1, l1 = M.pcg(L0, R, lambda r: vcycle(lv, 0, r), 1e-4, 300)
y2, it2, l2 = pcg_cg(L0, R, lambda r: vcycle(lv, 0, r), 1e-4, 300)
print("standard PCG:", it1, l1, " Chronopoulos-Gear:", it2, l2, " |dy|/|y|", np.abs(y1 - y2).max() / np.abs(y1).max())
