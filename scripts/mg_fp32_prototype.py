"""numpy prototype (design aid, not part of the product): the multigrid V-cycle of espic_mg.cuh evaluated in FP32 inside the FP64 CG.
The preconditioner only has to be a fixed SPD operator close to K^-1; if the iteration counts do not move, its vectors and
coefficients (x0, z, diag, minv and all coarse levels) can be stored in FP32, halving the traffic of the four fine-level passes."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mg_prototype as M, mg_semi_prototype as S
M.JW = 0.9

def cast(levels, dt):
    out = []
    for L in levels:
        out.append(dict(diag=L["diag"].astype(dt), c=[c.astype(dt) for c in L["c"]], mask=L["mask"], f=L.get("f")))
    return out

def vcycle(levels, l, b):
    L = levels[l]; x = np.zeros_like(b)
    if l == len(levels) - 1: return M.jacobi(L, x, b, 0.9, 6)
    x = M.jacobi(L, x, b, 0.9, 1); r = b - M.apply(L, x); f = levels[l + 1]["f"]
    x = x + S.prolong(vcycle(levels, l + 1, S.restrict(r, f)), b.shape, f) * L["mask"]
    return M.jacobi(L, x, b, 0.9, 1)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    p = M.build(n); phi = p["phi"].copy()
    rho = np.where(p["reg"], M.QE * p["n0"], 0.0) * (1 + 0.05 * np.random.default_rng(0).standard_normal(phi.shape))
    R = np.where(p["reg"], rho / M.EPS0, 0.0)
    for trial in range(2):
        L0 = M.fine_level(p, phi)
        lv = [L0]
        while lv[-1]["diag"].size > 4096 or len(lv) == 1:
            lv.append(S.coarsen_shape(lv[-1], (2, 2, 1) if len(lv) == 1 else (2, 2, 2)))
        lv32 = cast(lv, np.float32)
        y64, it64, l64 = M.pcg(L0, R, lambda r: vcycle(lv, 0, r), 1e-4, 300)
        # FP32 V-cycle: the residual is scaled to O(1) before the cast so that FP32 range is never an issue
        def m32(r):
            s = np.abs(r).max()
            if s == 0: return r
            return vcycle(lv32, 0, (r / s).astype(np.float32)).astype(np.float64) * s
        y32, it32, l32 = M.pcg(L0, R, m32, 1e-4, 300)
        print("n=%d  |R|=%.3e: FP64 V-cycle %d its (l2 %.2e), FP32 V-cycle %d its (l2 %.2e), |dy|/|y| = %.2e" % (
            n, np.sqrt((R * R).sum() / R.size), it64, l64, it32, l32, np.abs(y64 - y32).max() / np.abs(y64).max()))
        # a second, harder right-hand side: the potential has developed the sheath (phi = -100 V at the sphere screens out)
        phi = phi + y64
        R = np.where(p["reg"], R * 0.01 * np.random.default_rng(1).standard_normal(phi.shape), 0.0)
