"""numpy prototype (design aid, not part of the product) of the round-2 solver in espic_mg.cuh: inexact Newton with
Eisenstat-Walker forcing around the multigrid-preconditioned CG, FP32 storage of everything the preconditioner alone reads and
of the Jacobian diagonal.  Counts CG iterations per solve for forcing parameters and storage choices.

  python scripts/mg_newton_prototype.py [n=64] [warm.npz]
with warm.npz = bench.py --dump-warm (phi of step n, rho of step n+1 at the bench size); without it a synthetic warm start:
converge on a noisy rho, then solve for a second, independently perturbed rho."""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mg_prototype as M
import mg_semi_prototype as S

W = 0.9


def residual(p, phi, rho):
    g = p["g"]
    ph = phi.copy()
    f = p["face"]
    ph[0] = np.where(f[0], ph[1], ph[0]); ph[-1] = np.where(f[-1], ph[-2], ph[-1])
    ph[:, 0] = np.where(f[:, 0], ph[:, 1], ph[:, 0]); ph[:, -1] = np.where(f[:, -1], ph[:, -2], ph[:, -1])
    ph[:, :, -1] = np.where(f[:, :, -1], ph[:, :, -2], ph[:, :, -1])
    lap = np.zeros_like(ph)
    lap[1:-1, 1:-1, 1:-1] = (g[0] * (ph[2:, 1:-1, 1:-1] + ph[:-2, 1:-1, 1:-1]) + g[1] * (ph[1:-1, 2:, 1:-1] + ph[1:-1, :-2, 1:-1])
                             + g[2] * (ph[1:-1, 1:-1, 2:] + ph[1:-1, 1:-1, :-2]) - 2 * g.sum() * ph[1:-1, 1:-1, 1:-1])
    ne = p["n0"] * np.exp(phi / p["Te"])
    return np.where(p["reg"], lap + (rho - M.QE * ne) / M.EPS0, 0.0)


def hierarchy(p, L0):
    h = list(np.array([0.2, 0.2, 0.4]) / (p["n"] - 1))
    lv = [L0]
    while min(lv[-1]["diag"].shape) > 4 and lv[-1]["diag"].size > 4096:
        hm = min(h)
        f = tuple(2 if x <= 1.42 * hm else 1 for x in h)
        h = [x * fa for x, fa in zip(h, f)]
        lv.append(S.coarsen_shape(lv[-1], f))
    return lv


def cast32(lv):
    return [dict(diag=L["diag"].astype(np.float32), c=[c.astype(np.float32) for c in L["c"]], mask=L["mask"], f=L.get("f")) for L in lv]


def jac(L, x, b, sweeps):
    d = L["diag"]
    inv = np.where(d > 0, 1.0 / np.where(d > 0, d, 1), 0.0).astype(d.dtype)
    for _ in range(sweeps):
        x = (x + W * inv * (b - M.apply(L, x))).astype(b.dtype)
    return x


def vcycle(lv, l, b):
    L = lv[l]
    x = np.zeros_like(b)
    if l == len(lv) - 1:
        return jac(L, x, b, 7)
    x = jac(L, x, b, 1)
    r = b - M.apply(L, x)
    f = lv[l + 1]["f"]
    x = x + S.prolong(vcycle(lv, l + 1, S.restrict(r, f).astype(b.dtype)), b.shape, f).astype(b.dtype) * L["mask"]
    return jac(L, x, b, 1)


def newton(p, phi, rho, tol=1e-4, nr_tol=1e-3, eta0=1e-2, eta_max=0.1, gamma=0.9, fp32=True, exact=False, verbose=True):
    phi = phi.copy()
    nn = phi.size
    ynorm, Rprev, total = 0.0, 0.0, 0
    hist = []
    for nit in range(25):
        R = residual(p, phi, rho)
        Rn = np.sqrt((R * R).sum() / nn)
        if (nit == 0 and Rn < tol) or (nit > 0 and ynorm < nr_tol and Rn < tol):
            break
        L0 = M.fine_level(p, phi)
        if fp32:
            L0["diag"] = L0["diag"].astype(np.float32).astype(np.float64)       # the operator uses the FP32-rounded diagonal
        lv = hierarchy(p, L0)
        lvp = cast32(lv) if fp32 else lv
        eta = 0.0 if exact else (eta0 if nit == 0 else min(eta_max, gamma * (Rn / Rprev) ** 2))
        stop = max(0.5 * tol, eta * Rn)
        Rprev = Rn
        if fp32:
            Mop = lambda r: vcycle(lvp, 0, r.astype(np.float32)).astype(np.float64)
        else:
            Mop = lambda r: vcycle(lvp, 0, r)
        y, it, l2 = M.pcg(L0, R, Mop, stop, 500) if Rn >= stop else (np.zeros_like(R), 0, Rn)
        total += it
        phi = phi + y
        ynorm = np.sqrt((y * y).sum() / nn)
        hist.append((Rn, it, l2, ynorm))
        if verbose:
            print("   newton %d: |R| %.3e  eta %.1e  -> %d its (l2 %.2e)  |y| %.2e" % (nit, Rn, eta, it, l2, ynorm), flush=True)
    return phi, total, hist


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    p = M.build(n)
    if len(sys.argv) > 2:
        d = np.load(sys.argv[2])
        to3 = lambda a: a.reshape(n, n, n).transpose(2, 1, 0).copy()          # flat U order (k slowest) -> [i][j][k]
        phi0, rho1 = to3(d["phi"]), to3(d["rho_next"])
    else:
        rng = np.random.default_rng(0)
        base = np.where(p["reg"], M.QE * p["n0"], 0.0)
        rho0 = base * (1 + 0.10 * rng.standard_normal(base.shape))
        print("cold solve (sets up the warm state)")
        phi0, _, _ = newton(p, p["phi"].copy(), rho0, exact=False, eta0=0.1)
        rho1 = 0.97 * rho0 + base * 0.03 * (1 + 0.10 * np.sqrt(1 / 0.03) * rng.standard_normal(base.shape))     # ~3 % of the noise renewed
    for name, kw in (("exact Newton, FP64 precond", dict(exact=True, fp32=False)),
                     ("exact Newton, FP32 storage", dict(exact=True, fp32=True)),
                     ("EW eta0=1e-1, FP32", dict(eta0=1e-1)),
                     ("EW eta0=1e-2, FP32", dict(eta0=1e-2)),
                     ("EW eta0=1e-3, FP32", dict(eta0=1e-3)),
                     ("EW eta0=1e-2, FP64 precond", dict(eta0=1e-2, fp32=False))):
        t = time.time()
        phi, total, hist = newton(p, phi0, rho1, verbose=True, **kw)
        print("%-28s: %d Newton solves, %d CG iterations, final |R| %.2e  (%.0f s)" % (
            name, len(hist), total, np.sqrt((residual(p, phi, rho1) ** 2).sum() / phi.size), time.time() - t), flush=True)
