"""numpy prototype (design aid, not part of the product): iteration counts of the multigrid-preconditioned CG of espic_mg.cuh for
variants of the cycle -- smoothing sweeps per level, cycle shape, damping, link scale -- on the warm Poisson problem of one bench
step (gpurun_out/warm_128.npz from `bench.py --dump-warm`).  The hierarchy mirrors k_mg_setup_from_level: 2x2x2 aggregates
(a direction whose spacing exceeds 1.42 x the smallest is not coarsened), mass term restricted exactly, links and the Laplacian
part of the diagonal scaled by `link_scale` per coarsened direction.

  python scripts/mg_variants_prototype.py gpurun_out/warm_128.npz"""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mg_prototype as M


def fine_split(p, phi):
    """fine level with the diagonal split into mass + per-direction Laplacian parts"""
    reg, face, g = p["reg"], p["face"], p["g"]
    mass = np.where(reg, p["n0"] * M.QE / (M.EPS0 * p["Te"]) * np.exp(phi / p["Te"]), 0.0)
    dl, links = [], []
    for a in range(3):
        lo = [slice(None)] * 3; hi = [slice(None)] * 3
        lo[a] = slice(0, -1); hi[a] = slice(1, None)
        lo, hi = tuple(lo), tuple(hi)
        d = np.where(reg, 2 * g[a], 0.0)
        c = np.zeros_like(d)
        c[lo] = np.where(reg[lo] & reg[hi], g[a], 0.0)
        d[hi] -= np.where(reg[hi] & face[lo], g[a], 0.0)
        d[lo] -= np.where(reg[lo] & face[hi], g[a], 0.0)
        dl.append(d); links.append(c)
    return dict(mass=mass, dl=dl, c=links, mask=reg, f=None)


def finish(L):
    L["diag"] = L["mass"] + L["dl"][0] + L["dl"][1] + L["dl"][2]
    d = L["diag"]
    L["inv"] = np.where(d > 0, 1.0 / np.where(d > 0, d, 1), 0.0)
    return L


def coarsen(L, f, scale):
    pad = [(0, (-s) % fa) for s, fa in zip(L["mass"].shape, f)]
    P = lambda a: np.pad(a, pad)
    csh = tuple((s + p_[1]) // fa for s, p_, fa in zip(L["mass"].shape, pad, f))
    agg = lambda a: a.reshape(csh[0], f[0], csh[1], f[1], csh[2], f[2]).sum(axis=(1, 3, 5))
    out = dict(mass=agg(P(L["mass"])), dl=[], c=[], mask=agg(P(L["mask"]).astype(float)) > 0, f=f)
    for a in range(3):
        c = P(L["c"][a]); d = P(L["dl"][a])
        if f[a] == 1:
            out["c"].append(agg(c)); out["dl"].append(agg(d)); continue
        idx = np.arange(c.shape[a]) % 2
        shape = [1, 1, 1]; shape[a] = -1
        inside = (idx == 0).reshape(shape)
        out["dl"].append(scale * (agg(d) - 2 * agg(c * inside)))
        out["c"].append(scale * agg(c * (~inside)))
    return finish(out)


def hierarchy(p, L0, scale, coarsest=4096):
    h = list(np.array([0.2, 0.2, 0.4]) / (p["n"] - 1))
    lv = [finish(L0)]
    while min(lv[-1]["diag"].shape) > 4 and lv[-1]["diag"].size > coarsest:
        hm = min(h)
        f = tuple(2 if x <= 1.42 * hm else 1 for x in h)
        h = [x * fa for x, fa in zip(h, f)]
        lv.append(coarsen(lv[-1], f, scale))
    return lv


def restrict(r, f):
    pad = [(0, (-s) % fa) for s, fa in zip(r.shape, f)]
    r = np.pad(r, pad)
    csh = tuple(s // fa for s, fa in zip(r.shape, f))
    return r.reshape(csh[0], f[0], csh[1], f[1], csh[2], f[2]).sum(axis=(1, 3, 5))


def prolong(e, sh, f):
    out = np.repeat(np.repeat(np.repeat(e, f[0], 0), f[1], 1), f[2], 2)
    return out[:sh[0], :sh[1], :sh[2]]


def jac(L, x, b, w, sweeps):
    for _ in range(sweeps):
        x = x + w * L["inv"] * (b - M.apply(L, x))
    return x


def cheb(L, x, b, degree, lmax=2.0, ratio=0.25):
    """Chebyshev smoothing on D^-1 A over [ratio*lmax, lmax] (lmax of D^-1 A <= 2 for this M-matrix)"""
    lmin = ratio * lmax
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    r = L["inv"] * (b - M.apply(L, x))
    d = r / theta
    x = x + d
    for _ in range(degree - 1):
        r = L["inv"] * (b - M.apply(L, x))
        rho_n = 1.0 / (2 * sigma - rho)
        d = rho_n * rho * d + 2 * rho_n / delta * r
        rho = rho_n
        x = x + d
    return x


def cycle(lv, l, b, o):
    L = lv[l]
    x = np.zeros_like(b)
    wl = o["w"] if l == 0 else o.get("wc", o["w"])       # optional: another damping factor on the coarse levels
    if l == len(lv) - 1:
        return jac(L, x, b, wl, o["coarsest"])
    nu = o["nu0"] if l == 0 else o["nu"]
    if o.get("cheb") and l == 0:
        x = cheb(L, x, b, o["cheb"])
    else:
        x = jac(L, x, b, wl, nu)
    f = lv[l + 1]["f"]
    for _ in range(o["gamma"] if l >= o["gamma_from"] else 1):
        r = b - M.apply(L, x)
        x = x + o["alpha"] * prolong(cycle(lv, l + 1, restrict(r, f), o), b.shape, f) * L["mask"]
    if o.get("cheb") and l == 0:
        return cheb(L, x, b, o["cheb"])
    return jac(L, x, b, wl, nu)


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


BASE = dict(w=0.9, nu0=1, nu=1, coarsest=7, gamma=1, gamma_from=1, alpha=1.0, scale=0.6)

def newton(p, phi, rho, tol=1e-4, nr_tol=1e-3, eta0=1e-2, eta_max=0.1, gamma=0.9, eta_pow=1.5, tight_after=None, verbose=True):
    """the Newton loop of mgn_body (inexact, Eisenstat-Walker forcing); tight_after = k: from Newton step k on the linear solve
    goes straight to tol/2 (two steps instead of three when the first one already lands in the quadratic regime)"""
    phi = phi.copy()
    o = dict(BASE)
    total, hist, Rprev, ynorm, last_full = 0, [], 0.0, 0.0, False
    for nit in range(25):
        R = residual(p, phi, rho)
        Rn = np.sqrt((R * R).sum() / R.size)
        if (nit == 0 and Rn < tol) or (nit > 0 and ynorm < nr_tol and (Rn < tol or last_full)):
            break
        lv = hierarchy(p, fine_split(p, phi), o["scale"])
        eta = eta0 if nit == 0 else min(eta_max, gamma * (Rn / Rprev) ** eta_pow)
        if tight_after is not None and nit >= tight_after:
            eta = 0.0
        stop = max(0.5 * tol, eta * Rn)
        last_full = stop <= 0.5 * tol
        Rprev = Rn
        y, it, l2 = M.pcg(lv[0], R, lambda r: cycle(lv, 0, r, o), stop, 500) if Rn >= stop else (np.zeros_like(R), 0, Rn)
        total += it
        phi = phi + y
        ynorm = np.sqrt((y * y).sum() / y.size)
        hist.append((Rn, it, l2, ynorm))
        if verbose:
            print("   newton %d: |R| %.3e  eta %.1e  -> %d its (l2 %.2e)  |y| %.2e" % (nit, Rn, eta, it, l2, ynorm), flush=True)
    Rf = residual(p, phi, rho)
    return phi, total, len(hist), np.sqrt((Rf * Rf).sum() / Rf.size)


def newton_study(path):
    d = np.load(path)
    n = int(d["mesh"])
    p = M.build(n)
    to3 = lambda a: a.reshape(n, n, n).transpose(2, 1, 0).copy()
    phi0, rho1 = to3(d["phi"]), to3(d["rho_next"])
    for name, kw in (("as built (eta0 1e-2, EW 0.9 x ratio^1.5)", {}),
                     ("tight from the second step", dict(tight_after=1)),
                     ("eta0 3e-3, tight from the second step", dict(eta0=3e-3, tight_after=1)),
                     ("eta0 1e-3, tight from the second step", dict(eta0=1e-3, tight_after=1)),
                     ("eta0 3e-2, tight from the second step", dict(eta0=3e-2, tight_after=1)),
                     ("one tight step", dict(tight_after=0))):
        t = time.time()
        phi, total, steps, Rf = newton(p, phi0, rho1, **kw)
        print("%-45s: %d Newton steps, %d CG iterations, final |R| %.2e  (%.0f s)" % (name, steps, total, Rf, time.time() - t), flush=True)


if __name__ == "__main__":
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/warm_128.npz"
    if len(sys.argv) > 2 and sys.argv[2] == "newton":
        newton_study(path)
        sys.exit(0)
    d = np.load(path)
    n = int(d["mesh"])
    p = M.build(n)
    to3 = lambda a: a.reshape(n, n, n).transpose(2, 1, 0).copy()          # flat U order (k slowest) -> [i][j][k]
    phi0, rho1 = to3(d["phi"]), to3(d["rho_next"])
    R = residual(p, phi0, rho1)
    Rn = np.sqrt((R * R).sum() / R.size)
    print("mesh %d^3, warm residual %.3e; linear solve to 1e-6 of it" % (n, Rn), flush=True)
    variants = [("kernel as built: V(1,1), w 0.9, scale 0.6, 7 coarsest sweeps", {})]
    for a in sys.argv[2:]:                         # name=key:val,key:val
        name, kv = a.split("=", 1)
        variants.append((name, {k: float(v) if "." in v else int(v) for k, v in (x.split(":") for x in kv.split(","))}))
    if len(sys.argv) <= 2:
        variants += [
            ("2 sweeps on the coarse levels", dict(nu=2)),
            ("3 sweeps on the coarse levels", dict(nu=3)),
            ("W cycle from level 1", dict(gamma=2, gamma_from=1)),
            ("W cycle from level 2", dict(gamma=2, gamma_from=2)),
            ("coarse levels 2 sweeps + W from level 2", dict(nu=2, gamma=2, gamma_from=2)),
            ("w 0.8", dict(w=0.8)), ("w 1.0", dict(w=1.0)),
            ("scale 0.55", dict(scale=0.55)), ("scale 0.65", dict(scale=0.65)),
            ("14 coarsest sweeps", dict(coarsest=14)),
            ("fine level 2 sweeps", dict(nu0=2)),
            ("fine level Chebyshev degree 2", dict(cheb=2)),
            ("coarse correction x 1.2", dict(alpha=1.2)),
        ]
    for name, kw in variants:
        o = dict(BASE); o.update(kw)
        t = time.time()
        lv = hierarchy(p, fine_split(p, phi0), o["scale"])
        y, it, l2 = M.pcg(lv[0], R, lambda r: cycle(lv, 0, r, o), 1e-6 * Rn, 200)
        print("%-70s %3d CG iterations  (%d levels, %.0f s)" % (name, it, len(lv), time.time() - t), flush=True)
