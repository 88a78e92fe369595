"""numpy prototype: vertex-centred multigrid with trilinear interpolation and rediscretised coarse operators as CG
preconditioner for the SPD Newton system (compare with the aggregation hierarchy of scripts/mg_prototype.py)."""
import sys, time
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__file__))
import mg_prototype as M

EPS0, QE = M.EPS0, M.QE

def level_from(reg, face, g, Pb):
    """7-point operator on a grid: diag, links (REG-REG), Neumann faces folded"""
    diag = np.where(reg, 2 * g.sum() + Pb, 0.0)
    links = []
    for a in range(3):
        lo = [slice(None)] * 3; hi = [slice(None)] * 3
        lo[a] = slice(0, -1); hi[a] = slice(1, None)
        lo, hi = tuple(lo), tuple(hi)
        c = np.zeros_like(diag)
        c[lo] = np.where(reg[lo] & reg[hi], g[a], 0.0)
        links.append(c)
        diag[hi] -= np.where(reg[hi] & face[lo], g[a], 0.0)
        diag[lo] -= np.where(reg[lo] & face[hi], g[a], 0.0)
    return dict(diag=diag, c=links, mask=reg)

def coarsen_geo(L, reg, face, dirich, g, Pb):
    regc = reg[::2, ::2, ::2]; dirc = dirich[::2, ::2, ::2]
    sh = regc.shape
    I, J, K = np.indices(sh)
    # coarse faces: index 0 planes and the last coarse plane if it coincides with / lies next to the fine wall
    facec = face[::2, ::2, ::2].copy()
    Pbc = Pb[::2, ::2, ::2]
    gc = g / 4.0
    # a coarse node that is REG on the fine grid but whose +neighbour does not exist (fine n even): wall beyond it -> Neumann fold
    Lc = level_from(regc, facec, gc, Pbc)
    for a in range(3):
        n_f = reg.shape[a]
        if n_f % 2 == 0:     # last coarse node = fine n-2, wall at fine n-1 (distance h): treat as Neumann on the coarse grid
            sl = [slice(None)] * 3; sl[a] = -1; sl = tuple(sl)
            Lc["diag"][sl] -= np.where(regc[sl], gc[a], 0.0)
    return Lc, regc, facec, dirc, gc, Pbc

def prolong_geo(e, fsh):
    """trilinear interpolation from coarse (even fine indices) to fine"""
    out = e
    for a in range(3):
        n_f = fsh[a]
        shp = list(out.shape); shp[a] = n_f
        f = np.zeros(shp)
        ev = [slice(None)] * 3; ev[a] = slice(0, n_f, 2)
        f[tuple(ev)] = out
        nodd = n_f // 2
        od = [slice(None)] * 3; od[a] = slice(1, n_f, 2)
        lo = [slice(None)] * 3; lo[a] = slice(0, nodd)
        hi = [slice(None)] * 3; hi[a] = slice(1, nodd + 1)
        left = out[tuple(lo)]
        right = np.zeros_like(left)
        rr = out[tuple(hi)]
        sl = [slice(None)] * 3; sl[a] = slice(0, rr.shape[a])
        right[tuple(sl)] = rr
        f[tuple(od)] = 0.5 * (left + right)
        out = f
    return out

def restrict_geo(r):
    """(1/8) P^T"""
    out = r
    for a in range(3):
        n_f = out.shape[a]
        nc = (n_f + 1) // 2
        ev = [slice(None)] * 3; ev[a] = slice(0, n_f, 2)
        c = out[tuple(ev)].copy()
        od = [slice(None)] * 3; od[a] = slice(1, n_f, 2)
        o = out[tuple(od)]
        nodd = o.shape[a]
        lo = [slice(None)] * 3; lo[a] = slice(0, nodd)
        c[tuple(lo)] += 0.5 * o
        hi = [slice(None)] * 3; hi[a] = slice(1, min(nodd + 1, nc))
        src = [slice(None)] * 3; src[a] = slice(0, min(nodd + 1, nc) - 1)
        c[tuple(hi)] += 0.5 * o[tuple(src)]
        out = 0.5 * c
    return out

def vcycle(levels, l, b, nu, w, coarse_sweeps):
    L = levels[l]
    x = np.zeros_like(b)
    if l == len(levels) - 1:
        return M.jacobi(L, x, b, w, coarse_sweeps)
    x = M.jacobi(L, x, b, w, nu)
    r = (b - M.apply(L, x)) * L["mask"]
    ec = vcycle(levels, l + 1, restrict_geo(r) * levels[l + 1]["mask"], nu, w, coarse_sweeps)
    x = x + prolong_geo(ec, b.shape) * L["mask"]
    x = M.jacobi(L, x, b, w, nu)
    return x

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    nlev = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    p = M.build(n)
    phi = p["phi"].copy()
    rho = np.where(p["reg"], QE * p["n0"], 0.0) * (1 + 0.05 * np.random.default_rng(0).standard_normal(phi.shape))
    g = p["g"]
    for newton in range(3):
        L0 = M.fine_level(p, phi)
        Pb = p["n0"] * QE / (EPS0 * p["Te"]) * np.exp(phi / p["Te"])
        ph = phi.copy()
        ph[0] = np.where(p["face"][0], ph[1], ph[0]); ph[-1] = np.where(p["face"][-1], ph[-2], ph[-1])
        ph[:, 0] = np.where(p["face"][:, 0], ph[:, 1], ph[:, 0]); ph[:, -1] = np.where(p["face"][:, -1], ph[:, -2], ph[:, -1])
        ph[:, :, -1] = np.where(p["face"][:, :, -1], ph[:, :, -2], ph[:, :, -1])
        lap = np.zeros_like(ph)
        lap[1:-1, 1:-1, 1:-1] = (g[0] * (ph[2:, 1:-1, 1:-1] + ph[:-2, 1:-1, 1:-1]) + g[1] * (ph[1:-1, 2:, 1:-1] + ph[1:-1, :-2, 1:-1])
                                 + g[2] * (ph[1:-1, 1:-1, 2:] + ph[1:-1, 1:-1, :-2]) - 2 * g.sum() * ph[1:-1, 1:-1, 1:-1])
        ne = p["n0"] * np.exp(phi / p["Te"])
        R = np.where(p["reg"], lap + (rho - QE * ne) / EPS0, 0.0)
        # aggregation hierarchy (current product)
        lv = [L0]
        for _ in range(nlev - 1):
            lv.append(M.coarsen(lv[-1]))
        t = time.time(); y, it, l2 = M.pcg(L0, R, lambda r: M.vcycle(lv, 0, r, "jac", 1, 1.0, 3), 1e-4, 300)
        print("newton %d |R|=%.2e  aggregation jac nu=1: %d its %.1fs" % (newton, np.sqrt((R * R).sum() / R.size), it, time.time() - t))
        # geometric hierarchy
        levels = [L0]; reg, face, dirich, gg, PP = p["reg"], p["face"], p["dirich"], g, np.where(p["reg"], Pb, 0.0)
        for _ in range(nlev - 1):
            Lc, reg, face, dirich, gg, PP = coarsen_geo(levels[-1], reg, face, dirich, gg, PP)
            levels.append(Lc)
        for nu, w in ((1, 0.8), (1, 0.9), (2, 0.8)):
            t = time.time(); y2, it2, l22 = M.pcg(L0, R, lambda r: vcycle(levels, 0, r, nu, w, 8), 1e-4, 300)
            print("   geometric trilinear nu=%d w=%.1f: %d its %.1fs  |y-y2|/|y|=%.1e" % (nu, w, it2, time.time() - t, np.abs(y - y2).max() / np.abs(y).max()))
        phi = phi + y
