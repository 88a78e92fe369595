"""numpy prototype of the aggregation-multigrid preconditioner for the SPD Newton system (design aid for the CUDA
implementation in espic_mg.cu; not part of the product).  K = diag - sum(link * neighbour) on REG nodes."""
import sys, time
import numpy as np

EPS0, QE = 8.85418782e-12, 1.602176565e-19

def build(n, n0=1e12, Te=1.5):
    ni = nj = nk = n
    dh = np.array([0.2, 0.2, 0.4]) / (n - 1)
    g = 1.0 / dh ** 2
    I, J, K = np.meshgrid(np.arange(ni), np.arange(nj), np.arange(nk), indexing="ij")
    x = -0.1 + I * dh[0]; y = -0.1 + J * dh[1]; z = K * dh[2]
    sphere = (x ** 2 + y ** 2 + (z - 0.15) ** 2) <= 0.05 ** 2
    dirich = sphere | (K == 0)
    face = ((I == 0) | (I == ni - 1) | (J == 0) | (J == nj - 1) | (K == nk - 1)) & ~dirich
    reg = ~dirich & ~face
    phi = np.zeros((ni, nj, nk)); phi[sphere] = -100.0
    return dict(n=n, g=g, reg=reg, face=face, dirich=dirich, phi=phi, n0=n0, Te=Te)

def fine_level(p, phi):
    reg, face, g = p["reg"], p["face"], p["g"]
    P = p["n0"] * QE / (EPS0 * p["Te"]) * np.exp(phi / p["Te"])
    diag = np.where(reg, 2 * g.sum() + P, 0.0)
    links = []
    for a in range(3):
        # link between u and u+e_a
        sl_lo = [slice(None)] * 3; sl_hi = [slice(None)] * 3
        sl_lo[a] = slice(0, -1); sl_hi[a] = slice(1, None)
        c = np.zeros_like(diag)
        both = reg[tuple(sl_lo)] & reg[tuple(sl_hi)]
        c[tuple(sl_lo)] = np.where(both, g[a], 0.0)
        links.append(c)
        # Neumann neighbour folds into the diagonal
        nlo = reg[tuple(sl_hi)] & face[tuple(sl_lo)]      # node hi has a face neighbour below
        nhi = reg[tuple(sl_lo)] & face[tuple(sl_hi)]
        diag[tuple(sl_hi)] -= np.where(nlo, g[a], 0.0)
        diag[tuple(sl_lo)] -= np.where(nhi, g[a], 0.0)
    return dict(diag=diag, c=links, mask=reg)

def apply(L, v):
    out = L["diag"] * v
    for a in range(3):
        c = L["c"][a]
        lo = [slice(None)] * 3; hi = [slice(None)] * 3
        lo[a] = slice(0, -1); hi[a] = slice(1, None)
        lo, hi = tuple(lo), tuple(hi)
        out[lo] -= c[lo] * v[hi]
        out[hi] -= c[lo] * v[lo]
    return out

def coarsen(L):
    d, cs = L["diag"], L["c"]
    sh = d.shape
    pad = [(0, s % 2) for s in sh]
    def P(a): return np.pad(a, pad)
    d = P(d); cs = [P(c) for c in cs]; mask = P(L["mask"])
    csh = tuple(s // 2 for s in d.shape)
    def agg(a): return a.reshape(csh[0], 2, csh[1], 2, csh[2], 2).sum(axis=(1, 3, 5))
    dc = agg(d)
    cc = []
    for a in range(3):
        c = cs[a]
        idx = np.arange(c.shape[a]) % 2
        shape = [1, 1, 1]; shape[a] = -1
        inside = (idx == 0).reshape(shape)       # link from even to odd index: inside the aggregate
        dc -= 2 * agg(c * inside)
        cc.append(agg(c * (~inside)))
    return dict(diag=dc, c=cc, mask=agg(mask.astype(float)) > 0)

def rbgs(L, x, b, order):
    d = L["diag"]; inv = np.where(d > 0, 1.0 / np.where(d > 0, d, 1), 0.0)
    I, J, K = np.indices(d.shape)
    col = (I + J + K) & 1
    for color in order:
        r = b - apply(L, x)
        m = col == color
        x[m] += (r * inv)[m]
    return x

import os
JW = float(os.environ.get("MG_OMEGA", "0.8"))
def jacobi(L, x, b, w, sweeps):
    w = JW
    d = L["diag"]; inv = np.where(d > 0, 1.0 / np.where(d > 0, d, 1), 0.0)
    for _ in range(sweeps):
        x = x + w * inv * (b - apply(L, x))
    return x

def restrict(r):
    pad = [(0, s % 2) for s in r.shape]
    r = np.pad(r, pad)
    csh = tuple(s // 2 for s in r.shape)
    return r.reshape(csh[0], 2, csh[1], 2, csh[2], 2).sum(axis=(1, 3, 5))

def prolong(e, sh):
    f = np.repeat(np.repeat(np.repeat(e, 2, 0), 2, 1), 2, 2)
    return f[:sh[0], :sh[1], :sh[2]]

def vcycle(levels, l, b, smoother, nu, alpha, coarse_sweeps):
    L = levels[l]
    x = np.zeros_like(b)
    if l == len(levels) - 1:
        if smoother == "rbgs":
            for _ in range(coarse_sweeps):
                x = rbgs(L, x, b, (0, 1))
            for _ in range(coarse_sweeps):
                x = rbgs(L, x, b, (1, 0))
        else:
            x = jacobi(L, x, b, 0.8, 2 * coarse_sweeps)
        return x
    for _ in range(nu):
        x = rbgs(L, x, b, (0, 1)) if smoother == "rbgs" else jacobi(L, x, b, 0.8, 1)
    r = b - apply(L, x)
    ec = vcycle(levels, l + 1, restrict(r), smoother, nu, alpha, coarse_sweeps)
    x = x + alpha * prolong(ec, b.shape) * L["mask"]
    for _ in range(nu):
        x = rbgs(L, x, b, (1, 0)) if smoother == "rbgs" else jacobi(L, x, b, 0.8, 1)
    return x

def bpx(levels, r, w=0.8, scales=None):
    """additive multilevel preconditioner: z = sum_l P_l (w D_l^-1) P_l^T r"""
    z = np.zeros_like(r)
    bl = r
    shapes = []
    contrib = []
    for l, L in enumerate(levels):
        d = L["diag"]; inv = np.where(d > 0, 1.0 / np.where(d > 0, d, 1), 0.0)
        sc = w if scales is None else scales[l]
        contrib.append(sc * inv * bl)
        shapes.append(bl.shape)
        if l + 1 < len(levels):
            bl = restrict(bl)
    e = contrib[-1]
    for l in range(len(levels) - 2, -1, -1):
        e = contrib[l] + prolong(e, shapes[l]) * levels[l]["mask"]
    return e

def pcg(L, b, M, tol, maxit):
    x = np.zeros_like(b); r = b.copy(); z = M(r); d = z.copy(); rz = (r * z).sum()
    nn = b.size
    for it in range(1, maxit + 1):
        q = apply(L, d)
        a = rz / (d * q).sum()
        x += a * d; r -= a * q
        l2 = np.sqrt((r * r).sum() / nn)
        if l2 < tol: return x, it, l2
        z = M(r); rz2 = (r * z).sum(); d = z + (rz2 / rz) * d; rz = rz2
    return x, maxit, l2

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    nlev = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    p = build(n)
    phi = p["phi"].copy()
    rho = np.where(p["reg"], QE * p["n0"], 0.0) * (1 + 0.05 * np.random.default_rng(0).standard_normal(phi.shape))
    g = p["g"]
    for newton in range(3):
        L0 = fine_level(p, phi)
        # residual R of the nonlinear equations with face neighbours folded
        ph = phi.copy()
        # faces mirror inner values (approximation good enough for the prototype): pad by copying
        ph[0, :, :] = np.where(p["face"][0], ph[1], ph[0]); ph[-1] = np.where(p["face"][-1], ph[-2], ph[-1])
        ph[:, 0] = np.where(p["face"][:, 0], ph[:, 1], ph[:, 0]); ph[:, -1] = np.where(p["face"][:, -1], ph[:, -2], ph[:, -1])
        ph[:, :, -1] = np.where(p["face"][:, :, -1], ph[:, :, -2], ph[:, :, -1])
        lap = np.zeros_like(ph)
        lap[1:-1, 1:-1, 1:-1] = (g[0] * (ph[2:, 1:-1, 1:-1] + ph[:-2, 1:-1, 1:-1]) + g[1] * (ph[1:-1, 2:, 1:-1] + ph[1:-1, :-2, 1:-1])
                                 + g[2] * (ph[1:-1, 1:-1, 2:] + ph[1:-1, 1:-1, :-2]) - 2 * g.sum() * ph[1:-1, 1:-1, 1:-1])
        ne = p["n0"] * np.exp(phi / p["Te"])
        R = np.where(p["reg"], lap + (rho - QE * ne) / EPS0, 0.0)
        print("newton %d: |R| = %.3e" % (newton, np.sqrt((R * R).sum() / R.size)))
        levels = [L0]
        for _ in range(nlev - 1):
            levels.append(coarsen(levels[-1]))
        inv = np.where(L0["diag"] > 0, 1.0 / np.where(L0["diag"] > 0, L0["diag"], 1), 0.0)
        t = time.time(); y, it, l2 = pcg(L0, R, lambda r: inv * r, 1e-4, 3000)
        print("  jacobi-pcg: %d its (l2 %.2e) %.1fs" % (it, l2, time.time() - t))
        import os
        variants = eval(os.environ.get("MG_VARIANTS", '(("rbgs", 1, 1.0, 4), ("jac", 1, 1.0, 4), ("jac", 2, 1.0, 4))'))
        for sm, nu, alpha, cs in variants:
            t = time.time()
            y2, it2, l22 = pcg(L0, R, lambda r: vcycle(levels, 0, r, sm, nu, alpha, cs), 1e-4, 200)
            print("  mg-pcg %s nu=%d alpha=%.1f levels=%d: %d its (l2 %.2e) %.1fs, |y-y2|/|y| = %.2e" % (
                sm, nu, alpha, nlev, it2, l22, time.time() - t, np.abs(y - y2).max() / np.abs(y).max()))
        for nl in (3, 4, 5):
            lv = levels[:1]
            while len(lv) < nl:
                lv.append(coarsen(lv[-1]))
            for w in (0.8, 1.0):
                t = time.time()
                y3, it3, l23 = pcg(L0, R, lambda r: bpx(lv, r, w), 1e-4, 600)
                print("  bpx-pcg levels=%d w=%.1f: %d its (l2 %.2e) %.1fs, |y-y3|/|y| = %.2e" % (nl, w, it3, l23, time.time() - t, np.abs(y - y3).max() / np.abs(y).max()))
        phi = phi + y2
