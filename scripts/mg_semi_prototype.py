"""numpy prototype: aggregation multigrid with SEMI-coarsening on the first level (2x2x1 aggregates: the mesh spacing in z
is twice that in x,y, so the stencil is 4:4:1 anisotropic and point smoothers with full coarsening converge slowly)."""
import sys, time
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__file__))
import mg_prototype as M

def coarsen_shape(L, f):
    d, cs = L["diag"], L["c"]
    pad = [(0, (-s) % fa) for s, fa in zip(d.shape, f)]
    P = lambda a: np.pad(a, pad)
    d = P(d); cs = [P(c) for c in cs]; mask = P(L["mask"])
    csh = tuple(s // fa for s, fa in zip(d.shape, f))
    def agg(a): return a.reshape(csh[0], f[0], csh[1], f[1], csh[2], f[2]).sum(axis=(1, 3, 5))
    dc = agg(d); cc = []
    for a in range(3):
        c = cs[a]
        if f[a] == 1:
            cc.append(agg(c)); continue
        idx = np.arange(c.shape[a]) % 2
        shape = [1, 1, 1]; shape[a] = -1
        inside = (idx == 0).reshape(shape)
        dc -= 2 * agg(c * inside)
        cc.append(agg(c * (~inside)))
    return dict(diag=dc, c=cc, mask=agg(mask.astype(float)) > 0, f=f)

def restrict(r, f):
    pad = [(0, (-s) % fa) for s, fa in zip(r.shape, f)]
    r = np.pad(r, pad)
    csh = tuple(s // fa for s, fa in zip(r.shape, f))
    return r.reshape(csh[0], f[0], csh[1], f[1], csh[2], f[2]).sum(axis=(1, 3, 5))

def prolong(e, sh, f):
    out = np.repeat(np.repeat(np.repeat(e, f[0], 0), f[1], 1), f[2], 2)
    return out[:sh[0], :sh[1], :sh[2]]

def vcycle(levels, l, b, nu, w, cs):
    L = levels[l]
    x = np.zeros_like(b)
    if l == len(levels) - 1:
        return M.jacobi(L, x, b, w, cs)
    x = M.jacobi(L, x, b, w, nu)
    r = b - M.apply(L, x)
    f = levels[l + 1]["f"]
    ec = vcycle(levels, l + 1, restrict(r, f), nu, w, cs)
    x = x + prolong(ec, b.shape, f) * L["mask"]
    return M.jacobi(L, x, b, w, nu)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    p = M.build(n)
    phi = p["phi"].copy()
    rho = np.where(p["reg"], M.QE * p["n0"], 0.0) * (1 + 0.05 * np.random.default_rng(0).standard_normal(phi.shape))
    g = p["g"]
    for newton in range(3):
        L0 = M.fine_level(p, phi)
        ph = phi.copy()
        ph[0] = np.where(p["face"][0], ph[1], ph[0]); ph[-1] = np.where(p["face"][-1], ph[-2], ph[-1])
        ph[:, 0] = np.where(p["face"][:, 0], ph[:, 1], ph[:, 0]); ph[:, -1] = np.where(p["face"][:, -1], ph[:, -2], ph[:, -1])
        ph[:, :, -1] = np.where(p["face"][:, :, -1], ph[:, :, -2], ph[:, :, -1])
        lap = np.zeros_like(ph)
        lap[1:-1, 1:-1, 1:-1] = (g[0] * (ph[2:, 1:-1, 1:-1] + ph[:-2, 1:-1, 1:-1]) + g[1] * (ph[1:-1, 2:, 1:-1] + ph[1:-1, :-2, 1:-1])
                                 + g[2] * (ph[1:-1, 1:-1, 2:] + ph[1:-1, 1:-1, :-2]) - 2 * g.sum() * ph[1:-1, 1:-1, 1:-1])
        ne = p["n0"] * np.exp(phi / p["Te"])
        R = np.where(p["reg"], lap + (rho - M.QE * ne) / M.EPS0, 0.0)
        for name, shapes in (("full 2x2x2", [(2, 2, 2)] * 3), ("semi 2x2x1 first", [(2, 2, 1), (2, 2, 2), (2, 2, 2)]),
                             ("semi first two", [(2, 2, 1), (2, 2, 1), (2, 2, 2)]), ("semi + 4 levels", [(2, 2, 1), (2, 2, 2), (2, 2, 2), (2, 2, 2)])):
            lv = [L0]
            for f in shapes:
                lv.append(coarsen_shape(lv[-1], f))
            t = time.time(); y, it, l2 = M.pcg(L0, R, lambda r: vcycle(lv, 0, r, 1, 0.8, 6), 1e-4, 300)
            print("newton %d  %-18s: %d its %.1fs  coarse sizes %s" % (newton, name, it, time.time() - t, [l["diag"].shape for l in lv[1:]]))
        phi = phi + y
