"""Seeded synthetic cases shared by the oracle, golden and GPU parity tests."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402

QE, AMU, ME, EPS_0 = orc.QE, orc.AMU, orc.ME, orc.EPS_0


def sphere_world(ni=9, nj=9, nk=13, phi_sphere=-100.0, sphere=True, inlet=True):
    """ch3/ver2/Main.cpp:19-30 geometry on an (ni,nj,nk) mesh."""
    w = orc.World(ni, nj, nk, (-0.1, -0.1, 0.0), (0.1, 0.1, 0.4))
    if sphere:
        w.add_sphere((0.0, 0.0, 0.15), 0.05, phi_sphere)
    if inlet:
        w.add_inlet()
    w.set_reference_values(0.0, 1.5, 1e10)
    return w


def box_world(n=9):
    """ch2/Main.cpp:17-18 geometry."""
    return orc.World(n, n, n, (-0.1, -0.1, 0.0), (0.1, 0.1, 0.2))


def smooth_phi(w, rng, amp=5.0):
    """A smooth random potential on free nodes (fixed nodes keep their Dirichlet value)."""
    i, j, k = np.meshgrid(np.arange(w.ni), np.arange(w.nj), np.arange(w.nk), indexing="ij")
    a = rng.uniform(-1, 1, size=6)
    f = amp * (a[0] * np.sin(2 * np.pi * i / (w.ni - 1) + a[1]) * np.cos(np.pi * j / (w.nj - 1) + a[2])
               + a[3] * np.cos(2 * np.pi * k / (w.nk - 1) + a[4]) + a[5] * (i + j + k) / (w.ni + w.nj + w.nk))
    flat = np.zeros(w.nn)
    u = (k * w.ni * w.nj + j * w.ni + i).ravel()
    flat[u] = f.ravel()
    free = w.object_id == 0
    w.phi[free] = flat[free]


def random_particles(w, rng, n, v_drift=7000.0, v_th=300.0, mpw=50.0, outside_sphere=True, near_walls=0.0):
    """Uniform positions in the box (optionally rejecting the sphere), drifting Maxwellian-ish velocities.
    near_walls: fraction placed within one velocity-step of a wall/sphere so kills and reflections occur."""
    pos = w.x0[:, None] + rng.uniform(0, 1, size=(3, n)) * (w.xm - w.x0)[:, None]
    if near_walls > 0:
        m = int(n * near_walls)
        face = rng.integers(0, 6, size=m)
        for q in range(m):
            c, hi = face[q] % 3, face[q] // 3
            eps = rng.uniform(0, 2e-3) * (w.xm[c] - w.x0[c])
            pos[c, q] = (w.xm[c] - eps) if hi else (w.x0[c] + eps)
    if outside_sphere and w.sphere is not None:
        c, r, _ = w.sphere
        d2 = ((pos - np.array(c)[:, None]) ** 2).sum(0)
        bad = d2 <= (1.02 * r) ** 2
        # move offenders onto a shell just outside the sphere: they hit it within a few steps
        npos = np.array(c)[:, None] + (pos[:, bad] - np.array(c)[:, None]) / np.sqrt(d2[bad]) * (1.03 * r)
        pos[:, bad] = npos
    vel = rng.normal(0, v_th, size=(3, n))
    vel[2] += v_drift
    return np.vstack([pos, vel, np.full((1, n), float(mpw))])


def sphere_case(seed=1, ni=9, nj=9, nk=13, n=2000, **kw):
    rng = np.random.default_rng(seed)
    w = sphere_world(ni, nj, nk)
    smooth_phi(w, rng, amp=kw.pop("amp", 20.0))
    w.compute_ef()
    sp = orc.Species(w, 16 * AMU, QE, mpw0=kw.pop("mpw0", 50.0), cap=max(2 * n, 16))
    sp.set_particles(random_particles(w, rng, n, **kw))
    return w, sp


def box_case(seed=2, n=9, npart=3000):
    rng = np.random.default_rng(seed)
    w = box_world(n)
    smooth_phi(w, rng, amp=2.0)
    w.compute_ef()
    ions = orc.Species(w, 16 * AMU, QE, cap=2 * npart)
    eles = orc.Species(w, ME, -QE, cap=2 * npart)
    ions.set_particles(random_particles(w, rng, npart, v_drift=0.0, v_th=500.0, mpw=1e3, near_walls=0.2))
    # electrons are fast: many cross a wall within one dt of 2e-10*1e3
    eles.set_particles(random_particles(w, rng, npart, v_drift=0.0, v_th=2e5, mpw=1e3, near_walls=0.3))
    return w, [ions, eles]
