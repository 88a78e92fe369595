"""Seeded cases for the ch4 surface-interaction advance (Species::advance(neutrals, spherium), ch4/Species.cpp:8-100)."""
import numpy as np

import cases
from cases import orc, AMU, QE

DT = 2e-6            # long step: a 7 km/s particle crosses 1.4 cm (2.8 cm on its first, doubled step), so many hit the sphere
SPH_C, SPH_R = (0.0, 0.0, 0.15), 0.05


def make_world(seed=5, dims=(9, 9, 13), amp=30.0):
    w, _ = cases.sphere_case(seed=seed, ni=dims[0], nj=dims[1], nk=dims[2], n=10, amp=amp)
    return w


def make_particles(w, seed, n, mpw, v_th=3000.0):
    """Half of the particles start in a shell just outside the sphere flying roughly inwards (bounces / impacts), a few next to
    the outer walls flying out (kills), the rest anywhere.  Returns (soa[7,n], pdt[n]): a third carry dt = DT (particles added
    since the last advance, ch4/Species.h:65), the others 0."""
    rng = np.random.default_rng(seed)
    part = cases.random_particles(w, rng, n, v_drift=7000.0, v_th=v_th, mpw=mpw, near_walls=0.1)
    m = n // 2
    d = rng.normal(size=(3, m))
    d /= np.linalg.norm(d, axis=0)
    r = SPH_R * rng.uniform(1.002, 1.35, size=m)
    c = np.array(SPH_C)[:, None]
    part[0:3, n - m:] = c + d * r
    speed = rng.uniform(3000.0, 12000.0, size=m)
    jitter = rng.normal(0, 0.4, size=(3, m))
    dirn = -d + jitter
    dirn /= np.linalg.norm(dirn, axis=0)
    part[3:6, n - m:] = dirn * speed
    pdt = np.where(rng.uniform(size=n) < 1 / 3, DT, 0.0)
    return part, pdt


def species_triplet(w, charge, mpw0=(5.0, 2.0, 20.0), cap=64):
    """advancing species, neutrals, sputtered material"""
    a = orc.Species(w, 16 * AMU, charge, mpw0[0], cap=cap)
    nt = orc.Species(w, 16 * AMU, 0.0, mpw0[1], cap=cap)
    sp = orc.Species(w, 100 * AMU, 0.0, mpw0[2], cap=cap)
    return a, nt, sp
