"""Host model (numpy) of the spatial decomposition with particle migration (csrc/espic_migrate.cuh; ch9/MPI
World::initMPIDomain + Species::transferParticles, ch9/MPI/include/World.h:73-128, ch9/MPI/src/Species.cpp:204-313).

Test infrastructure: it states, independently of the CUDA code, which part owns a particle and in which ORDER every part holds its
particles after a migration -- the arrivals are appended by ascending source part, each source in its own particle order, and
then one swap-with-last removal sweep (ch3/ver2/Species.cpp:36-46; ch9/MPI/src/Species.cpp:205-210) closes the holes of the
dead and of the leavers.  The GPU path must reproduce that order bit for bit.
"""
import numpy as np


def owner_of(z, z0, dhz, nk, kb):
    """part of the CELL the particle's z lies in: lc = (z - z0)/dh (World::XtoL), k = (int)lc clamped to nk-2 as in the
    gather/scatter of the engine; part r owns kb[r] <= k < kb[r+1]"""
    lc = (np.asarray(z) - z0) / dhz
    k = np.clip(lc.astype(np.int64), 0, nk - 2)
    return np.searchsorted(np.asarray(kb[1:-1]), k, side="right")


def swap_remove(part, dead):
    """for (p=0;p<np;p++) if (dead) { particles[p] = particles[np-1]; np--; p--; }  -- hole r (ascending) receives the r-th
    survivor counted from the end"""
    n = part.shape[1]
    L = n - int(dead.sum())
    holes = np.nonzero(dead[:L])[0]
    fillers = np.nonzero(~dead[L:])[0][::-1] + L
    out = part[:, :L].copy()
    out[:, holes] = part[:, fillers]
    return out


def migrate(parts, z0, dhz, nk, kb, dead=None):
    """parts: list of (7, n_r) arrays; dead: optional list of boolean masks (particles the push killed but has not removed
    yet: espic_push(ESPIC_PUSH_MIGRATE) / Species::moveKernel leave them in place).  The sequence of ch9/MPI Species::move
    (Species.cpp:189-213): leavers are packed per destination in particle order, arrivals are appended by ascending source,
    then ONE swap-with-last sweep closes the holes of the dead and of the leavers -- arrivals are the first fillers.
    Returns (new parts, counts[src][dst])."""
    R = len(parts)
    counts = np.zeros((R, R), dtype=np.int64)
    seg = [[None] * R for _ in range(R)]
    gone = []
    for r, p in enumerate(parts):
        dmask = np.zeros(p.shape[1], dtype=bool) if dead is None else np.asarray(dead[r], dtype=bool)
        own = owner_of(p[2], z0, dhz, nk, kb)
        leave = (own != r) & ~dmask
        for d in range(R):
            if d != r:
                seg[r][d] = p[:, leave & (own == d)]
                counts[r, d] = seg[r][d].shape[1]
        gone.append(dmask | leave)
    out = []
    for r, p in enumerate(parts):
        arrivals = [seg[s][r] for s in range(R) if s != r]
        full = np.concatenate([p] + arrivals, axis=1)
        mask = np.concatenate([gone[r], np.zeros(full.shape[1] - p.shape[1], dtype=bool)])
        out.append(swap_remove(full, mask))
    return out, counts


def split_by_owner(part, z0, dhz, nk, kb):
    own = owner_of(part[2], z0, dhz, nk, kb)
    return [np.ascontiguousarray(part[:, own == r]) for r in range(len(kb) - 1)]


def canonical(part):
    """columns in lexicographic order of their bit patterns: compares particle SETS"""
    v = np.ascontiguousarray(part).view(np.uint64)
    return part[:, np.lexsort(v[::-1])]
