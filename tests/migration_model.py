"""Host model (numpy) of the spatial decomposition with particle migration (csrc/espic_migrate.cuh; ch9/MPI
World::initMPIDomain + Species::transferParticles, ch9/MPI/include/World.h:73-128, ch9/MPI/src/Species.cpp:204-313).

Test infrastructure: it states, independently of the CUDA code, which part owns a particle and in which ORDER every part holds its
particles after a migration -- stayers in the reference's swap-with-last removal order (ch3/ver2/Species.cpp:36-46), then the
arrivals by ascending source part, each source in its own particle order.  The GPU path must reproduce that order bit for bit.
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


def migrate(parts, z0, dhz, nk, kb):
    """parts: list of (7, n_r) arrays (after push + removal).  Returns (new parts, counts[src][dst])."""
    R = len(parts)
    own = [owner_of(p[2], z0, dhz, nk, kb) for p in parts]
    counts = np.zeros((R, R), dtype=np.int64)
    seg = [[None] * R for _ in range(R)]
    stay = []
    for r, p in enumerate(parts):
        for d in range(R):
            if d != r:
                seg[r][d] = p[:, own[r] == d]
                counts[r, d] = seg[r][d].shape[1]
        stay.append(swap_remove(p, own[r] != r))
    out = [np.concatenate([stay[r]] + [seg[s][r] for s in range(R) if s != r], axis=1) for r in range(R)]
    return out, counts


def split_by_owner(part, z0, dhz, nk, kb):
    own = owner_of(part[2], z0, dhz, nk, kb)
    return [np.ascontiguousarray(part[:, own == r]) for r in range(len(kb) - 1)]


def canonical(part):
    """columns in lexicographic order of their bit patterns: compares particle SETS"""
    v = np.ascontiguousarray(part).view(np.uint64)
    return part[:, np.lexsort(v[::-1])]
