"""CPU tests of the N>1 design with torch.distributed/gloo, world_size 2, no GPU:

particles are sharded by index, every rank scatters ITS shard into a private full-mesh accumulator, the accumulators
are summed with one all-reduce, and only then divided by the node volumes.  The per-rank work is done here by the CPU
oracle (test infrastructure); what is under test is the sharding arithmetic the GPU path uses:
  * FP64: all-reduced shard deposits == the single-rank deposit up to summation order (1e-13);
  * fixed point (int64 multiples of 2^-shift, shift agreed through a max-all-reduce as in prepare_acc(),
    csrc/espic_particles.cu): the all-reduced result is BIT-IDENTICAL for 1 and 2 ranks and any particle order;
  * weak-scaling bookkeeping of bench.py: per-rank seeds differ, mpw follows the global particle count.
Also checks div_check (the exact-division identity the kernels rely on).
"""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import statefile as sf
from cases import orc


def fixed_point_deposit(w, part, shift):
    """Host restatement of the fixed-point scatter: each of the 8 weights is rounded to an int64 multiple of 2^-shift
    (Field::scatter order and left-to-right products, Field.h:177-184), then summed as integers."""
    ni, nj, nk = w.ni, w.nj, w.nk
    dh = w.dh
    acc = np.zeros(w.nn, dtype=np.int64)
    lc = [(part[c] - w.x0[c]) / dh[c] for c in range(3)]
    idx = [np.minimum(l.astype(np.int64), n - 2) for l, n in zip(lc, (ni, nj, nk))]
    d = [l - i for l, i in zip(lc, idx)]
    mpw = part[6]
    scale = math.ldexp(1.0, shift)
    u = (idx[2] * nj + idx[1]) * ni + idx[0]
    sj, sk = ni, ni * nj
    ai, aj, ak = 1 - d[0], 1 - d[1], 1 - d[2]
    di, dj, dk = d
    terms = [(0, ai, aj, ak), (1, di, aj, ak), (1 + sj, di, dj, ak), (sj, ai, dj, ak),
             (sk, ai, aj, dk), (1 + sk, di, aj, dk), (1 + sj + sk, di, dj, dk), (sj + sk, ai, dj, dk)]
    for off, a, b, c in terms:
        q = np.rint(mpw * a * b * c * scale).astype(np.int64)
        np.add.at(acc, u + off, q)
    return acc


def agree_shift(n_local, mpw_max, nranks, allreduce_max):
    """prepare_acc(): bound = max over ranks(np*mpw_max) * nranks < 2^e  ->  shift = min(62 - e, 62)"""
    bound = allreduce_max(float(max(n_local, 1)) * mpw_max) * nranks
    _, e = math.frexp(bound)
    return min(62 - e, 62)


def _worker(rank, world, port, path):
    import torch.distributed as dist
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = np.load(path)
    part = d["part"]
    n = part.shape[1]
    lo, hi = rank * n // world, (rank + 1) * n // world        # index sharding
    w = cases.sphere_world(int(d["ni"]), int(d["nj"]), int(d["nk"]))
    shard = np.ascontiguousarray(part[:, lo:hi])

    # FP64 path: private accumulator (number density * node volume), all-reduce, divide
    sp = orc.Species(w, 16 * orc.AMU, orc.QE, 50.0, cap=max(16, 2 * shard.shape[1]))
    sp.set_particles(shard)
    sp.compute_number_density()
    acc = torch.from_numpy(sp.den * w.node_vol)
    dist.all_reduce(acc)
    den = np.where(w.node_vol != 0, acc.numpy() / w.node_vol, 0.0)

    # fixed-point path
    def allreduce_max(v):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])
    shift = agree_shift(shard.shape[1], float(part[6].max()), world, allreduce_max)
    iacc = torch.from_numpy(fixed_point_deposit(w, shard, shift))
    dist.all_reduce(iacc)
    if rank == 0:
        np.savez(path + ".out.npz", den=den, iacc=iacc.numpy(), shift=shift)
    dist.barrier()
    dist.destroy_process_group()


def test_index_sharding_with_density_allreduce(tmp_path):
    import torch.multiprocessing as mp
    ni, nj, nk, n = 9, 9, 13, 20001         # odd count: ragged shards
    w, sp = cases.sphere_case(seed=31, ni=ni, nj=nj, nk=nk, n=n)
    part = sp.particles()
    path = str(tmp_path / "in.npz")
    np.savez(path, part=part, ni=ni, nj=nj, nk=nk)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, path), nprocs=2, join=True)
    out = np.load(path + ".out.npz")
    sp.compute_number_density()
    scale = np.abs(sp.den).max()
    assert np.abs(out["den"] - sp.den).max() <= 1e-13 * scale, "FP64 shard deposits + all-reduce == single-rank deposit"
    # single rank, fixed point, same shift as the 2-rank run would agree on for ONE rank holding everything:
    shift2 = int(out["shift"])
    one = fixed_point_deposit(w, part, shift2)
    assert np.array_equal(one, out["iacc"]), "fixed-point deposit must be bit-identical for 1 and 2 ranks"
    rev = fixed_point_deposit(w, np.ascontiguousarray(part[:, ::-1]), shift2)
    assert np.array_equal(one, rev), "and for any particle order"
    den_fixed = np.where(w.node_vol != 0, one * math.ldexp(1.0, -shift2) / w.node_vol, 0.0)
    assert np.abs(den_fixed - sp.den).max() <= 1e-10 * scale
    # no overflow headroom problem: the largest accumulated value stays below 2^62
    assert np.abs(out["iacc"]).max() < 2 ** 62


def test_shift_is_rank_count_independent_bound():
    """The agreed shift only shrinks as ranks are added (the bound covers the whole system), never overflows."""
    for nranks in (1, 2, 4, 8):
        shift = agree_shift(200_000_000, 80.0, nranks, lambda v: v)
        total = 200_000_000 * nranks * 80.0
        assert total * math.ldexp(1.0, shift) < 2.0 ** 62


def test_bench_weak_scaling_bookkeeping():
    """bench.py: per-rank particle seeds differ, mpw = n0 * V / (particles_per_gpu * world)."""
    src = open(os.path.join(sf.ROOT, "bench.py")).read()
    assert "seed + 1000 * rank" in src and "n_total = n_local * world" in src and "mpw = N0 * box_vol / n_total" in src


def test_exact_division_identity():
    exe = os.path.join(sf.ROOT, "oracle", "div_check")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(sf.ROOT, "oracle")])
    r = subprocess.run([exe, "100000"], capture_output=True, text=True)
    assert r.returncode == 0 and "mismatches=0" in r.stdout, r.stdout
