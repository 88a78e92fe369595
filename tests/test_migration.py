"""Spatial domain decomposition with particle migration (ch9/MPI World::initMPIDomain + Species::transferParticles,
ch9/MPI/include/World.h:73-128, ch9/MPI/src/Species.cpp:204-313).  ch9/MPI itself cannot be built here (no MPI toolchain), so
parity is anchored on the property the decomposition must have: a run cut into k-slabs with migration is the SAME run --
the union of the parts' particles equals the single-domain particle set bit for bit after every step, every particle sits
on the part that owns its cell, and the summed scatter equals the single-domain scatter.  The order inside every part is the
deterministic one of tests/migration_model.py.

CPU (-m "not gpu"): the model with the oracle as the per-part engine, and the exchange protocol of espic_migrate
(all-gathered count matrix, per-peer SoA segments, arrivals by ascending source) over gloo with world size 2.
GPU (-m gpu): R parts as R contexts on ONE device through espic_migrate_pack / espic_migrate_segment /
espic_species_upload_device, bit for bit against the model; the NCCL path (espic_migrate) is in tests/test_multigpu.py.
"""
import os

import numpy as np
import pytest

import cases
import migration_model as mm
import statefile as sf
from cases import orc


def _oracle_parts(w, sp_all, kb):
    z0, dhz = w.x0[2], w.dh[2]
    parts = mm.split_by_owner(sp_all.particles(), z0, dhz, w.nk, kb)
    sps = []
    for p in parts:
        s = orc.Species(w, sp_all.mass, sp_all.charge, 50.0, cap=max(16, 4 * sp_all.particles().shape[1]))
        s.set_particles(p)
        sps.append(s)
    return sps


def _oracle_decomposed_steps(w, sps, kb, dt, steps, fused=False):
    """push every part with the oracle, migrate with the model; returns per-step lists of part arrays.
    fused: the push leaves its dead in place (mpw = 0) and the migration's single sweep removes them together with the leavers
    (espic_push(ESPIC_PUSH_MIGRATE) + espic_migrate); otherwise the push removes its dead first (espic_push + espic_migrate)."""
    hist = []
    for _ in range(steps):
        dead = None
        for s in sps:
            s.push_nocompact(dt) if fused else s.advance(dt)
        if fused:
            dead = [s.particles()[6] == 0 for s in sps]
        new, counts = mm.migrate([s.particles() for s in sps], w.x0[2], w.dh[2], w.nk, kb, dead)
        for s, p in zip(sps, new):
            s.set_particles(p)
        hist.append(([p.copy() for p in new], counts))
    return hist


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("kb,dt", [([0, 3, 9, 13], 1e-7), ([0, 2, 4, 13], 2e-5), ([0, 13], 1e-7)])
def test_decomposed_oracle_run_is_the_single_domain_run(kb, dt, fused):
    w, sp = cases.sphere_case(seed=77, ni=9, nj=8, nk=14, n=6000, near_walls=0.1)
    one = orc.Species(w, sp.mass, sp.charge, 50.0, cap=4 * 6000)
    one.set_particles(sp.particles())
    sps = _oracle_parts(w, sp, kb)
    assert sum(s.particles().shape[1] for s in sps) == 6000
    moved, skipped = 0, 0
    for parts, counts in _oracle_decomposed_steps(w, sps, kb, dt, 4, fused):
        one.advance(dt)
        moved += int(counts.sum())
        skipped += int(counts[0, 2]) if len(kb) > 3 else 0
        allp = np.concatenate(parts, axis=1)
        assert np.array_equal(mm.canonical(allp).view(np.uint64), mm.canonical(one.particles()).view(np.uint64))
        for r, p in enumerate(parts):
            assert np.all(mm.owner_of(p[2], w.x0[2], w.dh[2], w.nk, kb) == r)
    if len(kb) > 2:
        assert moved > 0
    if dt > 1e-6:
        assert skipped > 0, "the long step must carry particles past the middle part"
    # summed scatter of the parts == single-domain scatter (ch9/MPI Field::updateBoundaries adds the shared planes)
    acc = np.zeros(w.nn)
    for s in sps:
        s.compute_number_density()
        acc += s.den * w.node_vol
        k_nodes = np.nonzero((s.den.reshape(w.nk, w.nj, w.ni) != 0).any(axis=(1, 2)))[0]
        r = sps.index(s)
        assert k_nodes.size == 0 or (k_nodes.min() >= kb[r] and k_nodes.max() <= kb[r + 1]), "a part scatters into its own planes only"
    one.compute_number_density()
    assert np.abs(acc - one.den * w.node_vol).max() <= 1e-13 * np.abs(one.den * w.node_vol).max()


def test_swap_remove_is_the_reference_loop():
    rng = np.random.default_rng(3)
    for n in (1, 2, 31, 32, 33, 257):
        for frac in (0.0, 0.3, 1.0):
            part = rng.normal(size=(7, n))
            dead = rng.random(n) < frac
            lst = [part[:, i] for i in range(n)]
            dl = list(dead)
            p, cnt = 0, n
            while p < cnt:                      # ch3/ver2/Species.cpp:36-46
                if dl[p]:
                    lst[p], dl[p] = lst[cnt - 1], dl[cnt - 1]
                    cnt -= 1
                    continue
                p += 1
            ref = np.array(lst[:cnt]).T.reshape(7, cnt)
            assert np.array_equal(mm.swap_remove(part, dead), ref)


def test_slab_bounds_helpers():
    """espic.slab_bounds / balanced_bounds: valid cuts (0 .. nk-1, strictly increasing), balanced shares"""
    es = __import__("engines")._espic()
    for nk, parts in ((14, 3), (128, 8), (9, 8), (3, 2)):
        kb = es.slab_bounds(nk, parts)
        assert kb[0] == 0 and kb[-1] == nk - 1 and all(b > a for a, b in zip(kb, kb[1:])) and len(kb) == parts + 1
    rng = np.random.default_rng(5)
    for cells, parts in ((127, 8), (127, 2), (13, 13), (40, 7)):
        w = rng.uniform(0.5, 1.5, cells)
        kb = es.balanced_bounds(w, parts)
        assert kb[0] == 0 and kb[-1] == cells and all(b > a for a, b in zip(kb, kb[1:])) and len(kb) == parts + 1
        share = np.array([w[a:b].sum() for a, b in zip(kb, kb[1:])]) / w.sum()
        assert np.abs(share - 1.0 / parts).max() <= 1.6 * w.max() / w.sum()
    assert es.balanced_bounds([0, 0, 0, 5, 0, 0], 3) == [0, 3, 4, 6]       # every part keeps a cell plane
    import bench
    w = bench.free_volume_per_cell_plane(128)
    assert abs(w.sum() - (0.2 * 0.2 * 0.4 - 4.0 / 3.0 * np.pi * 0.05 ** 3)) < 1e-12


# ---- the exchange protocol of espic_migrate over gloo, world size 2 ------------------------------------------------------

def _gloo_worker(rank, world, port, path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = np.load(path)
    kb = [int(k) for k in d["kb"]]
    w = cases.sphere_world(int(d["ni"]), int(d["nj"]), int(d["nk"]))
    w.phi[:] = d["phi"]
    w.compute_ef()
    dt = float(d["dt"])
    mine = mm.split_by_owner(d["part"], w.x0[2], w.dh[2], w.nk, kb)[rank]
    s = orc.Species(w, 16 * orc.AMU, orc.QE, 50.0, cap=4 * d["part"].shape[1])
    s.set_particles(mine)
    for _ in range(int(d["steps"])):
        s.advance(dt)
        p = s.particles()
        own = mm.owner_of(p[2], w.x0[2], w.dh[2], w.nk, kb)
        # pack: one SoA segment per destination, leavers in particle order
        segs = {dst: np.ascontiguousarray(p[:, own == dst]) for dst in range(world) if dst != rank}
        # counts matrix M[src][dst]: every rank contributes its row, the all-gather returns all rows
        row = torch.zeros(world, dtype=torch.int64)
        for dst, sg in segs.items():
            row[dst] = sg.shape[1]
        rows = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(rows, row)
        M = torch.stack(rows).numpy()
        # one send and one matching receive per peer with a non-zero count; arrivals land behind the old particles by ascending source
        recv = {src: torch.zeros((7, int(M[src, rank])), dtype=torch.float64) for src in range(world) if src != rank}
        reqs = []
        for peer in range(world):
            if peer == rank:
                continue
            if M[rank, peer] > 0:
                reqs.append(dist.isend(torch.from_numpy(segs[peer]), peer))
            if M[peer, rank] > 0:
                reqs.append(dist.irecv(recv[peer], peer))
        for q in reqs:
            q.wait()
        # arrivals behind the last slot, then ONE swap-with-last sweep over old + new closes the leavers' holes
        full = np.concatenate([p] + [recv[src].numpy() for src in sorted(recv)], axis=1)
        gone = np.concatenate([own != rank, np.zeros(full.shape[1] - p.shape[1], dtype=bool)])
        s.set_particles(mm.swap_remove(full, gone))
    np.save(path + ".rank%d.npy" % rank, s.particles())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_protocol_gloo(tmp_path):
    import torch.multiprocessing as mp
    ni, nj, nk, n, steps, dt = 9, 8, 14, 5001, 3, 3e-6
    kb = [0, 6, 13]
    w, sp = cases.sphere_case(seed=78, ni=ni, nj=nj, nk=nk, n=n)
    path = str(tmp_path / "in.npz")
    np.savez(path, part=sp.particles(), phi=w.phi, ni=ni, nj=nj, nk=nk, kb=kb, steps=steps, dt=dt)
    mp.spawn(_gloo_worker, args=(2, 29300 + os.getpid() % 2000, path), nprocs=2, join=True)
    got = [np.load(path + ".rank%d.npy" % r) for r in range(2)]
    sps = _oracle_parts(w, sp, kb)
    want = _oracle_decomposed_steps(w, sps, kb, dt, steps)[-1][0]
    for r in range(2):
        assert np.array_equal(got[r].view(np.uint64), want[r].view(np.uint64)), "rank %d: particles and their order" % r


# ---- GPU: R parts on one device ----------------------------------------------------------------------------------------------

def _gpu_parts(st, kb):
    from engines import GpuEngine
    part = st.species[0]["part"]
    z0, dhz = st.x0[2], (st.xm[2] - st.x0[2]) / (st.nk - 1)
    split = mm.split_by_owner(part, z0, dhz, st.nk, kb)
    engines = []
    for r, p in enumerate(split):
        st.species[0]["part"] = p
        g = GpuEngine(st)
        g.e.set_domain(len(kb) - 1, r, kb)
        engines.append(g)
    st.species[0]["part"] = part
    return engines


def _gpu_migrate_on_one_device(engines):
    """espic_migrate with the transport done by hand: pack everywhere, append the segments by ascending source, removal sweep"""
    R = len(engines)
    counts = [g.e.migrate_pack(g.species[0]) for g in engines]
    for g in engines:
        g.e.sync()
    for dst in range(R):
        for src in range(R):
            if src == dst or counts[src][dst] == 0:
                continue
            ptr, cnt = engines[src].e.migrate_segment(dst)
            assert cnt == counts[src][dst]
            engines[dst].e.upload_device(engines[dst].species[0], [ptr + 8 * q * cnt for q in range(7)], cnt, 0.0, append=True)
    for g in engines:
        g.e.migrate_finish(g.species[0])
        g.e.sync()
    return np.array(counts)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("kb,dt,n", [([0, 3, 9, 13], 1e-7, 40000), ([0, 2, 4, 13], 2e-5, 20011), ([0, 1, 2, 3, 5, 8, 11, 12, 13], 4e-6, 30000),
                                     ([0, 5, 13], 3e-6, 300001)])
def test_gpu_parts_on_one_device_match_the_model(kb, dt, n, fused):
    es = __import__("engines")._espic()
    w, sp = cases.sphere_case(seed=79, ni=9, nj=8, nk=14, n=n, near_walls=0.1)
    st = sf.state_from_oracle(w, [sp], dt)
    engines = _gpu_parts(st, kb)
    sps = _oracle_parts(w, sp, kb)
    hist = _oracle_decomposed_steps(w, sps, kb, dt, 4, fused)
    one = __import__("engines").GpuEngine(st)
    total = 0
    for step, (want, want_counts) in enumerate(hist):
        for g in engines:
            g.e.push(g.species[0], dt, es.WALL_ABSORB, es.PUSH_MIGRATE if fused else 0)
            if fused:
                with pytest.raises(es.EspicError):          # dead and leavers are still in the arrays
                    g.e.deposit(g.species[0], es.DEPOSIT_FP64)
        counts = _gpu_migrate_on_one_device(engines)
        assert np.array_equal(counts, want_counts), step
        total += int(counts.sum())
        for r, g in enumerate(engines):
            got = g.e.download(g.species[0])
            assert got.shape == want[r].shape, (step, r, got.shape, want[r].shape)
            assert np.array_equal(got.view(np.uint64), want[r].view(np.uint64)), "step %d part %d: particles and their order" % (step, r)
        one.e.push(one.species[0], dt, es.WALL_ABSORB, 0)
        allp = np.concatenate([g.e.download(g.species[0]) for g in engines], axis=1)
        assert np.array_equal(mm.canonical(allp).view(np.uint64), mm.canonical(one.e.download(one.species[0])).view(np.uint64))
    assert total > 0
    # the parts' scatters add up to the single-domain scatter (summation order only)
    acc = np.zeros(w.nn)
    for g in engines:
        g.e.deposit(g.species[0], es.DEPOSIT_FP64)
        acc += g.e.field(es.DEN, g.species[0]) * w.node_vol
    one.e.deposit(one.species[0], es.DEPOSIT_FP64)
    ref = one.e.field(es.DEN, one.species[0]) * w.node_vol
    assert np.abs(acc - ref).max() <= 1e-12 * np.abs(ref).max()
    for g in engines + [one]:
        g.e.close()


@pytest.mark.gpu
def test_gpu_migrate_edge_cases():
    """empty parts, a part that loses everything, no leavers at all, and the error paths of the C ABI"""
    es = __import__("engines")._espic()
    from engines import GpuEngine
    w, sp = cases.sphere_case(seed=80, ni=9, nj=8, nk=14, n=3000)
    st = sf.state_from_oracle(w, [sp], 1e-7)
    kb = [0, 6, 13]
    engines = _gpu_parts(st, kb)
    # no push in between: nobody leaves
    counts = _gpu_migrate_on_one_device(engines)
    assert counts.sum() == 0
    n0 = [g.e.count(g.species[0]) for g in engines]
    # hand part 0 everything part 1 owns: all of it must leave again, in order
    p1 = engines[1].e.download(engines[1].species[0])
    before0 = engines[0].e.download(engines[0].species[0])
    engines[0].e.upload(engines[0].species[0], p1, append=True)
    engines[1].e.upload(engines[1].species[0], np.zeros((7, 0)))
    assert engines[1].e.count(engines[1].species[0]) == 0
    counts = _gpu_migrate_on_one_device(engines)
    assert counts[0][1] == n0[1] and counts[1][0] == 0
    assert np.array_equal(engines[1].e.download(engines[1].species[0]).view(np.uint64), p1.view(np.uint64))
    assert np.array_equal(engines[0].e.download(engines[0].species[0]).view(np.uint64), before0.view(np.uint64))
    # a part holding only foreign particles ends empty
    engines[1].e.upload(engines[1].species[0], before0)
    engines[0].e.upload(engines[0].species[0], np.zeros((7, 0)))
    counts = _gpu_migrate_on_one_device(engines)
    assert counts[1][0] == n0[0] and engines[1].e.count(engines[1].species[0]) == 0
    assert np.array_equal(engines[0].e.download(engines[0].species[0]).view(np.uint64), before0.view(np.uint64))
    # error paths
    g = GpuEngine(st)
    with pytest.raises(es.EspicError):
        g.e.parts = 1
        g.e.migrate_pack(g.species[0])          # no domain set
    with pytest.raises(es.EspicError):
        g.e.set_domain(2, 0, [0, 6, 12])        # must end at nk-1 cells
    with pytest.raises(es.EspicError):
        g.e.set_domain(2, 0, [0, 0, 13])        # empty part
    g.e.set_domain(2, 0, [0, 6, 13])
    with pytest.raises(es.EspicError):
        g.e.migrate(g.species[0])               # no communicator of that shape
    for x in engines + [g]:
        x.e.close()


def _gpu_migrate_species_on_one_device(engines, q):
    """as _gpu_migrate_on_one_device, for species index q of every part"""
    R = len(engines)
    counts = [g.e.migrate_pack(g.species[q]) for g in engines]
    for dst in range(R):
        for src in range(R):
            if src == dst or counts[src][dst] == 0:
                continue
            ptr, cnt = engines[src].e.migrate_segment(dst)
            engines[dst].e.upload_device(engines[dst].species[q], [ptr + 8 * c * cnt for c in range(7)], cnt, 0.0, append=True)
    for g in engines:
        g.e.migrate_finish(g.species[q])
    return np.array(counts)


@pytest.mark.gpu
def test_gpu_two_species_push_all_then_migrate_all():
    """the kill / leave bits of a MIGRATE push belong to the SPECIES.  The natural loop
    `for sp: push(MIGRATE)` then `for sp: migrate` must give, per species, exactly the run that pushes and migrates one
    species at a time; calls that would clobber a pending migration of the same species fail instead of corrupting it."""
    es = __import__("engines")._espic()
    from engines import GpuEngine
    kb, dt = [0, 3, 9, 13], 2e-6
    w, spa = cases.sphere_case(seed=81, ni=9, nj=8, nk=14, n=30000, near_walls=0.1)
    _, spb = cases.sphere_case(seed=82, ni=9, nj=8, nk=14, n=17001, near_walls=0.2, v_drift=9000.0)
    st = sf.state_from_oracle(w, [spa, spb], dt)
    z0, dhz = st.x0[2], (st.xm[2] - st.x0[2]) / (st.nk - 1)
    full = [r["part"] for r in st.species]
    split = [mm.split_by_owner(p, z0, dhz, st.nk, kb) for p in full]

    def make_parts():
        out = []
        for r in range(len(kb) - 1):
            for q in range(2):
                st.species[q]["part"] = split[q][r]
            g = GpuEngine(st)
            g.e.set_domain(len(kb) - 1, r, kb)
            out.append(g)
        for q in range(2):
            st.species[q]["part"] = full[q]
        return out

    inter, seq = make_parts(), make_parts()
    moved = 0
    for step in range(4):
        # interleaved: every species pushed first, then every species migrated
        for g in inter:
            for q in range(2):
                g.e.push(g.species[q], dt, es.WALL_ABSORB, es.PUSH_MIGRATE)
        for g in inter[:1]:        # guards: the pending species refuses anything that would move or overwrite its particles
            for call in (lambda: g.e.push(g.species[0], dt, es.WALL_ABSORB, 0),
                         lambda: g.e.add_particles(g.species[0], full[0][:, :5].copy(), dt),
                         lambda: g.e.upload(g.species[0], full[0][:, :5].copy()),
                         lambda: g.e.sort_by_cell(g.species[0]),
                         lambda: g.e.inject_cold_beam(g.species[0], 7000.0, 1e10, dt, 1, 0, step)):
                with pytest.raises(es.EspicError):
                    call()
        for q in range(2):
            moved += int(_gpu_migrate_species_on_one_device(inter, q).sum())
        # sequential: one species at a time
        for q in range(2):
            for g in seq:
                g.e.push(g.species[q], dt, es.WALL_ABSORB, es.PUSH_MIGRATE)
            _gpu_migrate_species_on_one_device(seq, q)
        for r, (a, b) in enumerate(zip(inter, seq)):
            for q in range(2):
                pa, pb = a.e.download(a.species[q]), b.e.download(b.species[q])
                assert pa.shape == pb.shape, (step, r, q, pa.shape, pb.shape)
                assert np.array_equal(pa.view(np.uint64), pb.view(np.uint64)), "step %d part %d species %d" % (step, r, q)
    assert moved > 0
    # and both equal the single-domain run, species by species
    one = GpuEngine(st)
    for step in range(4):
        for q in range(2):
            one.e.push(one.species[q], dt, es.WALL_ABSORB, 0)
    for q in range(2):
        allp = np.concatenate([g.e.download(g.species[q]) for g in inter], axis=1)
        assert np.array_equal(mm.canonical(allp).view(np.uint64), mm.canonical(one.e.download(one.species[q])).view(np.uint64))
    for g in inter + seq + [one]:
        g.e.close()
