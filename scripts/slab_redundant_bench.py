#!/usr/bin/env python
"""Slab-decomposed multigrid solve: time per CG iteration with different numbers of coarse levels solved redundantly by every
rank (ESPIC_MG_SLAB_REDUNDANT_NODES; the library reads it at every solve, so one process can switch it).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/slab_redundant_bench.py [--mesh 256]
Every solve starts from the same quasi-neutral guess on the same charge density, so all variants run the same iterations."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, default=256)
    ap.add_argument("--particles", type=float, default=2e7)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--limits", default="0,9000,70000,600000")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    es = B.load_espic()
    n = int(args.particles)
    mpw = B.N0 * 0.016 / (n * world)
    e = es.Engine(args.mesh, args.mesh, args.mesh, B.X0, B.XM, device=local)
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    e.add_sphere(*B.SPHERE); e.add_inlet(); e.set_reference_values(B.PHI0, B.TE0, B.N0)
    sp = e.add_species(16 * B.AMU, B.QE, mpw, capacity=n + 1024)
    uid = [e.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    e.comm_init(rank, world, uid[0])
    t = B.make_particles_device(torch, n, 12345 + rank, mpw, dev)
    e.upload_device(sp, [t[c].data_ptr() for c in range(7)], n, mpw)
    e.sync(); del t
    e.deposit(sp, es.DEPOSIT_FP64); e.compute_charge_density()
    for limit in [int(x) for x in args.limits.split(",")]:
        os.environ["ESPIC_MG_SLAB_REDUNDANT_NODES"] = str(limit)
        ms, its = [], None
        for r in range(args.reps + 1):
            e.solve(es.SOLVE_QN, 1, 1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier(); torch.cuda.synchronize(); a.record()
            info = e.solve(es.SOLVE_PCG_MG_SLAB, 5000, 1e-4)
            b.record(); torch.cuda.synchronize()
            if r > 0:
                ms.append(a.elapsed_time(b))
            its = info
        tm = torch.tensor([sorted(ms)[len(ms) // 2]], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        if rank == 0:
            li = its["lin_iters"] if isinstance(its, dict) else getattr(its, "lin_iters", -1)
            print("ranks %d mesh %d^3 redundant<=%-7d: %8.2f ms per solve, %s -> %.1f us per CG iteration" % (
                world, args.mesh, limit, tm.item(), its, 1e3 * tm.item() / max(li, 1)), flush=True)
    e.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
