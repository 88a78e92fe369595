#!/usr/bin/env python
"""Per-kernel timing of the particle path on the bench workload (CUDA events on the engine's stream).
   python scripts/kernel_bench.py [--particles 2e8] [--mesh 128] [--reps 5]
Prints one line per variant: ms, particles/s, algorithmic GB/s, fraction of the measured HBM peak."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=float, default=2e8)
    ap.add_argument("--mesh", type=int, default=128)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch
    es = B.load_espic()
    dev = torch.device("cuda", 0)
    n = int(args.particles)
    mpw = B.N0 * 0.016 / n
    e = es.Engine(args.mesh, args.mesh, args.mesh, B.X0, B.XM, device=0)
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    e.add_sphere(*B.SPHERE); e.add_inlet(); e.set_reference_values(B.PHI0, B.TE0, B.N0)
    sp = e.add_species(16 * B.AMU, B.QE, mpw, capacity=n + 1024)
    t = B.make_particles_device(torch, n, 12345, mpw, dev)
    e.upload_device(sp, [t[c].data_ptr() for c in range(7)], n, mpw)
    e.sync(); del t; torch.cuda.empty_cache()
    e.sort_by_cell(sp); e.deposit(sp, es.DEPOSIT_FP64); e.compute_charge_density()
    e.solve(es.SOLVE_QN, 1, 1.0); e.compute_ef(); e.sync()
    peak, _ = B.measured_peak_gbs()

    def timed(name, fn, bytes_pp, reps=args.reps, setup=None):
        if args.only and args.only not in name:
            return
        ms = []
        for r in range(reps):
            if setup:
                setup()
            np_ = e.count(sp)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
            ms.append((a.elapsed_time(b), np_))
        m, np_ = sorted(ms)[len(ms) // 2]
        gbs = bytes_pp * np_ / (m * 1e-3) / 1e9
        print("%-44s %9.3f ms  %10.3e part/s  %8.1f GB/s  %5.1f%% of %.0f   (n=%d; all: %s)" % (
            name, m, np_ / (m * 1e-3), gbs, 100 * gbs / peak, peak, np_, " ".join("%.2f" % x[0] for x in ms)), flush=True)

    NC = es.PUSH_NO_COMPACT
    timed("sort_by_cell", lambda: e.sort_by_cell(sp), 112 + 24)
    timed("deposit fp64 (sorted)", lambda: e.deposit(sp, es.DEPOSIT_FP64), 32)
    timed("deposit fixed (sorted)", lambda: e.deposit(sp, es.DEPOSIT_FIXED), 32)
    timed("push only, no compaction (sorted, 1st)", lambda: e.push(sp, B.DT, es.WALL_ABSORB, NC), 104, reps=1)
    timed("push+deposit fused, no compaction", lambda: e.push(sp, B.DT, es.WALL_ABSORB, NC | es.PUSH_FUSE_DEPOSIT), 104, reps=1)
    e.sort_by_cell(sp)
    timed("push+deposit fused + removal (sorted, 1st)", lambda: e.push(sp, B.DT, es.WALL_ABSORB, es.PUSH_FUSE_DEPOSIT), 104, reps=1)
    timed("push+deposit fused + removal (steps 2..)", lambda: e.push(sp, B.DT, es.WALL_ABSORB, es.PUSH_FUSE_DEPOSIT), 104, reps=4)
    timed("deposit fp64 (5 steps after sort)", lambda: e.deposit(sp, es.DEPOSIT_FP64), 32)
    timed("push only + removal (5 steps after sort)", lambda: e.push(sp, B.DT, es.WALL_ABSORB, 0), 104, reps=2)
    timed("diag", lambda: e.diag(sp), 32)
    timed("sample_moments (5+ steps after sort)", lambda: e.sample_moments(sp), 56, reps=3)
    e.sort_by_cell(sp)
    timed("sample_moments (sorted)", lambda: e.sample_moments(sp), 56, reps=3)
    # ch4 operators on the same particles (the ions stand in for a neutral gas: a second species with charge 0)
    if not args.only or "ch4" in args.only:
        n4 = min(n, 50_000_000)
        gas = e.add_species(16 * B.AMU, 0.0, mpw, capacity=n4 + 1024)
        t = B.make_particles_device(torch, n4, 999, mpw, dev)
        e.upload_device(gas, [t[c].data_ptr() for c in range(7)], n4, mpw)
        e.sync(); del t; torch.cuda.empty_cache()
        sp_saved, sp = sp, gas
        step = [0]

        def surf():
            step[0] += 1
            e.push_surface(gas, B.DT, gas, gas, 7, 0, step[0])
        timed("ch4 push_surface neutral + removal", surf, 104, reps=3)
        sig = [1e-14]

        def dsmc():
            step[0] += 1
            cols, sig[0] = e.dsmc_mex(gas, B.DT, sig[0], 7, 1, step[0])
        timed("ch4 dsmc_mex (unsorted)", dsmc, 48, reps=3)
        timed("ch4 compute_mpc", lambda: e.compute_mpc(gas), 24, reps=3)
        timed("ch4 mcc_cex", lambda: e.mcc_cex(gas, sp_saved, B.DT, 7, 2, 1), 72, reps=3)
        sp = sp_saved
    print("launches:", e.kernel_launches())


if __name__ == "__main__":
    main()
