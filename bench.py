#!/usr/bin/env python
"""bench.py -- particle-pushes/s per full PIC step (push + deposit + rho + Poisson + E) on the sphere case.

  python bench.py [--gpus N] [--steps K] [--warmup W]                 this repo's CUDA engine (N=1 default)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     one rank per GPU
  python bench.py --impl reference ...       the reference's own CPU implementation (oracle/_ref/ref_ch3,
                                             built from the unmodified ch3/ver2 sources) on the host cores

Workload (BASELINE.json configs[3], the configuration the metric is quoted on at 1/2/4/8 B200): 128^3 mesh,
sphere (0,0,0.15) r=0.05 at -100 V + inlet, O+ ions at n0=1e12 with v=(0,0,7000)+300*N(0,1) m/s, 2e8
macroparticles PER GPU (weak scaling: particles are sharded by index, each rank deposits its shard, the density
is summed with one NCCL all-reduce, the field solve is replicated), Boltzmann electrons, Newton + multigrid-preconditioned CG
Poisson solve (ch3/ver2 SolverType::PCG, tol 1e-4; --solver pcg = the reference's Jacobi preconditioner) warm-started from the
previous step, dt=1e-7.
Injection and diagnostics are outside the timed region; the periodic cell sort is inside it.
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

QE, AMU = 1.602176565e-19, 1.660538921e-27
X0, XM = (-0.1, -0.1, 0.0), (0.1, 0.1, 0.4)
SPHERE = ((0.0, 0.0, 0.15), 0.05, -100.0)
N0, TE0, PHI0 = 1e12, 1.5, 0.0
DT = 1e-7
PUSH_BYTES = 104          # algorithmic bytes per particle of the fused push+deposit kernel


T_START = time.time()

# stdout carries exactly ONE line, the JSON result: libraries that print to file descriptor 1 (NCCL's "NCCL version ..." banner
# under NCCL_DEBUG=VERSION, for one) are sent to stderr, the result goes to the saved descriptor
_RESULT_FD = None


def claim_stdout():
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _RESULT_FD is None:
        os.write(1, line)
    else:
        os.write(_RESULT_FD, line)


def log(msg):
    if os.environ.get("BENCH_VERBOSE", "1") != "0":
        print("[bench %7.1fs] %s" % (time.time() - T_START, msg), file=sys.stderr, flush=True)


def load_espic():
    import importlib.util
    path = os.path.join(ROOT, "plasma-simulations-by-example_b200", "espic.py")
    spec = importlib.util.spec_from_file_location("espic", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["espic"] = mod
    spec.loader.exec_module(mod)
    mod.load()          # raises if the CUDA extension is missing: no fallback
    return mod


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md).  In-process NVML queries (pynvml): a
    polling `nvidia-smi -lms` child was measured to stall this process's cudaStreamSynchronize calls for milliseconds."""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = False
        self.source = None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(hi).bus == bus:
                        h = hi
                        break
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
            while not self.stop_flag:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for n, bit in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(n)
                except Exception:
                    pass
                time.sleep(self.period)
        except Exception:
            self._run_smi()

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self.source = "nvidia-smi"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(1.0)

    def stop(self):
        self.stop_flag = True

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0, "source": self.source}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": self.source}


def make_particles_device(torch, n, seed, mpw, dev, zrange=None):
    """Uniform in the box (or in the z range of one slab) outside the sphere, drift + thermal velocity; generated on the device."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    t = torch.empty((7, n), dtype=torch.float64, device=dev)
    lo, hi = list(X0), list(XM)
    if zrange is not None:
        lo[2], hi[2] = zrange
    for c in range(3):
        t[c].uniform_(0.0, 1.0, generator=g)
        t[c].mul_(hi[c] - lo[c]).add_(lo[c])
    (cx, cy, cz), r, _ = SPHERE
    d2 = (t[0] - cx) ** 2 + (t[1] - cy) ** 2 + (t[2] - cz) ** 2
    if zrange is None:
        t[2][d2 <= (1.001 * r) ** 2] += 0.2          # out of the sphere, still inside the box
    else:
        for _ in range(64):                          # slab-local: redraw the points that fell into the sphere (uniform outside it)
            bad = torch.nonzero(d2 <= (1.001 * r) ** 2).flatten()
            if bad.numel() == 0:
                break
            for c in range(3):
                t[c][bad] = torch.rand(bad.numel(), generator=g, dtype=torch.float64, device=dev) * (hi[c] - lo[c]) + lo[c]
            d2[bad] = (t[0][bad] - cx) ** 2 + (t[1][bad] - cy) ** 2 + (t[2][bad] - cz) ** 2
    del d2
    for c in range(3, 6):
        t[c].normal_(0.0, 300.0, generator=g)
    t[5].add_(7000.0)
    t[6].fill_(mpw)
    return t


def free_volume_per_cell_plane(n_mesh):
    """volume of the box outside the sphere in each cell plane k (the particle load of the synthetic case is uniform there)"""
    dz = (XM[2] - X0[2]) / (n_mesh - 1)
    z = X0[2] + dz * np.arange(n_mesh)
    (cx, cy, cz), r, _ = SPHERE
    a = np.clip(z - cz, -r, r)
    cap = np.pi * (r * r * a - a ** 3 / 3.0)          # integral of pi (r^2 - s^2) ds
    return (XM[0] - X0[0]) * (XM[1] - X0[1]) * dz - np.diff(cap)


def host_particles(rng, n, mpw):
    p = np.empty((7, n))
    for c in range(3):
        p[c] = X0[c] + rng.random(n) * (XM[c] - X0[c])
    (cx, cy, cz), r, _ = SPHERE
    d2 = (p[0] - cx) ** 2 + (p[1] - cy) ** 2 + (p[2] - cz) ** 2
    p[2, d2 <= (1.001 * r) ** 2] += 0.2
    p[3:6] = rng.normal(0, 300.0, (3, n))
    p[5] += 7000.0
    p[6] = mpw
    return p


def run_ref(which, state, cmds, tmpdir, timeout):
    """Run an oracle/_ref harness binary (compiled from the unmodified reference) with per-command timing."""
    import statefile as sf
    fin = os.path.join(tmpdir, "in.state")
    sf.write_state(fin, state)
    env = dict(os.environ, ESPIC_REF_TIMING="1", ESPIC_REF_NODUMP="1")
    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", which), fin, os.path.join(tmpdir, "out.state")] + cmds,
                         check=True, capture_output=True, text=True, env=env, timeout=timeout)
    return [(line.split()[1], float(line.split()[2])) for line in out.stdout.splitlines() if line.startswith("T ")]


def reference_cpu_step(args, e, es, sp, n_total, mpw, steps, warmup, budget_s):
    """Time the reference's own CPU implementation of one full step of THIS workload on a bounded sample.

    particle phases  Species::advance + computeNumberDensity on `cpu_sample` particles (uniform sample of the same
                     distribution, E from the warm GPU state), `warmup`+`steps` repetitions, scaled linearly to the full
                     population; serial ch3/ver2 build and, if built, the ch9/MT std::thread build on all host cores.
    mesh phases      computeChargeDensity + computeEF as measured; the Poisson solve is the reference's solveGS
                     (the solver ch3/ver2/Main.cpp ships with; its Newton-PCG breaks down at n0=1e12)
                     run ONCE to its own tolerance from the previous step's phi on the next step's rho -- exactly the
                     warm-started solve a full-size reference step performs.
    """
    import statefile as sf
    t_begin = time.time()
    n_sample = int(args.cpu_sample)
    nn = args.mesh ** 3
    # fields of step n, then rho of step n+1 from the GPU engine (the reference's solve input at full statistics)
    st = sample_state(args, e, es, n_sample, mpw, n_total)
    e.push(sp, DT, es.WALL_ABSORB, es.PUSH_FUSE_DEPOSIT)
    e.deposit(sp, es.DEPOSIT_FP64)
    e.compute_charge_density()
    rho_next = e.field(es.RHO)
    res = {"cores": 1, "kind": "reference", "unit": "particle-pushes/s"}
    with tempfile.TemporaryDirectory() as tmp:
        t = run_ref("ref_ch3", st, ["advance", "deposit", "rho", "ef"] * (warmup + steps), tmp, timeout=budget_s)
        tt = np.array([x[1] for x in t]).reshape(warmup + steps, 4)[warmup:]
        t_adv, t_dep, t_rho, t_ef = tt.mean(axis=0)
        res["ns_per_particle_serial"] = {"advance": t_adv / n_sample * 1e9, "deposit": t_dep / n_sample * 1e9}
        cores = 1
        if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_mt")):
            try:
                nc = os.cpu_count() or 1
                t = run_ref("ref_mt", st, ["threads:%d" % nc] + ["advance", "deposit"] * (warmup + steps), tmp, timeout=budget_s)
                tm = np.array([x[1] for x in t if x[0] in ("advance", "deposit")]).reshape(warmup + steps, 2)[warmup:].mean(axis=0)
                res["ns_per_particle_threads"] = {"advance": tm[0] / n_sample * 1e9, "deposit": tm[1] / n_sample * 1e9, "threads": nc}
                if tm.sum() < t_adv + t_dep:
                    t_adv, t_dep, cores = tm[0], tm[1], nc
            except Exception as ex:
                res["ns_per_particle_threads"] = {"error": repr(ex)}
        # one warm-started full solve
        st2 = sample_state(args, e, es, 0, mpw, n_total)
        st2.phi, st2.rho = st.phi, rho_next
        left = max(30.0, budget_s - (time.time() - t_begin))
        solver_note = "solveGS(20000,1e-4) warm-started, run once"
        try:
            t = run_ref("ref_ch3", st2, ["solve_gs:20000:1e-4"], tmp, timeout=left)
            t_solve = t[0][1]
        except subprocess.TimeoutExpired:
            t_solve = left
            solver_note = "solveGS did not reach its tolerance within the %.0f s budget: lower bound used" % left
    scale = n_total / n_sample
    t_step = (t_adv + t_dep) * scale + t_rho + t_ef + t_solve
    res.update({"value": n_total / t_step, "cores": cores, "s_per_step_full_size": t_step,
                "phases_s": {"advance(sample)": t_adv, "deposit(sample)": t_dep, "rho": t_rho, "solve": t_solve, "ef": t_ef},
                "sample": "unmodified reference sources (oracle/_ref, g++ -O2): %d^3 mesh; advance+deposit on %d of %d particles "
                          "(x%.0f, %d core%s), rho+ef measured, Poisson = %s" % (args.mesh, n_sample, n_total, scale, cores,
                                                                                "s" if cores > 1 else "", solver_note)})
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=float, default=2e8, help="macroparticles per GPU")
    ap.add_argument("--mesh", type=int, default=128)
    ap.add_argument("--solver", default="mg", choices=["pcg", "mg", "mgslab", "gs", "qn"],
                    help="pcg: Newton + Jacobi-PCG (the reference's preconditioner); mg: same with a multigrid V-cycle "
                         "preconditioner, solved by every rank; mgslab: the multigrid solve decomposed into one k-slab per rank")
    ap.add_argument("--sort-every", type=int, default=8)
    ap.add_argument("--fixed-point", action="store_true", help="bit-reproducible int64 deposition")
    ap.add_argument("--decomp", action="store_true",
                    help="N>1: spatial decomposition into k-slabs with particle migration (espic_migrate, ch9/MPI's scheme) "
                         "instead of sharding the particles by index")
    ap.add_argument("--fuse", action="store_true", help="scatter inside the push kernel instead of the tiled deposit kernel")
    ap.add_argument("--cpu-sample", type=float, default=2e6, help="particles of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the extra QN-solver measurement")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi clocks during the timed region")
    ap.add_argument("--dump-warm", default=None, metavar="PATH",
                    help="after the timed region write phi of the last step and rho of the next one to PATH (.npz): the warm-started "
                         "Poisson problem of one full-size step, for the offline reference-solver convergence run (scripts/ref_gs_convergence.py)")
    ap.add_argument("--profile-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed region (ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference" and rank != 0:
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1 and args.impl == "ours":
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    es = load_espic()
    n_mesh = args.mesh
    n_local = int(args.particles)
    n_total = n_local * (world if args.impl == "ours" else args.gpus)
    box_vol = (XM[0] - X0[0]) * (XM[1] - X0[1]) * (XM[2] - X0[2])
    mpw = N0 * box_vol / n_total
    solver = {"pcg": es.SOLVE_PCG, "mg": es.SOLVE_PCG_MG, "mgslab": es.SOLVE_PCG_MG_SLAB, "gs": es.SOLVE_GS, "qn": es.SOLVE_QN}[args.solver]
    if args.solver == "mgslab" and world == 1:
        solver = es.SOLVE_PCG_MG
    max_it, tol = 5000, 1e-4
    workload = "sphere-%d^3-mesh-%.0e-ions-per-gpu-%s" % (n_mesh, n_local, args.solver)

    # ---------------------------------------------------------------- engine + warm state (untimed)
    e = es.Engine(n_mesh, n_mesh, n_mesh, X0, XM, device=local_rank)
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    e.add_sphere(*SPHERE)
    e.add_inlet()
    e.set_reference_values(PHI0, TE0, N0)
    sp = e.add_species(16 * AMU, QE, mpw, capacity=int(n_local * 1.02) + 1024)
    if world > 1 and args.impl == "ours":
        uid = [e.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        e.comm_init(rank, world, uid[0])
    n_gen = n_local if args.impl == "ours" else min(n_local, int(2e7))   # reference arm only needs a warm field
    mpw_gen = mpw if args.impl == "ours" else N0 * box_vol / n_gen
    decomp = args.decomp and world > 1 and args.impl == "ours"
    if decomp and args.fuse:
        raise SystemExit("--decomp needs the separate deposit kernel (the fused scatter would run before the migration)")
    zrange = None
    if decomp:
        dhz = (XM[2] - X0[2]) / (n_mesh - 1)
        kb = es.balanced_bounds(free_volume_per_cell_plane(n_mesh), world)     # equal particle counts, not equal node counts
        zrange = (X0[2] + kb[rank] * dhz, X0[2] + kb[rank + 1] * dhz)
        e.set_domain(world, rank, kb)
        workload += "-kslab-migration"
    t = make_particles_device(torch, n_gen, 12345 + rank, mpw_gen, dev, zrange)
    e.upload_device(sp, [t[c].data_ptr() for c in range(7)], n_gen, mpw_gen)
    e.sync()
    del t
    if decomp:
        log("initial migration: sent %d, received %d" % e.migrate(sp))     # particles moved out of the sphere change slab
    torch.cuda.empty_cache()
    dmode = es.DEPOSIT_FIXED if args.fixed_point else es.DEPOSIT_FP64
    pflags = (es.PUSH_FUSE_DEPOSIT | (es.PUSH_FIXED_POINT if args.fixed_point else 0)) if args.fuse else 0

    log("particles resident: %d on rank %d" % (n_gen, rank))
    e.sort_by_cell(sp)
    e.deposit(sp, dmode)
    e.compute_charge_density()
    e.sync()
    log("sorted + deposited")
    e.solve(es.SOLVE_QN, 1, 1.0)                     # the reference's own initial guess (ctor -> solveQN)
    info0 = e.solve(es.SOLVE_GS, 20000, 1e-2)        # robust nonlinear SOR to get near the solution
    log("initial SOR: %s" % (info0,))
    if args.solver != "qn":
        info0 = e.solve(solver, max_it, tol)
        log("initial %s: %s" % (args.solver, info0))
    e.compute_ef()

    def pic_step(i, count):
        e.push(sp, DT, es.WALL_ABSORB, pflags | (es.PUSH_MIGRATE if decomp else 0))
        if decomp:
            e.migrate(sp)
        n_live = e.count(sp)
        if args.sort_every > 0 and i % args.sort_every == 0 and not args.fuse:
            e.sort_by_cell(sp)       # between push and deposit: the scatter sees perfectly ordered particles
        e.deposit(sp, dmode)
        e.compute_charge_density()
        inf = e.solve(solver, max_it, tol)
        e.compute_ef()
        if count is not None:
            count.append((n_live, inf))

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        return reference_arm(args, e, es, sp, workload, n_total, mpw)

    # ---------------------------------------------------------------- timed region (device resident)
    for i in range(args.warmup):
        wc = []
        pic_step(i, wc)
        log("warm-up step %d: n=%d %s" % (i, wc[0][0], wc[0][1]))
    sampler = ClockSampler(local_rank)
    if not args.no_clocks:
        sampler.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    push_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    phase_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(args.steps)]
    counts = []
    kernel_ms = []
    migrated = []
    launches0 = e.kernel_launches()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if args.profile_range:
        torch.cuda.profiler.start()
    ev0.record()
    for i in range(args.steps):
        pe = phase_ev[i]
        pe[0].record()
        n_before = e.count(sp)
        e.push(sp, DT, es.WALL_ABSORB, pflags | (es.PUSH_MIGRATE if decomp else 0))
        if decomp:
            migrated.append(e.migrate(sp)[0])    # inside the push phase of the timed region
        pe[1].record()
        kernel_ms.append(e.last_push_ms())       # CUDA events around the k_push launch itself, on the launching stream
        n_live = e.count(sp)
        if args.sort_every > 0 and (i + args.warmup) % args.sort_every == 0 and not args.fuse:
            e.sort_by_cell(sp)
        pe[2].record()
        e.deposit(sp, dmode)
        e.compute_charge_density()
        pe[3].record()
        inf = e.solve(solver, max_it, tol)
        pe[4].record()
        e.compute_ef()
        pe[5].record()
        counts.append((n_before, n_live, inf))
    ev1.record()
    torch.cuda.synchronize()
    if args.profile_range:
        torch.cuda.profiler.stop()
    if world > 1:
        dist.barrier()
    launches = e.kernel_launches() - launches0
    sampler.stop()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    pushed_local = sum(c[0] for c in counts)
    if world > 1:
        tp = torch.tensor([pushed_local], dtype=torch.float64, device=dev)
        dist.all_reduce(tp, op=dist.ReduceOp.SUM)
        pushed = float(tp.item())
    else:
        pushed = float(pushed_local)
    value = pushed / (ms * 1e-3)
    log("timed region done: %.2f ms/step" % (ms / args.steps))
    ph = np.array([[p[j].elapsed_time(p[j + 1]) for j in range(5)] for p in phase_ev])     # push, sort, dep+rho, solve, ef
    ph = ph[:, [1, 0, 2, 3, 4]]
    phase_ms = ph.mean(axis=0)

    # dominant kernel: k_push (Species::advance).  Its duration is the mean over the timed steps of the CUDA-event time of
    # the kernel launch alone (events recorded by the library on the launching stream); the push PHASE additionally holds
    # the removal bookkeeping (popcount, scan, hole filling, one D2H count)
    push_ms = float(phase_ms[1])
    k_ms = float(np.mean(kernel_ms))
    peak, peak_src = measured_peak_gbs()
    achieved = PUSH_BYTES * (pushed_local / args.steps) / (k_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "push_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_per_particle"] * (pushed_local / args.steps)
        except Exception:
            traffic = None

    if args.dump_warm and rank == 0:
        phi_n = e.field(es.PHI)
        e.push(sp, DT, es.WALL_ABSORB, pflags)
        e.deposit(sp, dmode)
        e.compute_charge_density()
        np.savez(args.dump_warm, phi=phi_n, rho_next=e.field(es.RHO), mesh=n_mesh, particles=e.count(sp))
        inf = e.solve(solver, max_it, tol)
        e.compute_ef()
        log("warm state dumped to %s (next solve: %s)" % (args.dump_warm, inf))

    # ---------------------------------------------------------------- e2e: same steps through the API with host buffers
    e2e = None
    if not args.no_e2e:
        rng = np.random.default_rng(99 + rank)
        n_inj = max(1, int(0.003 * n_local))
        batches = [host_particles(rng, n_inj, mpw) for _ in range(2)]
        pinned = [torch.from_numpy(b).pin_memory() for b in batches]
        pb = [p.numpy() for p in pinned]
        h2d = 7 * 8 * n_inj
        d2h = 0
        phi_pinned = torch.empty(n_mesh ** 3, dtype=torch.float64).pin_memory().numpy()    # where Output::fields would read phi
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pushed_e2e = 0
        for i in range(args.steps):
            e.add_particles(sp, pb[i % 2], DT)                 # host -> device: this step's injected particles
            pushed_e2e += e.count(sp)
            pic_step(i + 1, None)
            dg = e.diag(sp)                                    # device -> host: the step's diagnostics ...
            phi_host = e.field(es.PHI, out=phi_pinned)         # ... and the potential (what Output::fields reads)
            d2h = dg.nbytes + phi_host.nbytes + 8
        e1.record()
        torch.cuda.synchronize()
        ms_e = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms_e, pushed_e2e], dtype=torch.float64, device=dev)
            mx = tms.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tms, op=dist.ReduceOp.SUM)
            ms_e, pushed_e2e = float(mx[0].item()), float(tms[1].item())
        log("e2e region done")
        e2e = {"value": pushed_e2e / (ms_e * 1e-3), "unit": "particle-pushes/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e / args.steps}

    # ---------------------------------------------------------------- the same step with ch9's own default field solver (QN)
    variants = None
    if not args.no_variants and args.solver != "qn":
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nq = max(3, min(args.steps, 5))
        pushed_q = 0
        v0.record()
        for i in range(nq):
            pushed_q += e.count(sp)
            e.push(sp, DT, es.WALL_ABSORB, pflags | (es.PUSH_MIGRATE if decomp else 0))
            if decomp:
                e.migrate(sp)
            if args.sort_every > 0 and i % args.sort_every == 0 and not args.fuse:
                e.sort_by_cell(sp)
            e.deposit(sp, dmode)
            e.compute_charge_density()
            e.solve(es.SOLVE_QN, 1, 1.0)
            e.compute_ef()
        v1.record()
        torch.cuda.synchronize()
        ms_q = v0.elapsed_time(v1)
        if world > 1:
            tq = torch.tensor([ms_q, pushed_q], dtype=torch.float64, device=dev)
            mq = tq.clone()
            dist.all_reduce(mq, op=dist.ReduceOp.MAX)
            dist.all_reduce(tq, op=dist.ReduceOp.SUM)
            ms_q, pushed_q = float(mq[0].item()), float(tq[1].item())
        variants = {"qn": {"value": pushed_q / (ms_q * 1e-3), "unit": "particle-pushes/s", "ms_per_step": ms_q / nq, "steps": nq,
                           "note": "same workload with SolverType::QN, the solver ch9/Main.cpp ships with (ch9/Main.cpp:40); "
                                   "run after the timed region, not part of `value`"}}

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(args, e, es, sp, n_total, mpw)
        except Exception as ex:       # the baseline is reporting, never the product
            cpu = {"value": None, "unit": "particle-pushes/s", "cores": 1, "kind": "reference", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        lin = [c[2]["lin_iters"] for c in counts]
        out = {
            "metric": "particle-pushes/sec per full PIC step", "value": value, "unit": "particle-pushes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "mesh": [n_mesh] * 3, "particles_per_gpu": n_local, "solver": args.solver,
                       "solver_tol": tol, "dt": DT, "sort_every": args.sort_every,
                       "deposit": "fixed-point int64" if args.fixed_point else "fp64 atomics",
                       "parallelism": "%s x%d, NCCL density all-reduce, %s" % (
                           "k-slab spatial decomposition with particle migration (%.3g particles sent per rank per step)" % np.mean(migrated)
                           if decomp else "particle-index sharding", world, "k-slab multigrid Poisson solve over peer memory" if (args.solver == "mgslab" and world > 1) else "replicated field solve"),
                       "l2": "inputs (%.1f GB of particles per GPU) are larger than L2" % (56 * n_local / 1e9),
                       "pcg_iters_per_step": float(np.mean(lin)), "newton_iters_per_step": float(np.mean([c[2]["nr_iters"] for c in counts]))},
            "phases_ms": {"sort(amortised)": float(phase_ms[0]), "push+removal": push_ms, "deposit+rho": float(phase_ms[2]),
                          "poisson": float(phase_ms[3]), "ef": float(phase_ms[4])},
            "roofline": {"bound": "hbm", "kernel": "k_push<ABSORB%s> (Species::advance: gather + leapfrog + kill flags%s)" % (
                             (",FUSE", " + fused deposit") if args.fuse else ("", "")),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "bytes_per_particle": PUSH_BYTES, "kernel_ms": k_ms,
                         "particles_per_launch": pushed_local / args.steps,
                         "note": "duration = CUDA events around the kernel launch on its stream, mean over the timed steps"},
            "cpu_baseline": cpu,
            "solver_variants": variants,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


def sample_state(args, e, es, n_sample, mpw_full, n_full):
    """State file for the CPU runs: the GPU engine's warm fields + an independent uniform particle sample whose weight is
    scaled so the deposited density matches the full population."""
    import statefile as sf
    st = sf.State()
    st.ni = st.nj = st.nk = args.mesh
    st.flags = 3
    st.x0, st.xm, st.dt = np.array(X0), np.array(XM), DT
    st.sphere_c, st.sphere_r, st.sphere_phi = np.array(SPHERE[0]), SPHERE[1], SPHERE[2]
    st.phi0, st.Te0, st.n0 = PHI0, TE0, N0
    st.phi, st.rho, st.ef = e.field(es.PHI), e.field(es.RHO), e.field(es.EF)
    st.node_vol, st.object_id = e.field(es.NODE_VOL), e.field(es.OBJECT_ID)
    rng = np.random.default_rng(4242)
    nn = args.mesh ** 3
    if n_sample > 0:
        part = host_particles(rng, n_sample, mpw_full * n_full / n_sample)
        st.species = [dict(mass=16 * AMU, charge=QE, mpw0=part[6, 0], den=np.zeros(nn), den_ave=np.zeros(nn), part=part)]
    return st


def cpu_baseline(args, e, es, sp, n_total, mpw):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_ch3")):
        return {"value": None, "unit": "particle-pushes/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref not built"}
    return reference_cpu_step(args, e, es, sp, n_total, mpw, steps=2, warmup=1, budget_s=150.0)


def reference_arm(args, e, es, sp, workload, n_total, mpw):
    """--impl reference: the reference's CPU implementation of the same step on the host cores.  The GPU engine above was
    used only to prepare the warm field state (untimed); nothing of this repo's engine is inside the timed commands."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_ch3")):
        emit({"impl": "reference", "unavailable": "oracle/_ref/ref_ch3 was not built (needs /root/reference at build time)"})
        return 0
    res = reference_cpu_step(args, e, es, sp, n_total, mpw, steps=args.steps, warmup=args.warmup, budget_s=240.0)
    e.close()
    value = res["value"]
    out = {"impl": "reference", "metric": "particle-pushes/sec per full PIC step", "value": value, "unit": "particle-pushes/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["s_per_step_full_size"] * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload, "mesh": [args.mesh] * 3, "particles_per_gpu": int(args.particles),
                      "solver": "gs (the reference's shipped solver; its PCG diverges on this case)"},
           "cpu_baseline": res,
           "e2e": {"value": value, "unit": "particle-pushes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
