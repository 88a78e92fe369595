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


def run_ref(which, state_path, cmds, timeout):
    """Run an oracle/_ref harness binary (compiled from the unmodified reference) with per-command timing."""
    env = dict(os.environ, ESPIC_REF_TIMING="1", ESPIC_REF_NODUMP="1")
    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", which), state_path, "/dev/null"] + cmds,
                         check=True, capture_output=True, text=True, env=env, timeout=timeout)
    return [(line.split()[1], float(line.split()[2])) for line in out.stdout.splitlines() if line.startswith("T ")]


GS_RECORD = os.path.join(ROOT, "profiles", "r2_reference_gs_convergence.json")


def reference_cpu_step(mesh, n_total, steps, warmup, n_sample, gs_sweeps, budget_s=1500.0):
    """Time the reference's own CPU implementation (oracle/_ref, the unmodified ch3/ver2 and ch9/MT sources, g++ -O2) of one
    full step of THIS workload on the host cores.  Pure CPU: no CUDA, none of this repo's engine -- the fields the particles
    move in are produced by the reference itself (deposit -> rho -> solveQN -> computeEF on the sample).

    particle phases  Species::advance + computeNumberDensity on `n_sample` particles of the workload's distribution (weight
                     scaled so the density is the workload's), `warmup`+`steps` repetitions, mean of the timed ones, scaled
                     linearly to the full population (`sample_factor`; both loops are O(N) with no N-dependent state).
                     Serial ch3/ver2 build and the ch9/MT std::thread build on all host cores; the faster one counts.
    mesh phases      computeChargeDensity, computeEF, solveQN: measured at full size, no scaling.
    Poisson          the solver ch3/ver2/Main.cpp ships with, solveGS (its Newton-PCG breaks down at n0 = 1e12):
                     `gs_sweeps` sweeps are timed live (max_it = gs_sweeps, tolerance 0 so it cannot stop early); the number
                     of sweeps one warm-started solve needs to reach the shipped tolerance 1e-4 is read from the committed
                     record of a run of the same binary to convergence on this workload's own warm state
                     (profiles/r2_reference_gs_convergence.json, scripts/ref_gs_convergence.py).
    """
    import statefile as sf
    t_begin = time.time()
    nn = mesh ** 3
    box_vol = (XM[0] - X0[0]) * (XM[1] - X0[1]) * (XM[2] - X0[2])
    st = sf.State()
    st.ni = st.nj = st.nk = mesh
    st.flags = 3 | 4                      # addSphere + addInlet, keep the potential they set
    st.x0, st.xm, st.dt = np.array(X0), np.array(XM), DT
    st.sphere_c, st.sphere_r, st.sphere_phi = np.array(SPHERE[0]), SPHERE[1], SPHERE[2]
    st.phi0, st.Te0, st.n0 = PHI0, TE0, N0
    part = host_particles(np.random.default_rng(4242), n_sample, N0 * box_vol / n_sample)
    st.species = [dict(mass=16 * AMU, charge=QE, mpw0=part[6, 0], den=np.zeros(nn), den_ave=np.zeros(nn), part=part)]
    res = {"cores": 1, "kind": "reference", "unit": "particle-pushes/s", "extrapolated": True,
           "sample_factor": n_total / n_sample, "sample_particles": n_sample}
    prep = ["deposit", "rho", "solve_qn", "ef"]
    reps = warmup + steps
    with tempfile.TemporaryDirectory() as tmp:
        fin = os.path.join(tmp, "in.state")
        sf.write_state(fin, st)
        del part, st
        left = lambda: max(30.0, budget_s - (time.time() - t_begin))
        t = run_ref("ref_ch3", fin, prep + ["advance", "deposit", "rho", "ef", "solve_qn"] * reps + ["solve_gs:%d:0" % gs_sweeps], left())
        body = t[len(prep):]
        tt = np.array([x[1] for x in body[:5 * reps]]).reshape(reps, 5)[warmup:]
        t_adv, t_dep, t_rho, t_ef, t_qn = tt.mean(axis=0)
        t_sweep = body[5 * reps][1] / gs_sweeps
        res["ns_per_particle_serial"] = {"advance": t_adv / n_sample * 1e9, "deposit": t_dep / n_sample * 1e9}
        res["spread"] = {"advance": float(tt[:, 0].std() / t_adv), "deposit": float(tt[:, 1].std() / t_dep)}
        cores = 1
        if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_mt")):
            try:
                nc = os.cpu_count() or 1
                t = run_ref("ref_mt", fin, prep + ["threads:%d" % nc] + ["advance", "deposit"] * reps, left())
                tm = np.array([x[1] for x in t if x[0] in ("advance", "deposit")][1:]).reshape(reps, 2)[warmup:].mean(axis=0)
                res["ns_per_particle_threads"] = {"advance": tm[0] / n_sample * 1e9, "deposit": tm[1] / n_sample * 1e9, "threads": nc}
                if tm.sum() < t_adv + t_dep:
                    t_adv, t_dep, cores = tm[0], tm[1], nc
            except Exception as ex:
                res["ns_per_particle_threads"] = {"error": repr(ex)}
    scale = n_total / n_sample
    t_particles = (t_adv + t_dep) * scale
    res["cores"] = cores
    res["phases_s"] = {"advance(sample)": t_adv, "deposit(sample)": t_dep, "rho": t_rho, "ef": t_ef, "solveQN": t_qn,
                       "solveGS per sweep": t_sweep, "solveGS sweeps timed": gs_sweeps}
    # the same step with the solver ch9/Main.cpp ships (SolverType::QN): nothing but the particle count is extrapolated
    t_step_qn = t_particles + t_rho + t_qn + t_ef
    res["qn_step"] = {"value": n_total / t_step_qn, "s_per_step_full_size": t_step_qn}
    rec = None
    if os.path.exists(GS_RECORD):
        try:
            rec = json.load(open(GS_RECORD))
            if int(rec["mesh"]) != mesh:
                rec = None
        except Exception:
            rec = None
    if rec is None:
        res.update({"value": None, "s_per_step_full_size": None,
                    "reason": "no committed convergence record of the reference's solveGS for a %d^3 mesh: the Poisson phase of its "
                              "shipped solver is not extrapolated from a guess; see qn_step for the fully measured variant" % mesh})
    else:
        sweeps = float(rec["sweeps_to_converge"])
        t_step = t_particles + t_rho + t_ef + t_sweep * sweeps
        res.update({"value": n_total / t_step, "s_per_step_full_size": t_step,
                    "poisson": {"solver": "solveGS(20000, 1e-4), warm-started", "per_sweep_s": t_sweep, "sweeps_to_converge": sweeps,
                                "solve_s": t_sweep * sweeps, "record": os.path.relpath(GS_RECORD, ROOT),
                                "record_note": rec.get("note", "")}})
    res["sample"] = ("unmodified reference sources (oracle/_ref, g++ -O2), %d^3 mesh: advance+deposit measured on %d of %d particles "
                     "(x%.0f, %d core%s); rho, computeEF, solveQN measured at full size; Poisson = solveGS per-sweep time measured over %d "
                     "sweeps x %s sweeps to its tolerance" % (mesh, n_sample, n_total, scale, cores, "s" if cores > 1 else "", gs_sweeps,
                                                               "%.0f recorded" % rec["sweeps_to_converge"] if rec else "(no record)"))
    res["wall_s"] = time.time() - t_begin
    return res


class Case:
    """One engine + one ion species for a configuration of the sphere case, warm-started (untimed), with the timed PIC step."""

    def __init__(self, ctx, mesh, n_local, n_total, solver_name, sort_every=8, fixed_point=False, fuse=False, decomp=False, seed=12345,
                 sort_order="xtoc"):
        es, torch, dist = ctx["es"], ctx["torch"], ctx["dist"]
        rank, world, local_rank, dev = ctx["rank"], ctx["world"], ctx["local_rank"], ctx["dev"]
        self.ctx, self.es, self.mesh, self.n_local, self.n_total = ctx, es, mesh, n_local, n_total
        self.solver_name, self.sort_every, self.fuse = solver_name, sort_every, fuse
        self.sort_order = es.SORT_DRIFT_Z if sort_order == "drift" else es.SORT_XTOC
        box_vol = (XM[0] - X0[0]) * (XM[1] - X0[1]) * (XM[2] - X0[2])
        self.mpw = mpw = N0 * box_vol / n_total
        solver = {"pcg": es.SOLVE_PCG, "mg": es.SOLVE_PCG_MG, "mgslab": es.SOLVE_PCG_MG_SLAB, "gs": es.SOLVE_GS, "qn": es.SOLVE_QN}[solver_name]
        if solver_name == "mgslab" and world == 1:
            solver = es.SOLVE_PCG_MG
        self.solver, self.max_it, self.tol = solver, 5000, 1e-4
        e = es.Engine(mesh, mesh, mesh, X0, XM, device=local_rank)
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        e.add_sphere(*SPHERE)
        e.add_inlet()
        e.set_reference_values(PHI0, TE0, N0)
        self.e = e
        self.sp = sp = e.add_species(16 * AMU, QE, mpw, capacity=int(n_local * 1.02) + 1024)
        if world > 1:
            uid = [e.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            e.comm_init(rank, world, uid[0])
        self.decomp = decomp and world > 1
        if self.decomp and fuse:
            raise SystemExit("--decomp needs the separate deposit kernel (the fused scatter would run before the migration)")
        zrange = None
        if self.decomp:
            dhz = (XM[2] - X0[2]) / (mesh - 1)
            kb = es.balanced_bounds(free_volume_per_cell_plane(mesh), world)     # equal particle counts, not equal node counts
            zrange = (X0[2] + kb[rank] * dhz, X0[2] + kb[rank + 1] * dhz)
            e.set_domain(world, rank, kb)
        # generated in pieces of at most 5e7 so that the temporaries stay small next to a 1e9-particle population
        done = 0
        while done < n_local:
            m = min(n_local - done, 50_000_000)
            t = make_particles_device(torch, m, seed + 1000 * rank + done // 50_000_000, mpw, dev, zrange)
            e.upload_device(sp, [t[c].data_ptr() for c in range(7)], m, mpw, append=done > 0)
            e.sync()
            del t
            done += m
        if self.decomp:
            log("initial migration: sent %d, received %d" % e.migrate(sp))     # particles moved out of the sphere change slab
        torch.cuda.empty_cache()
        self.dmode = es.DEPOSIT_FIXED if fixed_point else es.DEPOSIT_FP64
        self.pflags = ((es.PUSH_FUSE_DEPOSIT | (es.PUSH_FIXED_POINT if fixed_point else 0)) if fuse else 0) | (es.PUSH_MIGRATE if self.decomp else 0)
        log("particles resident: %d on rank %d (%d^3 mesh)" % (n_local, rank, mesh))
        e.sort_particles(sp, self.sort_order)
        e.deposit(sp, self.dmode)
        e.compute_charge_density()
        e.solve(es.SOLVE_QN, 1, 1.0)                     # the reference's own initial guess (ctor -> solveQN)
        info0 = e.solve(es.SOLVE_GS, 20000, 1e-2)        # robust nonlinear SOR to get near the solution
        log("initial SOR: %s" % (info0,))
        if solver_name != "qn":
            info0 = e.solve(solver, self.max_it, self.tol)
            log("initial %s: %s" % (solver_name, info0))
        e.compute_ef()
        self.step_no = 0

    def step(self, solver=None, rec=None, ev=None):
        """push (+ migration) | sort when due | deposit + rho | Poisson | E; ev: six CUDA events recorded at the phase boundaries"""
        e, es, sp = self.e, self.es, self.sp
        i = self.step_no
        self.step_no += 1
        if ev:
            ev[0].record()
        n_before = e.count(sp)
        e.push(sp, DT, es.WALL_ABSORB, self.pflags)
        sent = e.migrate(sp)[0] if self.decomp else 0     # inside the push phase
        if ev:
            ev[1].record()
        k_ms = e.last_push_ms() if rec is not None else 0.0    # CUDA events around the k_push launch itself, on the launching stream
        if self.sort_every > 0 and i % self.sort_every == 0:
            e.sort_particles(sp, self.sort_order)       # between push and deposit: the scatter sees perfectly ordered particles
        if ev:
            ev[2].record()
        e.deposit(sp, self.dmode)
        e.compute_charge_density()
        if ev:
            ev[3].record()
        inf = e.solve(self.solver if solver is None else solver, self.max_it, self.tol)
        if ev:
            ev[4].record()
        e.compute_ef()
        if ev:
            ev[5].record()
        if rec is not None:
            rec.append((n_before, inf, k_ms, sent))

    def timed(self, steps, warmup, solver=None, profile_range=False):
        """`warmup` untimed steps, then exactly `steps` timed ones between barrier + synchronize; device time, max over ranks"""
        torch, dist, world, dev = self.ctx["torch"], self.ctx["dist"], self.ctx["world"], self.ctx["dev"]
        for i in range(warmup):
            wc = []
            self.step(solver, wc)
            log("warm-up step %d: n=%d %s" % (i, wc[0][0], wc[0][1]))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phase_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(steps)]
        rec = []
        launches0 = self.e.kernel_launches()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if profile_range:
            torch.cuda.profiler.start()
        ev0.record()
        for i in range(steps):
            self.step(solver, rec, phase_ev[i])
        ev1.record()
        torch.cuda.synchronize()
        if profile_range:
            torch.cuda.profiler.stop()
        if world > 1:
            dist.barrier()
        launches = self.e.kernel_launches() - launches0
        ms = ev0.elapsed_time(ev1)
        pushed_local = float(sum(r[0] for r in rec))
        pushed = pushed_local
        if world > 1:
            tms = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
            tp = torch.tensor([pushed_local], dtype=torch.float64, device=dev)
            dist.all_reduce(tp, op=dist.ReduceOp.SUM)
            pushed = float(tp.item())
        ph = np.array([[p[j].elapsed_time(p[j + 1]) for j in range(5)] for p in phase_ev]).mean(axis=0)   # push, sort, dep+rho, solve, ef
        return {"ms": ms, "ms_per_step": ms / steps, "pushed": pushed, "pushed_local": pushed_local, "value": pushed / (ms * 1e-3),
                "phases_ms": {"sort(amortised)": float(ph[1]), "push+removal": float(ph[0]), "deposit+rho": float(ph[2]),
                              "poisson": float(ph[3]), "ef": float(ph[4])},
                "kernel_ms": float(np.mean([r[2] for r in rec])), "launches": int(launches),
                "pcg_iters_per_step": float(np.mean([r[1]["lin_iters"] for r in rec])),
                "newton_iters_per_step": float(np.mean([r[1]["nr_iters"] for r in rec])),
                "migrated_per_step": float(np.mean([r[3] for r in rec])), "steps": steps, "warmup": warmup}

    def close(self):
        self.e.close()
        self.ctx["torch"].cuda.empty_cache()


# algorithmic bytes of the multigrid Newton solve (espic_mg.cuh), per node
MG_FINE_BYTES_PER_IT = 93      # A: rf 4 + winv 4 | B: rf 4 + winv 4 + e 1, z 4 written | C: z 4 + d 8 + diag 4, d' 8 written | D: d' 8 + diag 4 + delta 8+8 + r 8+8, rf 4 written
MG_COARSE_BYTES_PER_IT = 61    # FP32 level: down 24 read + 4.5 written, up 28.5 read + 4 written
MG_NEWTON_BYTES = 83           # linearise 17 read + 36 written, coarse diagonals 5, update 17 read + 8 written
MG_LINEARISE_BYTES = 53        # the closing residual evaluation


def poisson_roofline(case, res, peak):
    """Achieved bytes/s of the Poisson phase: algorithmic bytes of every pass of the multigrid Newton kernel x the iteration
    counts of the timed steps, over the phase time (CUDA events around espic_solve).  Per rank: a k-slab in slab mode."""
    es = case.es
    dh = [(XM[a] - X0[a]) / (case.mesh - 1) for a in range(3)]
    dims, _, _ = es.mg_plan(case.mesh, case.mesh, case.mesh, dh, 1)
    share = 1.0 / case.ctx["world"] if (case.solver_name == "mgslab" and case.ctx["world"] > 1) else 1.0
    fine = float(np.prod(dims[0])) * share
    coarse = float(sum(np.prod(d) for d in dims[1:])) * share
    per_it = MG_FINE_BYTES_PER_IT * fine + MG_COARSE_BYTES_PER_IT * coarse
    byt = per_it * res["pcg_iters_per_step"] + MG_NEWTON_BYTES * fine * res["newton_iters_per_step"] + MG_LINEARISE_BYTES * fine
    t = res["phases_ms"]["poisson"] * 1e-3
    ach = byt / t / 1e9 if t > 0 else 0.0
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "bytes_per_cg_iteration": per_it, "bytes_per_step": byt, "phase_ms": res["phases_ms"]["poisson"],
            "levels": [list(d) for d in dims],
            "note": "algorithmic bytes per rank: (%d B/fine node + %d B/coarse node) x CG iterations + %d B/fine node x Newton steps + %d B/fine node, over "
                    "the Poisson phase time of the timed region" % (MG_FINE_BYTES_PER_IT, MG_COARSE_BYTES_PER_IT, MG_NEWTON_BYTES, MG_LINEARISE_BYTES)}


def digest(a):
    import hashlib
    return hashlib.sha1(np.ascontiguousarray(a).view(np.uint8)).hexdigest()


def parity_check(ctx, head):
    """N > 1, untimed, after the headline region: the assertions of tests/test_multigpu.py executed under torch.distributed.run.
      slab          ESPIC_SOLVE_PCG_MG_SLAB against the replicated ESPIC_SOLVE_PCG_MG on the headline state: same start, same rho
      deposit       fixed-point density of N index shards (all-reduced) == one rank's deposit of the gathered particles, bit for bit
      migration     k-slab decomposition with espic_migrate for 3 steps == the single-domain run: same particle multiset, bit for bit"""
    es, torch, dist, rank, world, dev = ctx["es"], ctx["torch"], ctx["dist"], ctx["rank"], ctx["world"], ctx["dev"]
    out = {}
    e, sp = head.e, head.sp
    # ---- slab vs replicated on the live 128^3 state: perturb by one more push + deposit so both solves have work to do
    try:
        phi_start = e.field(es.PHI)
        e.push(sp, DT, es.WALL_ABSORB, 0)
        e.deposit(sp, head.dmode)
        e.compute_charge_density()
        i_rep = e.solve(es.SOLVE_PCG_MG, head.max_it, head.tol)
        phi_rep = e.field(es.PHI)
        e.set_field(es.PHI, phi_start)
        i_slab = e.solve(es.SOLVE_PCG_MG_SLAB, head.max_it, head.tol)
        phi_slab = e.field(es.PHI)
        e.compute_ef()
        rel = float(np.abs(phi_slab - phi_rep).max() / np.abs(phi_rep).max())
        dg = [None] * world
        dist.all_gather_object(dg, digest(phi_slab))
        out["slab_vs_replicated"] = {"max_rel_diff_phi": rel, "newton": [i_rep["nr_iters"], i_slab["nr_iters"]],
                                     "cg_iterations": [i_rep["lin_iters"], i_slab["lin_iters"]],
                                     "phi_identical_on_all_ranks": len(set(dg)) == 1,
                                     "ok": bool(rel <= 1e-10 and len(set(dg)) == 1 and i_slab["converged"] == 1 and i_rep["converged"] == 1)}
    except Exception as ex:
        out["slab_vs_replicated"] = {"ok": False, "error": repr(ex)}
    # ---- small case: 33 x 33 x 65 mesh, 2e5 particles per rank
    try:
        ni, nj, nk, n_loc = 33, 33, 65, 200_000
        rng = np.random.default_rng(777 + rank)
        part = host_particles(rng, n_loc, 50.0)
        part[5] += rng.normal(0, 3000.0, n_loc)          # fast enough that particles cross slabs within 3 steps of 2e-6
        small_dt = 2e-6

        def engine(comm):
            g = es.Engine(ni, nj, nk, X0, XM, device=ctx["local_rank"])
            g.add_sphere(*SPHERE)
            g.add_inlet()
            s = g.add_species(16 * AMU, QE, 50.0, capacity=4 * n_loc * (1 if comm else world))
            if comm:
                uid = [g.unique_id() if rank == 0 else None]
                dist.broadcast_object_list(uid, src=0)
                g.comm_init(rank, world, uid[0])
            return g, s

        g, s = engine(True)
        g.upload(s, part)
        g.deposit(s, es.DEPOSIT_FIXED)
        den = g.field(es.DEN, s)
        dg = [None] * world
        dist.all_gather_object(dg, digest(den))
        allp = [None] * world
        dist.all_gather_object(allp, part)
        full = np.concatenate(allp, axis=1)
        one, s1 = engine(False)
        one.upload(s1, full)
        one.deposit(s1, es.DEPOSIT_FIXED)
        den1 = one.field(es.DEN, s1)
        out["fixed_point_deposit"] = {"identical_on_all_ranks": len(set(dg)) == 1, "equals_single_rank_deposit_bitwise": bool(np.array_equal(den.view(np.uint64), den1.view(np.uint64))),
                                      "ok": bool(len(set(dg)) == 1 and np.array_equal(den.view(np.uint64), den1.view(np.uint64)))}
        g.close()
        # migration: every rank owns a k-slab; particles start on their owner
        kb = es.slab_bounds(nk, world)
        dhz = (XM[2] - X0[2]) / (nk - 1)
        kcell = np.minimum(((full[2] - X0[2]) / dhz).astype(np.int64), nk - 2)
        mine = full[:, (kcell >= kb[rank]) & (kcell < kb[rank + 1])]
        g, s = engine(True)
        g.set_domain(world, rank, kb)
        g.upload(s, np.ascontiguousarray(mine))
        moved = 0
        for _ in range(3):
            g.push(s, small_dt, es.WALL_ABSORB, es.PUSH_MIGRATE)
            moved += g.migrate(s)[0]
            one.push(s1, small_dt, es.WALL_ABSORB, 0)
        parts = [None] * world
        dist.all_gather_object(parts, g.download(s))
        union = np.concatenate(parts, axis=1)
        ref = one.download(s1)

        def canon(a):
            return a[:, np.lexsort(a[::-1])]
        same = union.shape == ref.shape and bool(np.array_equal(canon(union).view(np.uint64), canon(ref).view(np.uint64)))
        mv = torch.tensor([moved], dtype=torch.float64, device=dev)
        dist.all_reduce(mv)
        out["migration"] = {"union_of_parts_equals_single_domain_bitwise": same, "particles": int(ref.shape[1]), "migrated_total": int(mv.item()),
                            "ok": bool(same and mv.item() > 0)}
        g.close()
        one.close()
    except Exception as ex:
        out.setdefault("fixed_point_deposit", {"ok": False, "error": repr(ex)})
        out.setdefault("migration", {"ok": False, "error": repr(ex)})
    out["ok"] = all(v.get("ok", False) for v in out.values() if isinstance(v, dict))
    return out


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
    ap.add_argument("--sort-order", default="xtoc", choices=["drift", "xtoc"],
                    help="key order of the periodic cell sort: drift = k fastest (a beam drifting along z keeps it), xtoc = ch4 World::XtoC order")
    ap.add_argument("--fixed-point", action="store_true", help="bit-reproducible int64 deposition")
    ap.add_argument("--decomp", action="store_true",
                    help="N>1: spatial decomposition into k-slabs with particle migration (espic_migrate, ch9/MPI's scheme) "
                         "instead of sharding the particles by index")
    ap.add_argument("--fuse", action="store_true", help="scatter inside the push kernel instead of the tiled deposit kernel")
    ap.add_argument("--cpu-sample", type=float, default=1e7, help="particles of the reference arm's bounded sample (the in-line cpu_baseline uses at most 2e6)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-mode", type=int, default=7, help="bit 0: diagnostics inside the push kernel, bit 1: particle H2D prefetched one "
                                                            "step ahead on the copy stream, bit 2: potential downloaded asynchronously")
    ap.add_argument("--no-variants", action="store_true", help="skip the extra QN-solver measurement")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra records (configs[2], strong scaling, configs[4], parity check)")
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

    if args.impl == "reference":
        return 0 if rank != 0 else reference_arm(args)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    es = load_espic()
    ctx = {"es": es, "torch": torch, "dist": dist, "rank": rank, "world": world, "local_rank": local_rank, "dev": dev}
    n_mesh = args.mesh
    n_local = int(args.particles)
    n_total = n_local * world
    workload = "sphere-%d^3-mesh-%.0e-ions-per-gpu-%s" % (n_mesh, n_local, args.solver)
    if args.decomp and world > 1:
        workload += "-kslab-migration"

    # ---------------------------------------------------------------- headline: BASELINE configs[3], weak scaling
    head = Case(ctx, n_mesh, n_local, n_total, args.solver, args.sort_every, args.fixed_point, args.fuse, args.decomp, sort_order=args.sort_order)
    e, sp = head.e, head.sp
    # warm-up first, then sample clocks over the timed steps only
    for i in range(args.warmup):
        wc = []
        head.step(None, wc)
        log("warm-up step %d: n=%d %s" % (i, wc[0][0], wc[0][1]))
    sampler = ClockSampler(local_rank)
    if not args.no_clocks:
        sampler.start()
        time.sleep(0.3)
    res = head.timed(args.steps, 0, profile_range=args.profile_range)
    sampler.stop()
    log("timed region done: %.2f ms/step" % res["ms_per_step"])
    value, ms = res["value"], res["ms"]
    pushed_local = res["pushed_local"]

    # dominant kernel: k_push (Species::advance).  Its duration is the mean over the timed steps of the CUDA-event time of
    # the kernel launch alone (events recorded by the library on the launching stream); the push PHASE additionally holds
    # the removal bookkeeping (popcount, scan, hole filling, one D2H count)
    k_ms = res["kernel_ms"]
    peak, peak_src = measured_peak_gbs()
    achieved = PUSH_BYTES * (pushed_local / args.steps) / (k_ms * 1e-3) / 1e9
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "push_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_per_particle"] * (pushed_local / args.steps)
            traffic_note = "ncu constant (profiles/push_traffic.json: %.1f DRAM bytes per particle from one `ncu --set full` capture) x particles per launch; not measured in this run" % tj["dram_bytes_per_particle"]
        except Exception:
            traffic = None

    if args.dump_warm and rank == 0:
        phi_n = e.field(es.PHI)
        e.push(sp, DT, es.WALL_ABSORB, head.pflags)
        e.deposit(sp, head.dmode)
        e.compute_charge_density()
        np.savez(args.dump_warm, phi=phi_n, rho_next=e.field(es.RHO), mesh=n_mesh, particles=e.count(sp))
        inf = e.solve(head.solver, head.max_it, head.tol)
        e.compute_ef()
        log("warm state dumped to %s (next solve: %s)" % (args.dump_warm, inf))

    # ---------------------------------------------------------------- e2e: same steps through the API with host buffers
    e2e = None
    if not args.no_e2e:
        # a second, identical case taken through the same warm-up: the e2e region then works on the population the device-timed
        # region worked on (the beam leaves the box at ~0.5 % per step; timing the e2e steps on what the headline region left
        # behind would compare 1.6e8 with 1.85e8 particles per step)
        ec = Case(ctx, n_mesh, n_local, n_total, args.solver, args.sort_every, args.fixed_point, args.fuse, args.decomp, sort_order=args.sort_order)
        for i in range(args.warmup):
            ec.step()
        e, sp = ec.e, ec.sp
        rng = np.random.default_rng(99 + rank)
        n_inj = max(1, int(0.003 * n_local))
        batches = [host_particles(rng, n_inj, ec.mpw) for _ in range(2)]
        pinned = [torch.from_numpy(b).pin_memory() for b in batches]
        pb = [p.numpy() for p in pinned]
        h2d = 7 * 8 * n_inj
        d2h = 0
        phi_host = [torch.empty(n_mesh ** 3, dtype=torch.float64).pin_memory().numpy() for _ in range(2)]
        m_diag, m_pre, m_async = bool(args.e2e_mode & 1), bool(args.e2e_mode & 2), bool(args.e2e_mode & 4)
        skip = int(os.environ.get("BENCH_E2E_SKIP", "0"))      # development aid: 1 no injection, 2 no diagnostics, 4 no potential
        if m_diag:
            ec.pflags |= es.PUSH_DIAG                          # the step's diagnostics ride in the push kernel's registers

        def e2e_steps(count, pev=None):
            pushed, d2h_b = 0, 0
            if m_pre and not skip & 1:
                e.prefetch_particles(sp, pb[0])                # the copy engine works one step ahead of the kernels
            for i in range(count):
                if not skip & 1:
                    e.add_particles(sp, pb[i % 2], DT)         # host -> device: this step's injected particles (staged by the prefetch)
                    if m_pre and i + 1 < count:
                        e.prefetch_particles(sp, pb[(i + 1) % 2])  # next step's batch travels while this step computes
                pushed += e.count(sp)
                ec.step(None, erec if pev else None, pev[i] if pev else None)
                dg = e.diag(sp) if not skip & 2 else np.zeros(5)   # device -> host: the step's diagnostics ...
                if skip & 4:
                    pass
                elif m_async:
                    if i > 0:
                        e.copy_sync()                          # (previous step's potential has arrived in its pinned buffer)
                    e.field_async(es.PHI, phi_host[i % 2])     # ... and the potential (what Output::fields reads), on the copy stream
                else:
                    e.field(es.PHI, out=phi_host[i % 2])
                d2h_b = dg.nbytes + phi_host[0].nbytes + 8
            e.copy_sync()
            return pushed, d2h_b

        if os.environ.get("BENCH_E2E_TRACE"):          # development aid: serialised cost of every call of the e2e step
            tr = {}

            def timed_call(name, fn):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r = fn()
                torch.cuda.synchronize()
                e.copy_sync()
                tr.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
                return r
            for i in range(6):
                timed_call("add_particles", lambda: e.add_particles(sp, pb[i % 2], DT))
                timed_call("push", lambda: e.push(sp, DT, es.WALL_ABSORB, ec.pflags))
                timed_call("deposit", lambda: e.deposit(sp, ec.dmode))
                timed_call("rho", lambda: e.compute_charge_density())
                timed_call("solve", lambda: e.solve(ec.solver, ec.max_it, ec.tol))
                timed_call("ef", lambda: e.compute_ef())
                timed_call("diag", lambda: e.diag(sp))
                timed_call("phi download", lambda: e.field(es.PHI, out=phi_host[0]))
            log("e2e trace (ms, serialised): %s" % {k: round(float(np.mean(v[1:])), 3) for k, v in tr.items()})
        e2e_steps(3)             # untimed: first-use allocations of the staging buffers (a cudaMalloc next to 22 GB of particles costs ~100 ms)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(args.steps)]
        erec = []
        pushed_e2e, d2h = e2e_steps(args.steps, pev)
        e1.record()
        torch.cuda.synchronize()
        ms_e = e0.elapsed_time(e1)
        eph = np.array([[p[j].elapsed_time(p[j + 1]) for j in range(5)] for p in pev]).mean(axis=0)
        e2e_phases = {"push+removal+diagnostics": float(eph[0]), "sort(amortised)": float(eph[1]), "deposit+rho": float(eph[2]),
                      "poisson": float(eph[3]), "ef": float(eph[4]),
                      "injection, downloads, host gaps": float(ms_e / args.steps - eph.sum())}
        ec.close()
        e, sp = head.e, head.sp
        if world > 1:
            tms = torch.tensor([ms_e, pushed_e2e], dtype=torch.float64, device=dev)
            mx = tms.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tms, op=dist.ReduceOp.SUM)
            ms_e, pushed_e2e = float(mx[0].item()), float(tms[1].item())
        log("e2e region done")
        e2e = {"value": pushed_e2e / (ms_e * 1e-3), "unit": "particle-pushes/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e / args.steps,
               "particles_per_step": pushed_e2e / args.steps / world, "phases_ms": e2e_phases, "k_push_ms": float(np.mean([r[2] for r in erec])),
               "note": "a second, identical case after the same warm-up steps (+3 untimed e2e steps), so that both regions push the same "
                       "population; every step: injected particles from pinned host memory, diagnostics and the potential back to the host"}

    # ---------------------------------------------------------------- the same step with ch9's own default field solver (QN)
    variants = None
    if not args.no_variants and args.solver != "qn":
        rq = head.timed(max(3, min(args.steps, 5)), 0, solver=es.SOLVE_QN)
        variants = {"qn": {"value": rq["value"], "unit": "particle-pushes/s", "ms_per_step": rq["ms_per_step"], "steps": rq["steps"],
                           "note": "same workload with SolverType::QN, the solver ch9/Main.cpp ships with (ch9/Main.cpp:40); "
                                   "run after the timed region, not part of `value`"}}

    # ---------------------------------------------------------------- N > 1: multi-GPU correctness, executed where the driver runs
    parity = None
    if world > 1 and not args.no_extra and args.solver in ("mg", "mgslab") and not args.decomp:
        parity = parity_check(ctx, head)
        log("parity check: %s" % (parity,))
    head.close()

    # ---------------------------------------------------------------- extra records (same timing contract, fewer steps)
    extra = {}
    if not args.no_extra and args.mesh == 128 and args.solver == "mg" and not args.decomp:
        ks = max(3, min(args.steps, 6))

        def record(tag, mesh, n_loc, n_tot, solver_name, note):
            try:
                free = torch.cuda.mem_get_info(dev)[0]
                need = 2 * 56 * n_loc * 1.03 + 40 * 8 * mesh ** 3 + (2 << 30)     # particles + sort double buffer, ~40 node arrays
                if need > free:
                    extra[tag] = {"skipped": "needs %.0f GB of device memory, %.0f GB free" % (need / 1e9, free / 1e9)}
                    return
                cs = Case(ctx, mesh, n_loc, n_tot, solver_name, args.sort_every, sort_order=args.sort_order)
                r = cs.timed(ks, 3)
                pk = r["kernel_ms"]
                extra[tag] = {"value": r["value"], "unit": "particle-pushes/s", "ms_per_step": r["ms_per_step"], "steps": ks, "warmup": 3,
                              "n_gpus": world, "mesh": [mesh] * 3, "particles_total": n_tot, "particles_per_gpu": n_loc, "solver": solver_name,
                              "phases_ms": r["phases_ms"], "pcg_iters_per_step": r["pcg_iters_per_step"],
                              "newton_iters_per_step": r["newton_iters_per_step"],
                              "push_roofline_frac": PUSH_BYTES * (r["pushed_local"] / ks) / (pk * 1e-3) / 1e9 / peak if pk > 0 else None,
                              "roofline_poisson": poisson_roofline(cs, r, peak), "note": note}
                cs.close()
            except Exception as ex:       # an extra record never takes the headline line down
                extra[tag] = {"error": repr(ex)}
            log("%s: %s" % (tag, extra[tag]))

        if world == 1:
            record("config2", 128, 50_000_000, 50_000_000, "mg", "BASELINE configs[2]: 128^3 mesh, 5e7 ions, one B200")
        else:
            record("strong", 128, 200_000_000 // world, 200_000_000, "mg",
                   "BASELINE configs[3] read as strong scaling: 2e8 ions in total, sharded by index over the ranks, replicated field solve")
        record("config5", 256, 1_000_000_000 // world, 1_000_000_000, "mgslab" if world > 1 else "mg",
               "BASELINE configs[4]: 256^3 mesh, 1e9 ions in total (strong scaling in N); N>1: k-slab multigrid solve over peer memory, "
               "N=1: the same solver on one GPU -- the denominator of the 8-GPU speed-up")

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(args, n_total)
        except Exception as ex:       # the baseline is reporting, never the product
            cpu = {"value": None, "unit": "particle-pushes/s", "cores": 1, "kind": "reference", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        out = {
            "metric": "particle-pushes/sec per full PIC step", "value": value, "unit": "particle-pushes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "mesh": [n_mesh] * 3, "particles_per_gpu": n_local, "solver": args.solver,
                       "solver_tol": head.tol, "dt": DT, "sort_every": args.sort_every, "sort_order": args.sort_order,
                       "deposit": "fixed-point int64" if args.fixed_point else "fp64 atomics",
                       "parallelism": "%s x%d, NCCL density all-reduce, %s" % (
                           "k-slab spatial decomposition with particle migration (%.3g particles sent per rank per step)" % res["migrated_per_step"]
                           if head.decomp else "particle-index sharding", world, "k-slab multigrid Poisson solve over peer memory" if (args.solver == "mgslab" and world > 1) else "replicated field solve"),
                       "l2": "inputs (%.1f GB of particles per GPU) are larger than L2" % (56 * n_local / 1e9),
                       "pcg_iters_per_step": res["pcg_iters_per_step"], "newton_iters_per_step": res["newton_iters_per_step"]},
            "phases_ms": res["phases_ms"],
            "roofline": {"bound": "hbm", "kernel": "k_push<ABSORB%s> (Species::advance: gather + leapfrog + kill flags%s)" % (
                             (",FUSE", " + fused deposit") if args.fuse else ("", "")),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_note": traffic_note,
                         "peak_source": peak_src, "bytes_per_particle": PUSH_BYTES, "kernel_ms": k_ms,
                         "particles_per_launch": pushed_local / args.steps,
                         "note": "duration = CUDA events around the kernel launch on its stream, mean over the timed steps"},
            "roofline_poisson": poisson_roofline(head, res, peak) if args.solver in ("mg", "mgslab") else None,
            "cpu_baseline": cpu,
            "solver_variants": variants,
            "e2e": e2e,
            "gpu_launches": res["launches"],
            "clocks": sampler.summary(),
        }
        if parity is not None:
            out["parity_check"] = parity
        out.update(extra)
        emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


def cpu_baseline(args, n_total):
    """bounded sample of the reference on the host cores beside the GPU number (about 20-30 s)"""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_ch3")):
        return {"value": None, "unit": "particle-pushes/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref not built"}
    return reference_cpu_step(args.mesh, n_total, steps=2, warmup=1, n_sample=int(min(args.cpu_sample, 2e6)), gs_sweeps=10, budget_s=300.0)


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the same step on the host cores.  No CUDA context, no library
    of this repo is loaded by this process or by the binaries it runs."""
    workload = "sphere-%d^3-mesh-%.0e-ions-per-gpu-%s" % (args.mesh, int(args.particles), args.solver)
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_ch3")):
        emit({"impl": "reference", "unavailable": "oracle/_ref/ref_ch3 was not built (needs /root/reference at build time)"})
        return 0
    n_total = int(args.particles) * args.gpus
    res = reference_cpu_step(args.mesh, n_total, steps=args.steps, warmup=args.warmup, n_sample=int(args.cpu_sample), gs_sweeps=40)
    value = res["value"]
    out = {"impl": "reference", "metric": "particle-pushes/sec per full PIC step", "value": value, "unit": "particle-pushes/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": None if value is None else res["s_per_step_full_size"] * 1e3,
           "extrapolated": True, "sample_factor": res["sample_factor"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload, "mesh": [args.mesh] * 3, "particles_per_gpu": int(args.particles),
                      "solver": "gs (the reference's shipped solver, PotentialSolver.cpp:334-430; its Newton-PCG diverges on this case) -- "
                                "the repo's arm solves the same equations to the same tolerance with Newton + multigrid-PCG",
                      "same_config": "same mesh, particles, dt, equations and tolerance; different iterative solver (see DESIGN.md 7)"},
           "cpu_baseline": res,
           "solver_variants": {"qn": dict(res["qn_step"], unit="particle-pushes/s",
                                          note="same step with SolverType::QN (ch9/Main.cpp:40); compare with the repo arm's solver_variants.qn")},
           "e2e": {"value": value, "unit": "particle-pushes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
