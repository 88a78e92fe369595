/*
 * espic.h -- C ABI of the B200-native ES-PIC engine (libespic_cuda.so, sm_100a).
 *
 * This is the drop-in boundary for the hot path  inject -> push -> deposit -> rho -> Poisson -> E
 * of particleincell/plasma-simulations-by-example (ch2, ch3/ver2, ch9).  The reference has no FFI:
 * its boundary is the class API that Main.cpp drives.  The host C++ shim in
 * plasma-simulations-by-example_b200/host/ (World.h, Species.h, PotentialSolver.h, Source.h, Output.h,
 * Field.h -- the reference's own file and class names) forwards every hot call to the entry points
 * below; each entry point cites the reference interface it replaces.  Plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns 0 on success, a negative value on error (espic_last_error() has the text);
 *     there is no CPU fallback: without a CUDA device espic_create fails.
 *   - node arrays are flat, u = k*ni*nj + j*ni + i (reference Field::U, ch3/ver2/Field.h:161);
 *     ef is interleaved ef[3*u+c]; object_id is int32.
 *   - particles are SoA: comp[0..6] = x, y, z, vx, vy, vz, mpw (reference struct Particle, Species.h:11-19).
 *   - all arithmetic is FP64 and follows the reference's operation order (no FMA contraction), so
 *     positions, velocities, kill masks, particle order, E and rho-from-density are bit-identical to
 *     the reference; reductions (deposition sums, dot products, diagnostics) differ by summation order only.
 *   - calls are asynchronous on the context's stream unless they return a host value.
 *   - a context is bound to one device and is not thread-safe: drive it from one host thread (the reference's drivers are
 *     single-threaded); several contexts (one per GPU / process) are independent.
 */
#ifndef ESPIC_H
#define ESPIC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct espic_ctx espic_ctx;

/* ---- lifetime ------------------------------------------------------------------------------- */

/* World::World(ni,nj,nk) + World::setExtents(x0,xm)  (ch3/ver2/World.cpp:14-36): allocates phi, rho,
 * node_vol, ef, object_id on `device`, computes dh=(xm-x0)/(n-1) and the node volumes (World.cpp:58-69). */
int  espic_create(espic_ctx **ctx, int ni, int nj, int nk, const double x0[3], const double xm[3], int device);
void espic_destroy(espic_ctx *ctx);
const char *espic_last_error(void);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the context's own */
int  espic_set_stream(espic_ctx *ctx, void *cuda_stream);
int  espic_sync(espic_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long long espic_kernel_launches(espic_ctx *ctx);
/* dh[3], xc[3] as computed by setExtents */
int  espic_get_mesh(espic_ctx *ctx, double dh[3], double xc[3]);

/* ---- geometry ------------------------------------------------------------------------------- */

/* World::addSphere (World.cpp:87-105): object_id=1, phi=phi_sphere on nodes with |x-c|^2 <= r^2 */
int espic_add_sphere(espic_ctx *ctx, const double c[3], double radius, double phi_sphere);
/* World::addInlet (World.cpp:108-115): object_id=2, phi=0 on k=0 */
int espic_add_inlet(espic_ctx *ctx);

/* ---- node fields: the public World/Species members Output.cpp and Main.cpp read -------------- */

enum { ESPIC_PHI = 0, ESPIC_RHO = 1, ESPIC_EF = 2, ESPIC_NODE_VOL = 3, ESPIC_OBJECT_ID = 4,
       ESPIC_DEN = 5, ESPIC_DEN_AVE = 6,
       /* per-species velocity-moment fields of ch4 (Species.h:85-95); VEL and NV_SUM are interleaved [3*u+c] */
       ESPIC_VEL = 7, ESPIC_T = 8, ESPIC_N_SUM = 9, ESPIC_NV_SUM = 10, ESPIC_NUU_SUM = 11, ESPIC_NVV_SUM = 12, ESPIC_NWW_SUM = 13,
       ESPIC_MPC = 14 };   /* macroparticles per CELL, (ni-1)(nj-1)(nk-1) doubles in World::XtoC order (ch4/Species.h:88) */
int espic_field_download(espic_ctx *ctx, int which, int species, void *host);
/* The same download without stalling the compute stream (what Output::fields needs while the next step already runs): the field is
 * snapshotted on the device, the snapshot is copied to `host` (pinned memory) on a second stream; espic_copy_sync waits for it. */
int espic_field_download_async(espic_ctx *ctx, int which, int species, void *host);
int espic_copy_sync(espic_ctx *ctx);
int espic_field_upload(espic_ctx *ctx, int which, int species, const void *host);
/* device pointer of a field (zero-copy interop: NCCL, torch.from_blob, ...) */
int espic_field_devptr(espic_ctx *ctx, int which, int species, void **dptr);

/* ---- species -------------------------------------------------------------------------------- */

/* Species::Species(name,mass,charge,mpw0,world) (Species.h:26-29).  Returns the species id (>=0). */
int espic_species_create(espic_ctx *ctx, double mass, double charge, double mpw0, long long capacity);
int espic_species_reserve(espic_ctx *ctx, int sp, long long capacity);
/* Species::getNp (Species.h:32) */
long long espic_species_count(espic_ctx *ctx, int sp);
/* raw particle upload/download (std::vector<Particle> particles, Species.h:65); append!=0 keeps existing ones */
int espic_species_upload(espic_ctx *ctx, int sp, const double *const comp[7], long long n, int append);
long long espic_species_download(espic_ctx *ctx, int sp, double *const comp[7], long long n_max);
/* same as espic_species_upload with DEVICE source pointers (device-to-device copy on the context's stream) */
int espic_species_upload_device(espic_ctx *ctx, int sp, const double *const dcomp[7], long long n, double mpw_max, int append);
/* Species::addParticle for n particles (Species.cpp:65-81): drop positions outside [x0,xm), gather E,
 * rewind the velocity by half a step, append in input order. */
int espic_species_add(espic_ctx *ctx, int sp, const double *const comp[7], long long n, double dt, long long *n_added);
/* Start copying the candidates of a LATER espic_species_add(ctx, sp, comp, n, ...) to the device on the copy stream (comp in pinned
 * host memory, unchanged until that call): issued one step ahead, the transfer overlaps the current step's kernels. */
int espic_species_prefetch(espic_ctx *ctx, int sp, const double *const comp[7], long long n);

/* Species::advance (ch3/ver2/Species.cpp:7-48; ch2/Species.cpp:7-38). */
enum { ESPIC_WALL_ABSORB = 0,      /* ch3/ch9: kill on sphere / outside box, swap-with-last removal (same order) */
       ESPIC_WALL_REFLECT = 1 };   /* ch2: specular reflection at the six walls, nothing removed */
enum { ESPIC_PUSH_FUSE_DEPOSIT = 1,   /* also scatter the survivors: the next espic_deposit only finalises */
       ESPIC_PUSH_NO_COMPACT = 2,     /* leave dead particles in place with mpw=0 (kill-mask tests) */
       ESPIC_PUSH_MIGRATE = 4,        /* spatial decomposition: also flag the survivors that left this part and postpone the removal to
                                         the espic_migrate that must follow (one pass closes both kinds of holes) */
       ESPIC_PUSH_DIAG = 8,           /* also sum the survivors' weight, momentum and kinetic energy while they are in registers: the
                                         next espic_species_diag returns them without a pass over the particles (Species.cpp:84-108;
                                         honoured by the plain absorbing push, ignored otherwise) */
       ESPIC_PUSH_FIXED_POINT = 256 };/* with FUSE_DEPOSIT: accumulate in int64 fixed point */
int espic_push(espic_ctx *ctx, int sp, double dt, int wall_mode, int flags);
/* device time (ms) of the push kernel of the most recent espic_push alone -- CUDA events on the launching stream,
 * without the removal bookkeeping that follows it (bench.py's roofline numerator) */
int espic_last_push_ms(espic_ctx *ctx, double *ms);

/* Species::computeNumberDensity (Species.cpp:51-62): den = scatter(mpw) / node_vol. */
enum { ESPIC_DEPOSIT_FP64 = 0,     /* FP64 atomics (order-dependent rounding) */
       ESPIC_DEPOSIT_FIXED = 1 };  /* int64 fixed-point accumulation: bit-reproducible for any order / GPU count */
int espic_deposit(espic_ctx *ctx, int sp, int mode);

/* periodic maintenance, no reference counterpart: reorder particles by cell for gather/scatter locality (counting sort; the
 * particle MULTISET is unchanged, the order is not the reference's any more).  The cell index is that of ch4 World::XtoC
 * (ch4/World.h:88-98), refined by 8 slabs of the cell along z. */
enum { ESPIC_SORT_XTOC = 0,        /* key = XtoC cell index: k slowest, i fastest */
       ESPIC_SORT_DRIFT_Z = 1 };   /* same cells, k FASTEST: a population drifting along z keeps this order from step to step
                                      (only the thermal x/y motion breaks runs), so sorts can be rare and cheap */
int espic_sort_by_cell(espic_ctx *ctx, int sp);                 /* = espic_sort_particles(ctx, sp, ESPIC_SORT_XTOC) */
int espic_sort_particles(espic_ctx *ctx, int sp, int order);

/* ColdBeamSource::sample (Source.cpp:4-27) with Philox4x32-10 counters (seed, stream, step, particle). */
int espic_inject_cold_beam(espic_ctx *ctx, int sp, double v_drift, double den, double dt,
                           uint64_t seed, uint32_t stream, uint32_t step, long long *n_added);

/* WarmBeamSource::sample (ch4/Source.cpp:31-56): Maxwellian beam at temperature T (Kelvin) -- Birdsall's sum-of-three-uniforms
 * speed times an isotropic direction (Species::sampleIsotropicVel / sampleVth, ch4/Species.cpp:149-173) plus the drift along z,
 * from Philox counters (seven blocks per particle), admitted and rewound like espic_inject_cold_beam. */
int espic_inject_warm_beam(espic_ctx *ctx, int sp, double v_drift, double den, double T, double dt,
                           uint64_t seed, uint32_t stream, uint32_t step, long long *n_added);

/* ch4 Species::advance(neutrals, spherium) (ch4/Species.cpp:8-100): push with surface interactions.  Every particle runs the
 * reference's sub-step loop: leave the box -> removed; enter the sphere -> World::lineSphereIntersect (ch4/World.cpp:160-183),
 * step back to 0.999 of the way to the surface; a NEUTRAL species (charge 0) is re-emitted diffusely with the Birdsall speed of
 * the 1000 K wall (sampleReflectedVelocity, Species.cpp:93-100; World::sphereDiffuseVector, ch4/World.cpp:185-199) and keeps
 * moving for the rest of its step; an ION dies and appends (int)(mpw0/neutrals.mpw0 + R) particles to species `neutrals_sp`
 * and (int)(yield*mpw0/sput.mpw0 + R) (yield 0.1 above 5 km/s impact speed) to species `sput_sp` through addParticle, in the
 * reference's order; emitted[0], emitted[1] return those counts (their sum in emitted[0] when both targets are the same
 * species).  Removal is the reference's swap-with-last order.  R comes from Philox counters keyed (seed, stream, step).
 * Particle::dt (ch4/Species.h:15): particles added since the species' last advance move for 2*dt on their first step, as in the
 * reference (addParticle(pos,vel) stores dt = world dt, advance adds another); the engine tracks them by index instead of an
 * eighth array, so espic_sort_by_cell refuses to run between an injection and the next advance of such a species.
 * For a neutral species neutrals_sp / sput_sp are ignored. */
int espic_push_surface(espic_ctx *ctx, int sp, double dt, int neutrals_sp, int sput_sp,
                       uint64_t seed, uint32_t stream, uint32_t step, long long emitted[2]);

/* ch4 DSMC_MEX::apply(dt) (ch4/Collisions.cpp:84-182): momentum-exchange collisions among the particles of species `sp`, cell by
 * cell (World::XtoC), Bird's no-time-counter pair selection with VHS cross-sections (ch4/Collisions.h:59-83).  *sigma_cr_max is
 * the object's running (sigma*cr)_max: read, and replaced by the maximum seen when at least one collision happened (start from
 * 1e-14 like the reference).  Pairs inside a cell are processed in the reference's order (cell lists in particle order);
 * random numbers are Philox counters keyed (seed, stream, step; cell, draw). */
int espic_dsmc_mex(espic_ctx *ctx, int sp, double dt, double *sigma_cr_max, uint64_t seed, uint32_t stream, uint32_t step,
                   long long *num_cols);
/* ch4 MCC_CEX::apply(dt) (ch4/Collisions.cpp:43-82): Monte Carlo collisions of the particles of `source_sp` with the gas of
 * `target_sp` given by its mesh fields (ESPIC_DEN, ESPIC_VEL): P = 1 - exp(-n*1e-16*|v - u|*dt) per particle; a colliding
 * particle's velocity is set to zero, as in the reference. */
int espic_mcc_cex(espic_ctx *ctx, int source_sp, int target_sp, double dt, uint64_t seed, uint32_t stream, uint32_t step,
                  long long *num_cols);
/* ch4 Species::computeMPC (ch4/Species.cpp:228-235): macroparticles per cell -> field ESPIC_MPC */
int espic_compute_mpc(espic_ctx *ctx, int sp);

/* Species::getRealCount/getMomentum/getKE (Species.cpp:84-108): out = {sum mpw, px, py, pz, KE} */
int espic_species_diag(espic_ctx *ctx, int sp, double out[5]);
/* Species::updateAverages -> Field::updateAverage (Field.h:214-221) */
int espic_update_average(espic_ctx *ctx, int sp);

/* ---- velocity moments (the "next" row 8f-1: mesh-averaged velocity and temperature) ----------- */

/* Species::sampleMoments (ch4/Species.cpp:190-200): n_sum, nv_sum, nuu_sum, nvv_sum, nww_sum += trilinear scatter of
 * mpw, mpw*vel, mpw*vx*vx, mpw*vy*vy, mpw*vz*vz */
int espic_sample_moments(espic_ctx *ctx, int sp);
/* Species::computeGasProperties (ch4/Species.cpp:203-226): vel = nv_sum/n_sum, T = m/(2K) * sum of velocity variances */
int espic_compute_gas_properties(espic_ctx *ctx, int sp);
/* Species::clearSamples (ch4/Species.cpp:239-241) */
int espic_clear_samples(espic_ctx *ctx, int sp);

/* ---- fields --------------------------------------------------------------------------------- */

/* World::computeChargeDensity (World.cpp:46-54): rho = sum_s charge_s * den_s over all species */
int espic_charge_density(espic_ctx *ctx);

enum { ESPIC_SOLVE_GS = 0,      /* PotentialSolver::solveGS, nonlinear Boltzmann SOR w=1.4 (PotentialSolver.cpp:334-430) */
       ESPIC_SOLVE_PCG = 1,     /* solveNRPCG (:225-296): Newton + Jacobi-PCG on the same discrete equations, with the Dirichlet and
                                   Neumann rows eliminated exactly so the linear system is SPD and CG cannot break down */
       ESPIC_SOLVE_QN = 2,      /* solveQN (:204-222) */
       ESPIC_SOLVE_GS_BOX = 3,  /* ch2 PotentialSolver::solve, linear SOR on interior nodes (ch2/PotentialSolver.cpp:11-67) */
       ESPIC_SOLVE_PCG_REF = 4, /* solveNRPCG + solvePCGLinear + solveGSLinear fallback restated operation for operation on the
                                   reference's non-symmetric 7-band matrix (:225-331,:433-461); inherits its breakdowns */
       ESPIC_SOLVE_PCG_MG = 5,  /* ESPIC_SOLVE_PCG with the Jacobi preconditioner of solvePCGLinear (:304) replaced by one
                                   aggregation-multigrid V-cycle: same Newton iteration, same equations, same stopping tests */
       ESPIC_SOLVE_PCG_MG_SLAB = 6 };/* ESPIC_SOLVE_PCG_MG decomposed into one k-slab per rank (needs espic_comm_init; nk must be a
                                   multiple of nranks * 2^(levels-1)): every rank computes its slab, boundary planes are stored
                                   straight into the neighbours' memory (CUDA IPC over NVLink), dot products and barriers go
                                   through peer memory too; replaces ch9/MPI updateGhosts + MPI_Allreduce
                                   (ch9/MPI/src/PotentialSolver.cpp:137,189,425-479).  Every rank ends with the full phi. */

typedef struct {
    int type;
    int max_it;          /* ctor argument max_solver_it */
    double tol;          /* ctor argument tolerance */
    double phi0, Te0, n0;/* setReferenceValues (PotentialSolver.h:53-57) */
    int nr_max_it;       /* NR_MAX_IT, 20 in the reference (:228) */
    double nr_tol;       /* NR_TOL, 1e-3 in the reference (:229) */
} espic_solve_params;

typedef struct {
    int converged;       /* what the reference's solve() returns */
    int nr_iters;
    long long lin_iters; /* PCG iterations over all Newton steps */
    int gs_fallbacks;    /* PCG failures handed to the linear GS */
    long long gs_iters;  /* SOR sweeps */
    double residual;     /* last L2 / norm */
} espic_solve_info;

int espic_solve(espic_ctx *ctx, const espic_solve_params *p, espic_solve_info *info);
/* Planning query, host only (no context, no device work): the multigrid hierarchy ESPIC_SOLVE_PCG_MG / _PCG_MG_SLAB build for a
 * mesh of ni x nj x nk nodes with spacings dh.  Returns the number of levels (<= 8) or a negative error; dims[l] = nodes per
 * dimension of level l (0 beyond the last level); *first_redundant = with nranks > 1 the first level every rank solves in full
 * (levels before it are split into k-slabs), else the coarsest level; *slab_plane_unit = fine k-planes per coarsest plane: the slab
 * solver needs nk to be a multiple of nranks * slab_plane_unit.  Replaces nothing in the reference (its PCG has no hierarchy,
 * PotentialSolver.cpp:299-331); it exposes the host logic of the preconditioner that replaces the Jacobi one. */
int espic_mg_plan(int ni, int nj, int nk, const double dh[3], int nranks, long long dims[8][3], int *first_redundant,
                  int *slab_plane_unit);
/* PotentialSolver::computeEF (PotentialSolver.cpp:465-504) */
int espic_compute_ef(espic_ctx *ctx);
/* World::getPE (World.cpp:72-84) */
int espic_field_pe(espic_ctx *ctx, double *pe);

/* ---- multi-GPU: particles sharded by index, density summed over ranks ------------ */

int espic_comm_unique_id(void *id128);                                  /* ncclGetUniqueId */
int espic_comm_init(espic_ctx *ctx, int rank, int nranks, const void *id128);
/* Once a communicator exists, espic_deposit itself sums the scatter ACCUMULATOR over all ranks (FP64, or int64 when deposited
 * in fixed point: identical bits for any rank count) before dividing by the node volumes -- that replaces ch9/MPI
 * Field::updateBoundaries (ch9/MPI/include/Field.h:122-179).  There is deliberately no separate "all-reduce the density" entry:
 * called after espic_deposit it would count every rank's particles nranks times. */

/* ---- multi-GPU: spatial decomposition with particle migration ------------------- */

/* World::initMPIDomain (ch9/MPI/include/World.h:73-128): this context simulates part `part` of `parts` slabs cut along k.
 * k_bounds[parts+1] are CELL planes, k_bounds[0] = 0 < ... < k_bounds[parts] = nk-1; part r owns the cells
 * k_bounds[r] <= k < k_bounds[r+1].  Coordinates stay global and every node array keeps its full size (the reference shifts
 * x0 per process), so gather, scatter and the push are the single-domain arithmetic; a part's scatter only touches its own node
 * planes, and the plane two parts share receives both contributions through the density all-reduce of espic_deposit, as
 * ch9/MPI Field::updateBoundaries (ch9/MPI/include/Field.h:122-179) adds them. */
int espic_domain_set(espic_ctx *ctx, int parts, int part, const int *k_bounds);
int espic_domain_get(espic_ctx *ctx, int *parts, int *part, int *k_bounds /* parts+1 ints, may be NULL */);

/* Species::move's tail (ch9/MPI/src/Species.cpp:189-213): transferParticles + the removal sweep.  Every live particle whose
 * cell belongs to another part is sent there (any part, not only a face neighbour; nothing is discarded -- the reference drops
 * particles that overshoot the neighbour, :285-288); arrivals are appended by ascending source part in the sender's particle
 * order; then ONE swap-with-last sweep (the order of espic_push) closes the holes of the dead and of the leavers over old
 * particles + arrivals, as the reference does.  Call it right after espic_push(..., ESPIC_PUSH_MIGRATE), which leaves the kill
 * and leave bits and removes nothing; without such a push (initial placement, after an injection) the leave bits are computed
 * here.  Counts travel as one all-gathered matrix, the particles in one NCCL send/recv group that writes straight behind the
 * receiver's particles.  Needs espic_comm_init with nranks == parts and rank == part.  Collective: every part must call it. */
int espic_migrate(espic_ctx *ctx, int sp, long long *n_sent, long long *n_received);

/* The pieces of espic_migrate for callers that move the segments themselves (several parts on one device, other transports):
 * pack lists and packs the leavers (nothing is removed yet) and returns counts[parts]; segment returns the device address of
 * the packed particles bound for `dest` as SoA [7][count] doubles (valid until the next pack); the receiver appends them with
 * espic_species_upload_device(..., append=1) in ascending source order; finish runs the removal sweep. */
int espic_migrate_pack(espic_ctx *ctx, int sp, long long *counts);
int espic_migrate_segment(espic_ctx *ctx, int dest, void **dptr, long long *count);
int espic_migrate_finish(espic_ctx *ctx, int sp);

#ifdef __cplusplus
}
#endif
#endif
