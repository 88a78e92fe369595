// espic_internal.cuh -- shared state and device helpers of libespic_cuda.so (sm_100a only).
// Compiled with -fmad=false: every FP64 expression keeps the reference's operation order and
// rounding, so per-particle and per-node results are bit-identical to g++ -O2 on x86-64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "espic.h"

#define ESPIC_MAX_SPECIES 8

struct MeshC {
    int ni, nj, nk;
    long long nn;
    double x0[3], xm[3], dh[3];
    double rdh[3];    // RN(1/dh): div_by_dh() turns it back into the exactly rounded quotient
    double sc[3];     // sphere centre
    double sr2;       // sphere radius^2 (0: World default, World.h:136-137)
};

struct Species {
    double mass = 0, charge = 0, mpw0 = 0;
    long long np = 0, cap = 0;
    double *p[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};      // x y z vx vy vz mpw
    double *alt[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};    // sort double buffer
    long long alt_cap = 0;
    double *den = nullptr, *den_ave = nullptr;
    double *acc = nullptr;             // FP64 scatter accumulator (also aliased as int64 in fixed-point mode)
    double *mpc = nullptr;             // macroparticles per cell (ch4 Species::mpc), allocated on first use
    double *mom = nullptr;             // velocity moments, allocated on first use: n_sum | nv_sum[3] | nuu | nvv | nww | vel[3] | T
    int ave_samples = 0;
    bool acc_fresh = false;            // accumulator holds the scatter of the current particle state
    int acc_mode = ESPIC_DEPOSIT_FP64;
    int acc_shift = 0;                 // fixed point: value * 2^shift
    double mpw_max = 0;                // upper bound of any mpw seen (fixed-point scale)
    int pushes_since_sort = 1 << 20;   // how scrambled the cell order is: chooses the deposit kernel
    int sort_order = 0;                // key order of the most recent sort (ESPIC_SORT_*)
    // ch4 Particle::dt by index (espic_surface.cuh): particles [0, n_settled) went through the last advance (dt = 0),
    // particles [n_settled, np) were added since (dt = world dt)
    long long n_settled = 0;
    bool substep = false;              // the species is advanced with espic_push_surface
    // migration in flight (espic_migrate.cuh): 0 none; 1 espic_push(ESPIC_PUSH_MIGRATE) left kill + leave bits for the first
    // mig_n particles, nothing removed yet; 2 leavers packed, arrivals may be appended, removal still to come
    int mig_stage = 0;
    long long mig_n = 0;
    // diagnostics of the particles as the last espic_push(ESPIC_PUSH_DIAG) left them: sum mpw, sum mpw v (3), sum mpw v^2
    bool diag_valid = false;
    double diag_sums[5] = {0, 0, 0, 0, 0};
    // kill bits (one per particle) written by the push kernels and consumed by the removal; leave bits of a MIGRATE push.
    // Per species: after espic_push(A, ESPIC_PUSH_MIGRATE) they must survive calls on other species until espic_migrate(A).
    uint32_t *kill_words = nullptr;  long long kill_cap = 0;
    uint32_t *leave_words = nullptr; long long leave_cap = 0;
};

struct espic_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    MeshC m;
    double xc[3];
    double *phi = nullptr, *rho = nullptr, *ef = nullptr, *node_vol = nullptr;
    double *ef4 = nullptr;             // gather copy of ef, one 32-byte {ex,ey,ez,0} record per node (one LDG.256 per node)
    int32_t *object_id = nullptr;
    int nsp = 0;
    Species sp[ESPIC_MAX_SPECIES];
    long long launches = 0;
    // CUDA events around the most recent k_push launch on `stream` (bench.py's roofline numerator uses the kernel alone)
    cudaEvent_t push_ev0 = nullptr, push_ev1 = nullptr;
    bool push_timed = false;
    // scratch
    uint32_t *dead_words = nullptr; long long dead_words_cap = 0;   // per-cell fill cursors of espic_dsmc_mex (transient scratch)
    uint32_t *hit_words = nullptr;  long long hit_words_cap = 0;    // ions that hit the sphere (espic_push_surface)
    int dom_klo = 0, dom_khi = 1 << 30;                             // cell planes this part owns (espic_domain_set)
    uint32_t *scan_pre = nullptr;  long long scan_cap = 0;
    uint32_t *scan_coff = nullptr; long long scan_coff_cap = 0;
    long long *lists = nullptr;    long long lists_cap = 0;   // holes | fillers
    double *red = nullptr;         long long red_cap = 0;     // reduction partials
    unsigned long long *dscal = nullptr;                      // small device scalars (128 x 8 B)
    void *hpin = nullptr;                                     // pinned host mirror of dscal (128 x 8 B)
    uint32_t *cell_cnt = nullptr;  long long cell_cap = 0;
    uint32_t *sort_key = nullptr;  long long sort_key_cap = 0;   // cell sort: key of every particle, then ...
    uint32_t *sort_src = nullptr;  long long sort_src_cap = 0;   // ... the particle that goes to every place of the new order
    // solver work vectors
    double *sv[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint8_t *node_type = nullptr;
    int node_type_mode = -1;       // which solver family node_type was built for
    long long geom_version = 0, node_type_version = -1, diag0_version = -1;
    int sm_count = 148;
    void *mg = nullptr;            // MgHierarchy of the multigrid-preconditioned solver (espic_mg.cuh)
    void *slab = nullptr;          // SlabState of the slab-decomposed multi-GPU variant (espic_mg.cuh)
    void *mig = nullptr;           // MigState of the spatial decomposition with particle migration (espic_migrate.cuh)
    // copy engine side: a second stream for host <-> device traffic that overlaps the compute stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_stage = nullptr, ev_snap = nullptr, ev_copied = nullptr;
    double *stage = nullptr; long long stage_cap = 0;       // particles prefetched by espic_species_prefetch ([7][stage_n])
    const void *stage_host = nullptr; long long stage_n = 0; int stage_sp = -1;
    void *snap = nullptr; size_t snap_cap = 0;               // device snapshot of a field on its way to the host
    bool copy_pending = false;
    // comm
    void *nccl = nullptr; int rank = 0, nranks = 1;
};

void espic_set_error(const char *fmt, ...);
int  espic_cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return espic_cuda_fail(e_, #call, __FILE__, __LINE__); } while (0)
#define LAUNCH_CHECK(ctx) do { (ctx)->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return espic_cuda_fail(e_, "kernel launch", __FILE__, __LINE__); } while (0)

// between espic_push(ESPIC_PUSH_MIGRATE) / espic_migrate_pack and the end of the migration the arrays still hold the dead and the leavers
#define MIG_GUARD(c, s, who) do { if ((s).mig_stage != 0) { espic_set_error("%s: a migration of this species is pending (call espic_migrate / espic_migrate_finish first)", who); return -1; } } while (0)

int espic_ensure(void **ptr, long long *cap, long long need, size_t elem, cudaStream_t s);
template <typename T> static inline int ensure_buf(T **ptr, long long *cap, long long need, cudaStream_t s)
{
    return espic_ensure((void **)ptr, cap, need, sizeof(T), s);
}

// generic exclusive scan of uint32 (two level).  After the call: prefix of element w is
// pre[w] + coff[w >> SCAN_CHUNK_LOG2]; the grand total is in *d_total (device, 64 bit).
#define SCAN_CHUNK_LOG2 13
int espic_scan_u32(espic_ctx *ctx, const uint32_t *in, long long n, unsigned long long *d_total);

int espic_ensure_moments(espic_ctx *c, int sp);   // espic_api.cu: allocate + zero the 11 moment arrays of a species
int espic_repack_ef(espic_ctx *c);
void espic_mg_destroy(espic_ctx *c);   // espic_fields.cu     // espic_api.cu: refresh ef4 after ef was written from outside

// espic_comm.cu
void espic_comm_destroy(espic_ctx *c);
int  espic_comm_max_double(espic_ctx *c, double *v);
int  espic_comm_allreduce_acc(espic_ctx *c, Species &s);
int  espic_comm_allgather_doubles(espic_ctx *c, double *buf, size_t count);
int  espic_comm_allgather_bytes(espic_ctx *c, void *buf, size_t bytes);
int  espic_comm_exchange_segments(espic_ctx *c, int parts, int me, const double *const *sendp, const long long *sendn,
                                  double *(*recvp)[7], const long long *recvn);
void espic_migrate_destroy(espic_ctx *c);   // espic_migrate.cuh (espic_particles.cu)

// ---- device helpers ---------------------------------------------------------------------------

#ifdef __CUDACC__

__device__ __forceinline__ long long node_u(const MeshC &m, int i, int j, int k)
{
    return ((long long)k * m.nj + j) * (long long)m.ni + i;
}

// World::XtoL (World.h:75-81) + (int) truncation of Field::gather/scatter (Field.h:169-176).
// lc can round to exactly n-1 just below xm; the reference then reads node n (UB) with weight 0:
// clamp the cell to n-2 (fraction becomes exactly 1), see oracle cell_of().
// a / dh, IEEE correctly rounded, without the division sequence (about 10 FP64 instructions + MUFU.RCP64H): with
// y = RN(1/dh) from the host, q = RN(a*y), r = a - dh*q (exact, one FMA), q' = RN(q + r*y) is the correctly rounded
// quotient (Markstein's division theorem; checked against a/dh on 9.6e8 host samples incl. cell boundaries +-3 ulp,
// oracle/div_check.c).  Tiny |a| (subnormal intermediate results) takes the true division.
__device__ __forceinline__ double div_by_dh(double a, double dh, double rdh)
{
    if (fabs(a) < 1e-280) return a / dh;
    const double q = a * rdh;
    const double r = __fma_rn(-dh, q, a);
    return __fma_rn(r, rdh, q);
}

__device__ __forceinline__ void cell_frac(double x, double x0, double dh, double rdh, int n, int &i, double &d)
{
    double lc = div_by_dh(x - x0, dh, rdh);
    int ii = (int)lc;
    if (ii > n - 2) ii = n - 2;
    i = ii;
    d = lc - (double)ii;
}

// Field3::gather (Field.h:189-211): eight terms, each data*w_i*w_j*w_k left to right, summed in the reference's order.
// one node of the padded gather field: a single 256-bit read-only load (LDG.E.256, sm_100+)
__device__ __forceinline__ void ld_node(const double *__restrict__ ef4, long long u, double v[3])
{
    double pad;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(pad) : "l"(ef4 + 4 * u));
}

__device__ __forceinline__ void gather_ef(const MeshC &m, const double *__restrict__ ef4,
                                          int i, int j, int k, double di, double dj, double dk, double e[3])
{
    const long long u000 = node_u(m, i, j, k);
    const long long sj = m.ni, sk = (long long)m.ni * m.nj;
    const double ai = 1 - di, aj = 1 - dj, ak = 1 - dk;
    double n0[3], n1[3], n2[3], n3[3], n4[3], n5[3], n6[3], n7[3];
    ld_node(ef4, u000, n0);           ld_node(ef4, u000 + 1, n1);
    ld_node(ef4, u000 + 1 + sj, n2);  ld_node(ef4, u000 + sj, n3);
    ld_node(ef4, u000 + sk, n4);      ld_node(ef4, u000 + 1 + sk, n5);
    ld_node(ef4, u000 + 1 + sj + sk, n6); ld_node(ef4, u000 + sj + sk, n7);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double v = n0[c] * ai * aj * ak;
        v = v + n1[c] * di * aj * ak;
        v = v + n2[c] * di * dj * ak;
        v = v + n3[c] * ai * dj * ak;
        v = v + n4[c] * ai * aj * dk;
        v = v + n5[c] * di * aj * dk;
        v = v + n6[c] * di * dj * dk;
        v = v + n7[c] * ai * dj * dk;
        e[c] = v;
    }
}

// World::inSphere (World.cpp:118-125)
__device__ __forceinline__ bool in_sphere(const MeshC &m, double x, double y, double z)
{
    double r0 = x - m.sc[0], r1 = y - m.sc[1], r2 = z - m.sc[2];
    double r_mag2 = (r0 * r0 + r1 * r1 + r2 * r2);
    return r_mag2 <= m.sr2;
}

// World::inBounds (World.h:59-63)
__device__ __forceinline__ bool in_bounds(const MeshC &m, double x, double y, double z)
{
    if (x < m.x0[0] || x >= m.xm[0]) return false;
    if (y < m.x0[1] || y >= m.xm[1]) return false;
    if (z < m.x0[2] || z >= m.xm[2]) return false;
    return true;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum (fixed tree); result valid in thread 0.  blockDim.x must be a multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v, double *sh /* >= 32 */)
{
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (l < (int)(blockDim.x >> 5)) ? sh[l] : 0.0;
        v = warp_sum(v);
    }
    return v;
}

#endif  // __CUDACC__
