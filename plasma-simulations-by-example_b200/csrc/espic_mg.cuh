// espic_mg.cuh -- inexact Newton + aggregation-multigrid preconditioned CG for the Boltzmann-electron Poisson solve,
// the WHOLE nonlinear solve in one persistent cooperative kernel.  Textually included by espic_fields.cu (it uses
// StencilC, the node types and the reductions defined there).
//
// The reference runs Newton-Raphson with a Jacobi-preconditioned CG per step (PotentialSolver::solveNRPCG / solvePCGLinear,
// PotentialSolver.cpp:225-331): ~300 CG iterations per Newton step on the 128^3 mesh.  ESPIC_SOLVE_PCG_MG keeps the
// Newton iteration on the same discrete equations (the SPD system K = -(L - diag P) on the REG nodes, espic_fields.cu) and
// the reference's stopping tests, and changes how the work is done:
//   * preconditioner: one multigrid V(1,1) cycle instead of the matrix diagonal (:304) -- 2x2x2 node aggregation (a
//     direction whose spacing exceeds sqrt(2) x the smallest is not coarsened), piecewise-constant prolongation P,
//     restriction P^T, Galerkin coarse operators (7-point again: a diagonal and three link arrays per level), damped-Jacobi
//     smoothing (w = 0.9) fused with the grid transfers; the coarsest level (<= 4096 nodes) is solved redundantly by EVERY
//     block in its own shared memory (no grid barrier inside it).  The cycle is a fixed SPD operator, which CG requires.
//     Everything the preconditioner alone touches (z, the smoother's copy of the diagonal, all coarse levels) is stored in
//     FP32: it only has to be a fixed operator close to K^-1; r, d, delta and the operator itself stay FP64.
//   * inexact Newton: the linear solve of Newton step k stops at eta_k |R_k| with the Eisenstat-Walker forcing term
//     eta_k = gamma (|R_k| / |R_{k-1}|)^2 (eta_0 fixed) -- every quantity comes from THIS solve, the result is a pure
//     function of (rho, phi).  The solve ends when the reference's test holds (update norm sqrt(sum y^2 / n) < nr_tol,
//     :282-285) AND the nonlinear residual of the equations the reference's solveGS tests (:389-421) is below tol,
//     sqrt(sum R^2 / n) < tol (or, at the rounding floor of that residual, after a linear solve that went to tol/2): at
//     least as strict as the reference, which stops on the update alone.
//   * fine-level passes march along k with the column's values in registers (each vector is read once per pass, the x/y
//     neighbours come from L1); the pre-smoothed iterate x0 = w D^-1 r and q = K d are recomputed instead of stored:
//     96 B per node and CG iteration instead of 170.
//   * one launch and one host synchronisation per solve (the round-1 version: 3 synchronisations per Newton step).
#pragma once
#include <cuda.h>       // CUtensorMap (the encoder itself is fetched through cudaGetDriverEntryPoint: no link against libcuda)

#define MG_MAX_LEVELS 8
#define MG_OMEGA 0.9                    // damped Jacobi on the fine level
// ... and on the coarse levels: 1.0 is the largest factor for which Jacobi is a convergent smoother whatever the mesh (the
// spectrum of D^-1 K ends just below 2: 1.993 on the 128^3 bench mesh), i.e. for which the V-cycle is guaranteed SPD.  With the
// rescaled coarse operators it saves CG iterations at no cost (scripts/mg_variants_prototype.py: 25 -> 22 per solve together
// with three more sweeps on the coarsest level; 1.1-1.3 would save more but are not safe, 1.6 diverges).
#define MG_OMEGA_C 1.0
#define MG_COARSE_SWEEPS 9
#define MG_COARSEST_NODES 4096          // stop coarsening once a level is this small; it must fit 3 FP32 vectors in shared memory
#define MG_SLAB_REDUNDANT_NODES 65536   // slab mode: levels up to this size are solved by every rank in full
#define MG_THREADS 256                  // one 32 x 8 tile of columns per block and marching step
#define MG_TX 32
#define MG_TY 8

typedef float mgf;                      // storage type of everything only the preconditioner reads

struct MgLevel {
    int ni, nj, nk;
    int fi, fj, fk;           // log2 of the coarsening factor from the next finer level to this one, per dimension (0 or 1)
    long long nn;
    int pitch;                // elements per row of this level's arrays (level 0: ni rounded up to 4 for the TMA tiles; else ni)
    mgf *diag, *minv;         // diagonal and its inverse (0 on nodes without unknowns): dlx + dly + dlz + mass
    mgf *cx, *cy, *cz;        // link to the +x / +y / +z neighbour (>= 0; K = diag - sum links); level 0: not stored
    mgf *x, *xn, *b;          // pre-smoothed iterate, post-smoothed iterate, right-hand side
    mgf *dlx, *dly, *dlz;     // Laplacian part of the diagonal per direction (geometry only, set up once)
    mgf *mass;                // Boltzmann part of the diagonal: sum of the children's (changes with phi every Newton step)
};
#define MG_LEVEL_ARRAYS 12

struct MgHierarchy {
    int nlev = 0;
    MgLevel L[MG_MAX_LEVELS];
    long long geom_version = -1;
    mgf *pool = nullptr;      // one allocation for all coarse-level arrays
    bool own_pool = false;
    // single-GPU solver: the fine vectors in the padded layout (r | delta | d0 | d1 doubles, z | diagf floats)
    double *fine = nullptr;
    long long fine_nnp = 0;
};

static MgHierarchy *g_mg_of(espic_ctx *c);   // stored in the context (espic_internal.cuh: void *mg)

// ---- setup kernels (geometry only: links never change with phi) --------------------------------------------------------

// Coarse operators.  Plain Galerkin P^T K P with piecewise-constant P is too stiff in every coarsened direction: summing the
// 2x2 fine links across an aggregate face gives the 7-point Laplacian of spacing 2h times 2 (relative to what the summed
// restriction of the right-hand side calls for), so smooth error is under-corrected by that factor and CG needs ~45 % more
// iterations (measured on the bench problem: 28 -> 19 for six decades).  The links and the Laplacian part of the diagonal
// are therefore scaled by `scale` (1/2 = rediscretisation on the coarse mesh; 1 = Galerkin) in each direction that is
// coarsened, level by level; the Boltzmann (mass) part of the diagonal is restricted exactly.  Any SPD coarse operator keeps
// the V-cycle SPD (symmetric smoother, R = P^T), which is all CG needs.

// level 1 from the fine node types: a fine link (u, u+e) exists iff both ends are REG, weight g = 1/dh^2; the fine Laplacian
// diagonal of a REG node is 2g per direction minus g per Neumann face neighbour (folded into the node, espic_fields.cu)
__global__ void __launch_bounds__(256) k_mg_setup_from_types(StencilC s, const uint8_t *__restrict__ type, MgLevel C, double scale)
{
    long long I = blockIdx.x * 256ll + threadIdx.x;
    if (I >= C.nn) return;
    const int ci = (int)(I % C.ni), cj = (int)((I / C.ni) % C.nj), ck = (int)(I / ((long long)C.ni * C.nj));
    double lx = 0, ly = 0, lz = 0, dx = 0, dy = 0, dz = 0;
    for (int dk = 0; dk <= C.fk; dk++)
        for (int dj = 0; dj <= C.fj; dj++)
            for (int di = 0; di <= C.fi; di++) {
                const int i = (ci << C.fi) + di, j = (cj << C.fj) + dj, k = (ck << C.fk) + dk;
                if (i >= s.ni || j >= s.nj || k >= s.nk) continue;
                const long long u = (long long)k * s.sk + (long long)j * s.sj + i;
                if (type[u] != NT_REG) continue;            // REG nodes are interior: all six neighbours exist
                dx += s.gdx2 * (2 - (type[u - 1] >= NT_I0) - (type[u + 1] >= NT_I0));
                dy += s.gdy2 * (2 - (type[u - s.sj] >= NT_I0) - (type[u + s.sj] >= NT_I0));
                dz += s.gdz2 * (2 - (type[u - s.sk] >= NT_I0) - (type[u + s.sk] >= NT_I0));
                if (type[u + 1] == NT_REG) { if (di == C.fi) lx += s.gdx2; else dx -= 2 * s.gdx2; }
                if (type[u + s.sj] == NT_REG) { if (dj == C.fj) ly += s.gdy2; else dy -= 2 * s.gdy2; }
                if (type[u + s.sk] == NT_REG) { if (dk == C.fk) lz += s.gdz2; else dz -= 2 * s.gdz2; }
            }
    const double sx = C.fi ? scale : 1.0, sy = C.fj ? scale : 1.0, sz = C.fk ? scale : 1.0;
    C.cx[I] = (mgf)(sx * lx); C.cy[I] = (mgf)(sy * ly); C.cz[I] = (mgf)(sz * lz);
    C.dlx[I] = (mgf)(sx * dx); C.dly[I] = (mgf)(sy * dy); C.dlz[I] = (mgf)(sz * dz);
}

// level l+1 from level l (l >= 1)
__global__ void __launch_bounds__(256) k_mg_setup_from_level(MgLevel F, MgLevel C, double scale)
{
    long long I = blockIdx.x * 256ll + threadIdx.x;
    if (I >= C.nn) return;
    const int ci = (int)(I % C.ni), cj = (int)((I / C.ni) % C.nj), ck = (int)(I / ((long long)C.ni * C.nj));
    double lx = 0, ly = 0, lz = 0, dx = 0, dy = 0, dz = 0;
    for (int dk = 0; dk <= C.fk; dk++)
        for (int dj = 0; dj <= C.fj; dj++)
            for (int di = 0; di <= C.fi; di++) {
                const int i = (ci << C.fi) + di, j = (cj << C.fj) + dj, k = (ck << C.fk) + dk;
                if (i >= F.ni || j >= F.nj || k >= F.nk) continue;
                const long long u = ((long long)k * F.nj + j) * F.ni + i;
                dx += F.dlx[u]; dy += F.dly[u]; dz += F.dlz[u];
                if (di == C.fi) lx += F.cx[u]; else dx -= 2 * (double)F.cx[u];       // a link at the mesh edge is 0
                if (dj == C.fj) ly += F.cy[u]; else dy -= 2 * (double)F.cy[u];
                if (dk == C.fk) lz += F.cz[u]; else dz -= 2 * (double)F.cz[u];
            }
    const double sx = C.fi ? scale : 1.0, sy = C.fj ? scale : 1.0, sz = C.fk ? scale : 1.0;
    C.cx[I] = (mgf)(sx * lx); C.cy[I] = (mgf)(sy * ly); C.cz[I] = (mgf)(sz * lz);
    C.dlx[I] = (mgf)(sx * dx); C.dly[I] = (mgf)(sy * dy); C.dlz[I] = (mgf)(sz * dz);
}

// ---- who owns what: single GPU, or one k-slab per rank with peer-mapped pools -----------------------------------------
// Every pass below is written once and instantiated twice.  `Own` tells a pass which planes of a level this rank computes
// and what a store has to do besides writing locally:
//   OwnAll   one GPU: the whole level, plain stores, grid-wide barrier.
//   OwnSlab  rank r of R owns planes [k0,k1) of every level (boundaries are multiples of 2^(k-coarsenings) fine planes, so an
//            aggregate never straddles two ranks).  All solver vectors live at the same offset of a pool that every rank
//            maps from every other rank (CUDA IPC over NVLink).  A value written on the first / last plane of the slab is
//            ALSO stored straight into the lower / upper neighbour's copy (peer store fused into the producing pass: the
//            halo exchange costs no extra pass and no NCCL call); dot-product partials are stored to every rank so all
//            ranks add the same numbers in the same order; the barrier is a grid barrier plus one flag per peer.
struct OwnAll {
    static constexpr bool slab = false;
    __device__ __forceinline__ int klo(const MgLevel &, int) const { return 0; }
    __device__ __forceinline__ int khi(const MgLevel &L, int) const { return L.nk; }
    __device__ __forceinline__ long long lo(const MgLevel &L, int) const { return 0; }
    __device__ __forceinline__ long long hi(const MgLevel &L, int) const { return (long long)L.pitch * L.nj * L.nk; }
    template <typename T> __device__ __forceinline__ void st(T *A, long long u, const MgLevel &, int, T v) const { A[u] = v; }
    // store at flat index u known to lie in plane k of level l (any row pitch)
    template <typename T> __device__ __forceinline__ void st_k(T *A, long long u, int, int, T v) const { A[u] = v; }
    template <typename T> __device__ __forceinline__ void st_all(T *A, long long u, T v) const { A[u] = v; }
    __device__ __forceinline__ int nparts(int nb) const { return nb; }
    __device__ __forceinline__ void put_partial(double *base, int nb, double v) const { base[blockIdx.x] = v; }
    __device__ __forceinline__ bool barrier(cg::grid_group &grid) { grid.sync(); return true; }
    __device__ __forceinline__ bool lbarrier(cg::grid_group &grid) { grid.sync(); return true; }
    __device__ __forceinline__ bool nbarrier(cg::grid_group &grid) { grid.sync(); return true; }
};

#define MG_MAX_RANKS 8
// A barrier that waits longer than this many polls (~ seconds) gives up: it raises the abort word on every rank, every
// later barrier returns at once and the kernel unwinds with an error code instead of hanging 8 GPUs on one faulted peer.
#define MG_SPIN_BUDGET (1ull << 31)
struct OwnSlab {
    static constexpr bool slab = true;
    int rank, nranks;
    int k0[MG_MAX_LEVELS], k1[MG_MAX_LEVELS];
    long long peer[MG_MAX_RANKS];        // byte distance from an address in this rank's pool to the same address in rank p's pool
    unsigned long long *flags;           // in the pool: flags[p] = last all-rank barrier epoch rank p has reached (written by rank p)
    unsigned long long *nflags;          // the same for the neighbour-only barriers
    unsigned long long nepoch;           // neighbour-only barriers passed so far
    unsigned long long lepoch_all;       // barriers of either kind passed so far (local arrive / release counters)
    int neighbour_sync;                  // 0: nbarrier() is a full barrier (ESPIC_MG_SLAB_NEIGHBOUR_SYNC=0)
    unsigned long long *abort_word;      // in the pool: non-zero once any rank gave up waiting (written by that rank to every pool)
    unsigned long long epoch;            // barriers passed so far (identical on every rank)
    unsigned long long *arrive, *release;     // in the local pool
    unsigned long long *larrive, *lrelease;   // the same pair for the barrier among this rank's blocks only
    unsigned long long lepoch;
    __device__ __forceinline__ int klo(const MgLevel &, int l) const { return k0[l]; }
    __device__ __forceinline__ int khi(const MgLevel &, int l) const { return k1[l]; }
    __device__ __forceinline__ long long lo(const MgLevel &L, int l) const { return (long long)k0[l] * L.pitch * L.nj; }
    __device__ __forceinline__ long long hi(const MgLevel &L, int l) const { return (long long)k1[l] * L.pitch * L.nj; }
    template <typename T> __device__ __forceinline__ T *at(T *p, int r) const
    {
        return reinterpret_cast<T *>(reinterpret_cast<char *>(p) + peer[r]);
    }
    template <typename T> __device__ __forceinline__ void st(T *A, long long u, const MgLevel &L, int l, T v) const
    {
        A[u] = v;
        const long long plane = (long long)L.pitch * L.nj;
        if (rank > 0 && u < lo(L, l) + plane) *at(A + u, rank - 1) = v;
        if (rank + 1 < nranks && u >= hi(L, l) - plane) *at(A + u, rank + 1) = v;
    }
    // store at flat index u known to lie in plane k of level l (any row pitch): first / last plane of the slab goes to the neighbour too
    template <typename T> __device__ __forceinline__ void st_k(T *A, long long u, int k, int l, T v) const
    {
        A[u] = v;
        if (rank > 0 && k == k0[l]) *at(A + u, rank - 1) = v;
        if (rank + 1 < nranks && k == k1[l] - 1) *at(A + u, rank + 1) = v;
    }
    // store into every rank's copy (the right-hand side of the first level that every rank solves in full)
    template <typename T> __device__ __forceinline__ void st_all(T *A, long long u, T v) const
    {
        for (int p = 0; p < nranks; p++) *at(A + u, p) = v;
    }
    __device__ __forceinline__ int nparts(int nb) const { return nb * nranks; }
    __device__ __forceinline__ void put_partial(double *base, int nb, double v) const
    {
        for (int p = 0; p < nranks; p++) *at(base + rank * nb + blockIdx.x, p) = v;
    }
    // All blocks of the ranks [pfirst, plast] (this rank included).  Local arrival (one atomic per block on a cumulative
    // counter), then block 0 exchanges one flag with every peer of the set over NVLink, then it releases the local blocks:
    // one local round trip plus one remote one.
    // Ordering: every block's thread 0 fences at GPU scope after the block barrier and before arriving; block 0, having seen
    // all arrivals, fences once at SYSTEM scope before it signals the peers (fences are cumulative: what block 0 has observed
    // is ordered before its own later writes at the wider scope), so a peer that sees the flag also sees every halo value and
    // partial sum stored before this barrier.  Returns false once the abort word is up.
    // kind 0: every rank (flags[]), kind 1: the two neighbours only (nflags[]: halo planes need no more).
    __device__ __forceinline__ bool sync_ranks(int kind)
    {
        __syncthreads();
        unsigned long long &ep = kind == 0 ? epoch : nepoch;
        ep++;
        lepoch_all++;
        volatile unsigned long long *ab = abort_word;
        unsigned long long *fl = kind == 0 ? flags : nflags;
        const int pfirst = kind == 0 ? 0 : max(rank - 1, 0), plast = kind == 0 ? nranks - 1 : min(rank + 1, nranks - 1);
        if (blockIdx.x == 0) {
            if (threadIdx.x < 32) {          // first warp: lane 0 collects the local arrivals, lane p talks to peer pfirst + p
                const int lane = threadIdx.x;
                if (lane == 0) {
                    volatile unsigned long long *arr = arrive;
                    const unsigned long long want = (unsigned long long)(gridDim.x - 1) * lepoch_all;
                    unsigned long long spins = 0;
                    while (*arr < want && !*ab) { if (++spins > MG_SPIN_BUDGET) break; }
                    __threadfence_system();
                }
                __syncwarp();
                const int peer_rank = pfirst + lane;
                if (peer_rank <= plast) {
                    *reinterpret_cast<volatile unsigned long long *>(at(fl + rank, peer_rank)) = ep;
                    volatile unsigned long long *mine = fl + peer_rank;
                    unsigned long long spins = 0;
                    while (*mine < ep && !*ab) {
                        if (++spins > MG_SPIN_BUDGET) {          // a peer never arrived: tell everybody and give up
                            for (int p = 0; p < nranks; p++) *reinterpret_cast<volatile unsigned long long *>(at(abort_word, p)) = ep;
                            break;
                        }
                    }
                    __threadfence_system();
                }
                __syncwarp();
                if (lane == 0) {
                    volatile unsigned long long *rel = release;
                    *rel = lepoch_all;
                    __threadfence();
                }
            }
        } else if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(arrive, 1ull);
            volatile unsigned long long *rel = release;
            while (*rel < lepoch_all && !*ab) { }
            __threadfence();
        }
        __syncthreads();
        return *ab == 0;
    }
    __device__ __forceinline__ bool barrier(cg::grid_group &) { return sync_ranks(0); }
    __device__ __forceinline__ bool nbarrier(cg::grid_group &) { return sync_ranks(neighbour_sync ? 1 : 0); }
    // the blocks of THIS rank only (levels every rank solves in full): same arrive / release scheme without the peer exchange
    __device__ __forceinline__ bool lbarrier(cg::grid_group &)
    {
        __syncthreads();
        lepoch++;
        volatile unsigned long long *ab = abort_word;
        if (threadIdx.x == 0) {
            __threadfence();
            volatile unsigned long long *rel = lrelease;
            if (blockIdx.x == 0) {
                volatile unsigned long long *arr = larrive;
                const unsigned long long want = (unsigned long long)(gridDim.x - 1) * lepoch;
                while (*arr < want && !*ab) { }
                __threadfence();
                *rel = lepoch;
            } else {
                atomicAdd(larrive, 1ull);
                while (*rel < lepoch && !*ab) { }
            }
            __threadfence();
        }
        __syncthreads();
        return *ab == 0;
    }
};

// ---- coarse-level passes (grid-stride; the caller separates them with barriers) --------------------------------------
// The coarse levels are small: what matters there is the length of the dependent-load chain of a thread, not bandwidth.
// The passes spread one node over 8 consecutive lanes (the 8 children of an aggregate, or the 7 stencil terms of a node)
// and combine with three shuffles; the sum order is fixed, so results are reproducible.

// Index arithmetic of the coarse passes is 32 bit (a coarse level has < 2^31 nodes) and every load is issued before the first
// use: a pass is one memory round trip per node, not a chain (node data -> branch -> neighbour data).

struct MgIdx { int i, j, k; };
__device__ __forceinline__ MgIdx mg_ijk(const MgLevel &L, unsigned u)
{
    const unsigned ni = (unsigned)L.ni, nij = (unsigned)(L.ni * L.nj);
    MgIdx q;
    q.k = (int)(u / nij);
    const unsigned r = u - (unsigned)q.k * nij;
    q.j = (int)(r / ni);
    q.i = (int)(r - (unsigned)q.j * ni);
    return q;
}

__device__ __forceinline__ mgf mg_sum8(mgf v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// The whole V-cycle is evaluated in FP32 (storage AND arithmetic): it only has to be a fixed SPD operator close to K^-1
// (scripts/mg_fp32_prototype.py: same CG iteration counts as in FP64), FP32 instructions issue at twice the FP64 rate and
// nothing has to be converted.  Sums use a fixed order (shuffle trees), so every block and every rank gets the same bits.

// Down pass between coarse levels F (level lf) -> C: x_F = w D^-1 b_F (smoothing from zero), residual b_F - K x_F summed
// over each aggregate -> b_C.  Lane c of a group of 8 handles child c of coarse node I.
template <class Own>
__device__ __forceinline__ void mg_down(const Own &own, const MgLevel &F, int lf, const MgLevel &C, long long t0, long long stride,
                                        bool to_all)
{
    const long long first = own.lo(C, lf + 1), total = (own.hi(C, lf + 1) - first) * 8;
    const int sj = F.ni, sk = F.ni * F.nj;
    const mgf W = (mgf)MG_OMEGA_C;
    for (long long w = t0; (w & ~31LL) < total; w += stride) {
        const unsigned I = (unsigned)(first + (w >> 3));
        const int c = (int)(w & 7);
        const bool live = w < total;
        mgf res = 0;
        if (live) {
            const MgIdx q = mg_ijk(C, I);
            const int di = c & 1, dj = (c >> 1) & 1, dk = c >> 2;
            const int i = (q.i << C.fi) + di, j = (q.j << C.fj) + dj, k = (q.k << C.fk) + dk;
            if (di <= C.fi && dj <= C.fj && dk <= C.fk && i < F.ni && j < F.nj && k < F.nk) {
                const int u = (k * F.nj + j) * F.ni + i;
                const bool xm = i > 0, xp = i + 1 < F.ni, ym = j > 0, yp = j + 1 < F.nj, zm = k > 0, zp = k + 1 < F.nk;
                // everything this child needs, loaded up front (index predicates only)
                const mgf mi = F.minv[u], bu = F.b[u], dg = F.diag[u];
                const mgf lxm = xm ? F.cx[u - 1] : (mgf)0, lxp = xp ? F.cx[u] : (mgf)0;
                const mgf lym = ym ? F.cy[u - sj] : (mgf)0, lyp = yp ? F.cy[u] : (mgf)0;
                const mgf lzm = zm ? F.cz[u - sk] : (mgf)0, lzp = zp ? F.cz[u] : (mgf)0;
                const mgf bxm = xm ? F.b[u - 1] : (mgf)0, bxp = xp ? F.b[u + 1] : (mgf)0;
                const mgf bym = ym ? F.b[u - sj] : (mgf)0, byp = yp ? F.b[u + sj] : (mgf)0;
                const mgf bzm = zm ? F.b[u - sk] : (mgf)0, bzp = zp ? F.b[u + sk] : (mgf)0;
                const mgf mxm = xm ? F.minv[u - 1] : (mgf)0, mxp = xp ? F.minv[u + 1] : (mgf)0;
                const mgf mym = ym ? F.minv[u - sj] : (mgf)0, myp = yp ? F.minv[u + sj] : (mgf)0;
                const mgf mzm = zm ? F.minv[u - sk] : (mgf)0, mzp = zp ? F.minv[u + sk] : (mgf)0;
                const mgf xu = W * bu * mi;
                const mgf off = W * (lxm * (bxm * mxm) + lxp * (bxp * mxp) + lym * (bym * mym) + lyp * (byp * myp) + lzm * (bzm * mzm) + lzp * (bzp * mzp));
                res = mi != (mgf)0 ? bu - (dg * xu - off) : (mgf)0;
                own.st(F.x, (long long)u, F, lf, xu);
            }
        }
        res = mg_sum8(res);
        if (c == 0 && live) {
            if (to_all) own.st_all(C.b, (long long)I, res);
            else own.st(C.b, (long long)I, C, lf + 1, res);
        }
    }
}

// Up pass on a coarse level F (index lf) with the correction e of the next coarser level C:
//   xn = (x + P e) + w D^-1 (b - K (x + P e))
// A big level takes one thread per node, a small one 8 lanes per node (one stencil term each, combined with three shuffles:
// what matters on the small levels is the length of a thread's dependent chain, not bandwidth).
template <class Own>
__device__ __forceinline__ void mg_up(const Own &own, const MgLevel &F, int lf, const MgLevel &C, const mgf *e,
                                      long long t0, long long stride)
{
    const long long first = own.lo(F, lf), count = own.hi(F, lf) - first;
    const int sj = F.ni, sk = F.ni * F.nj, csj = C.ni, csk = C.ni * C.nj;
    const mgf W = (mgf)MG_OMEGA_C;
    // value of the prolongated iterate at node (vi,vj,vk) = flat v; links to nodes without unknowns are zero on coarse levels,
    // so no mask is needed on the neighbours
    auto val = [&](int v, int vi, int vj, int vk) {
        return F.x[v] + e[(vk >> C.fk) * csk + (vj >> C.fj) * csj + (vi >> C.fi)];
    };
    if (count * 2 > stride) {
        for (long long uu = first + t0; uu < first + count; uu += stride) {
            const int u = (int)uu;
            const MgIdx q = mg_ijk(F, (unsigned)u);
            const int i = q.i, j = q.j, k = q.k;
            const bool xm = i > 0, xp = i + 1 < F.ni, ym = j > 0, yp = j + 1 < F.nj, zm = k > 0, zp = k + 1 < F.nk;
            const mgf mi = F.minv[u], bu = F.b[u], dg = F.diag[u];
            const mgf lxm = xm ? F.cx[u - 1] : (mgf)0, lxp = xp ? F.cx[u] : (mgf)0;
            const mgf lym = ym ? F.cy[u - sj] : (mgf)0, lyp = yp ? F.cy[u] : (mgf)0;
            const mgf lzm = zm ? F.cz[u - sk] : (mgf)0, lzp = zp ? F.cz[u] : (mgf)0;
            const mgf vc = val(u, i, j, k);
            const mgf vxm = xm ? val(u - 1, i - 1, j, k) : (mgf)0, vxp = xp ? val(u + 1, i + 1, j, k) : (mgf)0;
            const mgf vym = ym ? val(u - sj, i, j - 1, k) : (mgf)0, vyp = yp ? val(u + sj, i, j + 1, k) : (mgf)0;
            const mgf vzm = zm ? val(u - sk, i, j, k - 1) : (mgf)0, vzp = zp ? val(u + sk, i, j, k + 1) : (mgf)0;
            const mgf tot = (bu - dg * vc) + ((lxm * vxm + lxp * vxp) + (lym * vym + lyp * vyp) + (lzm * vzm + lzp * vzp));
            own.st(F.xn, uu, F, lf, mi != (mgf)0 ? vc + W * mi * tot : (mgf)0);
        }
        return;
    }
    const long long total = count * 8;
    for (long long w = t0; (w & ~31LL) < total; w += stride) {
        const int u = (int)(first + (w >> 3));
        const int c = (int)(w & 7);
        const bool live = w < total;
        mgf term = 0, centre = 0, mi = 0;
        if (live) {
            const MgIdx q = mg_ijk(F, (unsigned)u);
            const int i = q.i, j = q.j, k = q.k;
            mi = F.minv[u];
            // term c: 0 centre (b - diag*v), 1..6 the six links, 7 nothing -- every address is known without waiting for data
            int v = u, vi = i, vj = j, vk = k, lu = u;
            bool ok = true;
            const mgf *lnk = F.cx;
            switch (c) {
                case 1: ok = i > 0; v = u - 1; vi = i - 1; lu = u - 1; break;
                case 2: ok = i + 1 < F.ni; v = u + 1; vi = i + 1; break;
                case 3: ok = j > 0; v = u - sj; vj = j - 1; lu = u - sj; lnk = F.cy; break;
                case 4: ok = j + 1 < F.nj; v = u + sj; vj = j + 1; lnk = F.cy; break;
                case 5: ok = k > 0; v = u - sk; vk = k - 1; lu = u - sk; lnk = F.cz; break;
                case 6: ok = k + 1 < F.nk; v = u + sk; vk = k + 1; lnk = F.cz; break;
                case 7: ok = false; break;
                default: break;
            }
            if (ok) {
                const mgf vv = val(v, vi, vj, vk);
                if (c == 0) { centre = vv; term = F.b[u] - F.diag[u] * vv; }
                else term = lnk[lu] * vv;
            }
        }
        const mgf tot = mg_sum8(term);
        centre = __shfl_sync(0xffffffffu, centre, (threadIdx.x & 31) & ~7);
        if (c == 0 && live) own.st(F.xn, (long long)u, F, lf, (mi != (mgf)0) ? centre + W * mi * tot : (mgf)0);
    }
}

// The coarsest level, solved by every block on its own in shared memory: x = w D^-1 b, then `sweeps` damped-Jacobi sweeps in
// FP32.  Identical arithmetic in every block (and on every rank): identical result everywhere, no grid barrier.
// Shared-memory layout (floats): eight arrays of n + 2 pad entries, pad = one plane, zero filled -- b | x | y | minv | diag |
// cx | cy | cz.  Links to neighbours that do not exist are zero in the link arrays and every index u +- 1, +- ni, +- ni*nj
// stays inside the padded arrays, so a sweep has no index tests at all.  A level of more than MG_STAGE_NODES nodes keeps
// its coefficients in global memory (three vectors in shared memory only).
#define MG_STAGE_NODES 2048

__device__ __forceinline__ int mg_cpad(const MgLevel &L) { return L.ni * L.nj; }
__device__ __forceinline__ bool mg_cstaged(const MgLevel &L) { return L.nn <= MG_STAGE_NODES; }

// once per Newton step, after the Galerkin diagonals are complete: coefficients -> shared memory, pads -> 0
__device__ __forceinline__ void mg_coarsest_stage(const MgLevel &L, mgf *smem)
{
    const int n = (int)L.nn, pad = mg_cpad(L), len = n + 2 * pad;
    if (!mg_cstaged(L)) {
        for (int t = threadIdx.x; t < 3 * n; t += blockDim.x) smem[t] = 0;
        __syncthreads();
        return;
    }
    for (int t = threadIdx.x; t < 8 * len; t += blockDim.x) smem[t] = 0;
    __syncthreads();
    mgf *minv = smem + 3 * len + pad, *diag = minv + len, *cx = diag + len, *cy = cx + len, *cz = cy + len;
    for (int u = threadIdx.x; u < n; u += blockDim.x) {
        minv[u] = L.minv[u]; diag[u] = L.diag[u]; cx[u] = L.cx[u]; cy[u] = L.cy[u]; cz[u] = L.cz[u];
    }
    __syncthreads();
}

// Returns the shared-memory array that holds the solution (index 0 = node 0).  Ends with a __syncthreads.
__device__ __forceinline__ const mgf *mg_coarsest(const MgLevel &L, int sweeps, mgf *smem)
{
    const int n = (int)L.nn;
    const int sj = L.ni, sk = L.ni * L.nj;
    if (mg_cstaged(L)) {
        const int pad = sk, len = n + 2 * pad;
        mgf *sb = smem + pad, *sx = sb + len, *sy = sx + len;
        const mgf *minv = sy + len, *diag = minv + len, *cx = diag + len, *cy = cx + len, *cz = cy + len;
        for (int u = threadIdx.x; u < n; u += blockDim.x) {
            const mgf b = L.b[u];
            sb[u] = b;
            sx[u] = (mgf)MG_OMEGA_C * b * minv[u];
        }
        __syncthreads();
        for (int sweep = 0; sweep < sweeps; sweep++) {
            for (int u = threadIdx.x; u < n; u += blockDim.x) {
                const mgf xc = sx[u];
                mgf t = sb[u] - diag[u] * xc;
                t += cx[u - 1] * sx[u - 1];
                t += cx[u] * sx[u + 1];
                t += cy[u - sj] * sx[u - sj];
                t += cy[u] * sx[u + sj];
                t += cz[u - sk] * sx[u - sk];
                t += cz[u] * sx[u + sk];
                sy[u] = xc + (mgf)MG_OMEGA_C * minv[u] * t;          // minv = 0 on nodes without unknowns, where x stays 0
            }
            __syncthreads();
            mgf *tmp = sx; sx = sy; sy = tmp;
        }
        return sx;
    }
    mgf *sb = smem, *sx = smem + n, *sy = smem + 2 * n;
    for (int u = threadIdx.x; u < n; u += blockDim.x) {
        const mgf b = L.b[u];
        sb[u] = b;
        sx[u] = (mgf)MG_OMEGA_C * b * L.minv[u];
    }
    __syncthreads();
    for (int sweep = 0; sweep < sweeps; sweep++) {
        for (int u = threadIdx.x; u < n; u += blockDim.x) {
            const MgIdx q = mg_ijk(L, (unsigned)u);
            const mgf xc = sx[u];
            mgf t = sb[u] - L.diag[u] * xc;
            if (q.i > 0) t += L.cx[u - 1] * sx[u - 1];
            if (q.i + 1 < L.ni) t += L.cx[u] * sx[u + 1];
            if (q.j > 0) t += L.cy[u - sj] * sx[u - sj];
            if (q.j + 1 < L.nj) t += L.cy[u] * sx[u + sj];
            if (q.k > 0) t += L.cz[u - sk] * sx[u - sk];
            if (q.k + 1 < L.nk) t += L.cz[u] * sx[u + sk];
            sy[u] = xc + (mgf)MG_OMEGA_C * L.minv[u] * t;
        }
        __syncthreads();
        mgf *tmp = sx; sx = sy; sy = tmp;
    }
    return sx;
}

// ---- coarse diagonals, every Newton step: the Boltzmann term changes with phi, links and Laplacian parts do not ----------

// level 1: mass = sum over the children of (Jacobian diagonal - Laplacian diagonal)
template <class Own>
__device__ __forceinline__ void mg_diag1(const Own &own, const StencilC &s, const uint8_t *__restrict__ type, const mgf *diagf, int pitch,
                                         const MgLevel &C, long long t0, long long stride, bool to_all)
{
    for (long long I = own.lo(C, 1) + t0; I < own.hi(C, 1); I += stride) {
        const int ci = (int)(I % C.ni), cj = (int)((I / C.ni) % C.nj), ck = (int)(I / ((long long)C.ni * C.nj));
        double m = 0;
        for (int dk = 0; dk <= C.fk; dk++)
            for (int dj = 0; dj <= C.fj; dj++)
                for (int di = 0; di <= C.fi; di++) {
                    const int i = (ci << C.fi) + di, j = (cj << C.fj) + dj, k = (ck << C.fk) + dk;
                    if (i >= s.ni || j >= s.nj || k >= s.nk) continue;
                    const long long u = (long long)k * s.sk + (long long)j * s.sj + i;
                    if (type[u] != NT_REG) continue;
                    const double d0 = s.gdx2 * (2 - (type[u - 1] >= NT_I0) - (type[u + 1] >= NT_I0)) +
                                      s.gdy2 * (2 - (type[u - s.sj] >= NT_I0) - (type[u + s.sj] >= NT_I0)) +
                                      s.gdz2 * (2 - (type[u - s.sk] >= NT_I0) - (type[u + s.sk] >= NT_I0));
                    m += fmax((double)diagf[((long long)k * s.nj + j) * pitch + i] - d0, 0.0);
                }
        const double d = (double)C.dlx[I] + (double)C.dly[I] + (double)C.dlz[I] + m;
        const mgf df = (mgf)d, mi = d > 0 ? (mgf)(1.0 / d) : (mgf)0, mf = (mgf)m;
        if (to_all) { own.st_all(C.diag, I, df); own.st_all(C.minv, I, mi); own.st_all(C.mass, I, mf); }
        else { own.st(C.diag, I, C, 1, df); own.st(C.minv, I, C, 1, mi); C.mass[I] = mf; }
    }
}

template <class Own>
__device__ __forceinline__ void mg_diagl(const Own &own, const MgLevel &F, const MgLevel &C, int lc, long long t0, long long stride,
                                         bool to_all)
{
    for (long long I = own.lo(C, lc) + t0; I < own.hi(C, lc); I += stride) {
        const int ci = (int)(I % C.ni), cj = (int)((I / C.ni) % C.nj), ck = (int)(I / ((long long)C.ni * C.nj));
        double m = 0;
        for (int dk = 0; dk <= C.fk; dk++)
            for (int dj = 0; dj <= C.fj; dj++)
                for (int di = 0; di <= C.fi; di++) {
                    const int i = (ci << C.fi) + di, j = (cj << C.fj) + dj, k = (ck << C.fk) + dk;
                    if (i >= F.ni || j >= F.nj || k >= F.nk) continue;
                    m += (double)F.mass[((long long)k * F.nj + j) * F.ni + i];
                }
        const double d = (double)C.dlx[I] + (double)C.dly[I] + (double)C.dlz[I] + m;
        const mgf df = (mgf)d, mi = d > 0 ? (mgf)(1.0 / d) : (mgf)0, mf = (mgf)m;
        if (to_all) { own.st_all(C.diag, I, df); own.st_all(C.minv, I, mi); own.st_all(C.mass, I, mf); }
        else { own.st(C.diag, I, C, lc, df); own.st(C.minv, I, C, lc, mi); C.mass[I] = mf; }
    }
}

// ---- the kernel's arguments ------------------------------------------------------------------------------------------

struct MgnArgs {
    StencilC s;
    int nlev;
    int coarse_sweeps;            // Jacobi sweeps on the coarsest level
    int first_redundant;          // slab mode: first level that every rank solves in full (<= nlev-1, >= 1); unused on one GPU
    MgLevel L[MG_MAX_LEVELS];     // L[0]: dimensions only
    const uint8_t *type;
    const double *rho;
    double *phi;                  // the potential the kernel works on (slab mode: the pool copy, halos kept current by peer stores)
    double *phi_user;             // slab mode: the context's phi (read at the start, this rank's slab written at the end); else == phi
    // the solver's own fine vectors: rows padded to `pitch` elements (a multiple of 4, the pads stay zero) so that TMA can tile them
    int pitch;
    long long psk;                // pitch * nj: plane stride of the padded arrays
    mgf *diagf;                   // Jacobian diagonal as the operator uses it (FP32 storage)
    mgf *winv;                    // w / diag: the damped-Jacobi smoother's factor (0 on nodes without unknowns)
    mgf *rf;                      // FP32 copy of r: what the V-cycle reads
    mgf *z;
    double *r, *d0, *d1, *delta;
    // 3-D tensor maps (x = pitch, y = nj, z = nk) of the vectors read with a stencil; box = one plane of a 32 x 8 tile + halo
    alignas(64) CUtensorMap tm_rf, tm_w, tm_g, tm_z, tm_d0, tm_d1;
    int smem_ring_off;            // byte offset of the TMA ring inside the dynamic shared memory (behind the coarsest level's arrays)
    double *part;                 // 3 x nparts partial sums
    double phi0, Te0, n0;
    int max_it, nr_max_it;
    double tol, nr_tol;
    double eta0, eta_max, gamma, eta_pow;  // forcing: eta_0, then min(eta_max, gamma (|R_k|/|R_k-1|)^eta_pow)
    double *out;                  // see MGN_OUT_*
    unsigned long long *prof;     // optional: nanoseconds per phase as seen by block 0 (ESPIC_MG_PROFILE=1), 16 slots
};
enum { MGN_OUT_CONVERGED = 0, MGN_OUT_NEWTON = 1, MGN_OUT_LIN = 2, MGN_OUT_YNORM = 3, MGN_OUT_RNORM = 4, MGN_OUT_RNORM0 = 5,
       MGN_OUT_LINFAIL = 6, MGN_OUT_LINL2 = 7, MGN_OUT_ABORT = 8, MGN_OUT_HIST = 9 /* then (|R|, its) per Newton step, up to 10 */ };

__device__ __forceinline__ unsigned long long mg_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// phase ids: 0 down0, 1 coarser down passes, 2 coarsest, 3 coarse up passes, 4 up0 + r.z, 5 d pass, 6 r pass, 7 linearise,
// 8 Galerkin diagonals, 9 update
#define MG_TICK(id) do { if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long n_ = mg_now(); a.prof[id] += n_ - tick; tick = n_; } } while (0)

// ---- fine-level passes: marching along k through a TMA-fed shared-memory ring ----------------------------------------
// The owned planes of the fine level are cut into 32 x 8 tiles of columns; the work list is (tile, plane unit) with the
// plane unit fastest, and block b takes the b-th of gridDim equal contiguous pieces of it: a run of consecutive planes of
// one tile (possibly continuing in the next tile).  A plane unit is 2^fk planes of level 1, so a thread finishes whole
// aggregates.  Lane layout inside a tile: a warp covers 16 (i) x 2 (j) columns -> the x partner of an aggregate is lane^1,
// the y partner lane^16.
//
// Every vector a pass reads with a stencil comes through the TMA unit: one cp.async.bulk.tensor.3d per vector and plane
// copies the tile plus its one-node halo (box 36 x 10 x 1 doubles, 40 x 10 x 1 floats; out-of-mesh coordinates are zero
// filled, which is exactly the value a vector has outside the unknowns) into slot (plane mod MG_RING) of a ring in shared
// memory and signals that slot's mbarrier.  One thread keeps the ring MG_RING planes ahead of the plane being computed, so
// the L2/HBM latency of a plane is hidden behind the arithmetic of the planes before it, every value crosses L2 -> SM once
// per pass, and the threads issue no global loads and no address arithmetic for the stencil at all (the first version of
// these passes loaded straight from global memory and was bound by one memory round trip per plane: 131 us per CG
// iteration for the four fine passes at 128^3, profiles/r2_mg_newton_history.txt).

// TMA wants the first byte of a box 16-byte aligned in global memory: a double tile starts 2 columns left of the tile (one
// halo column + one unused), a float tile 4 columns left (scripts/probes/tma_probe.cu: an odd start column is an illegal
// instruction).
#define MG_RING 5
#define MG_TD_W 36                      // doubles per tile row: columns tx0-2 .. tx0+33
#define MG_TD_X0 2
#define MG_TD_SLOT 368                  // doubles per ring slot: 36 x 10 = 360, rounded up to a multiple of 128 bytes
#define MG_TF_W 40                      // floats per tile row: columns tx0-4 .. tx0+35
#define MG_TF_X0 4
#define MG_TF_SLOT 416                  // floats per ring slot: 40 x 10 = 400, rounded up to a multiple of 128 bytes
#define MG_TD_BYTES (MG_TD_W * (MG_TY + 2) * 8)
#define MG_TF_BYTES (MG_TF_W * (MG_TY + 2) * 4)
#define MG_RING_BYTES (MG_RING * (MG_TD_SLOT * 8 + 2 * MG_TF_SLOT * 4) + MG_RING * 8)

__device__ __forceinline__ unsigned mg_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

struct MgRing {
    double *sd;                  // MG_RING slots of MG_TD_SLOT doubles
    mgf *sf, *sg;                // two float rings of MG_RING slots of MG_TF_SLOT floats
    unsigned long long *bar;     // one mbarrier per slot
    unsigned g;                  // planes staged so far by this block (identical in all its threads): slot = g % MG_RING, phase = g / MG_RING
};

__device__ __forceinline__ void mg_ring_init(MgRing &ring, unsigned char *base)
{
    ring.sd = reinterpret_cast<double *>(base);
    ring.sf = reinterpret_cast<mgf *>(base + MG_RING * MG_TD_SLOT * 8);
    ring.sg = ring.sf + MG_RING * MG_TF_SLOT;
    ring.bar = reinterpret_cast<unsigned long long *>(ring.sg + MG_RING * MG_TF_SLOT);
    ring.g = 0;
    if (threadIdx.x == 0) {
        for (int q = 0; q < MG_RING; q++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mg_smem_u32(ring.bar + q)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

__device__ __forceinline__ void mg_tma_3d(void *dst, const CUtensorMap *map, unsigned long long *bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(mg_smem_u32(dst)), "l"((unsigned long long)map), "r"(mg_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// thread 0: stage plane k of up to three vectors (tile origin tx0, ty0) as ring entry g.  Lane 0 of the ring (double-sized
// slots) takes either a double tile (md) or a float tile (mfd), lanes 1 and 2 float tiles (mf, mg).
__device__ __forceinline__ void mg_ring_issue(const MgRing &ring, unsigned g, const CUtensorMap *md, const CUtensorMap *mf,
                                              const CUtensorMap *mg, const CUtensorMap *mfd, int tx0, int ty0, int k)
{
    const unsigned slot = g % MG_RING;
    unsigned long long *bar = ring.bar + slot;
    const unsigned bytes = (md ? MG_TD_BYTES : 0) + (mf ? MG_TF_BYTES : 0) + (mg ? MG_TF_BYTES : 0) + (mfd ? MG_TF_BYTES : 0);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mg_smem_u32(bar)), "r"(bytes) : "memory");
    if (md) mg_tma_3d(ring.sd + slot * MG_TD_SLOT, md, bar, tx0 - MG_TD_X0, ty0 - 1, k);
    if (mfd) mg_tma_3d(ring.sd + slot * MG_TD_SLOT, mfd, bar, tx0 - MG_TF_X0, ty0 - 1, k);
    if (mf) mg_tma_3d(ring.sf + slot * MG_TF_SLOT, mf, bar, tx0 - MG_TF_X0, ty0 - 1, k);
    if (mg) mg_tma_3d(ring.sg + slot * MG_TF_SLOT, mg, bar, tx0 - MG_TF_X0, ty0 - 1, k);
}

__device__ __forceinline__ void mg_ring_wait(const MgRing &ring, unsigned g)
{
    const unsigned addr = mg_smem_u32(ring.bar + g % MG_RING), parity = (g / MG_RING) & 1u;
    unsigned done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}

struct MgCol {
    int i, j, tx0, ty0;
    bool inmesh;      // the column exists
    bool inner;       // 1 <= i <= ni-2 and 1 <= j <= nj-2: REG nodes only live here, every x/y neighbour exists
    long long nbase;  // j*ni + i      (native layout: phi, rho, type)
    long long pbase;  // j*pitch + i   (padded layout: the solver's own vectors)
    int cd, cf;       // this column's centre inside a double / float tile plane
};

struct MgRuns {
    int ntx, nty, kunit, klo, khi;
    long long nku, first, last;      // plane units per tile; this block's piece [first, last) of the flattened (tile, unit) list
};

template <class Own>
__device__ __forceinline__ MgRuns mg_runs(const Own &own, const MgnArgs &a)
{
    MgRuns R;
    R.ntx = (a.s.ni + MG_TX - 1) / MG_TX;
    R.nty = (a.s.nj + MG_TY - 1) / MG_TY;
    R.kunit = a.nlev > 1 ? (1 << a.L[1].fk) : 1;
    R.klo = own.klo(a.L[0], 0);
    R.khi = own.khi(a.L[0], 0);
    R.nku = (R.khi - R.klo + R.kunit - 1) / R.kunit;
    const long long total = (long long)R.ntx * R.nty * R.nku;
    R.first = total * blockIdx.x / gridDim.x;
    R.last = total * (blockIdx.x + 1) / gridDim.x;
    return R;
}

__device__ __forceinline__ MgCol mg_col(const MgnArgs &a, const MgRuns &R, long long tile)
{
    const StencilC &s = a.s;
    const int tx = (int)(tile % R.ntx), ty = (int)(tile / R.ntx);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int li = (w & 1) * 16 + (l & 15), lj = (w >> 1) * 2 + (l >> 4);
    MgCol c;
    c.tx0 = tx * MG_TX; c.ty0 = ty * MG_TY;
    c.i = c.tx0 + li;
    c.j = c.ty0 + lj;
    c.inmesh = c.i < s.ni && c.j < s.nj;
    c.inner = c.i >= 1 && c.i <= s.ni - 2 && c.j >= 1 && c.j <= s.nj - 2;
    c.nbase = (long long)c.j * s.sj + c.i;
    c.pbase = (long long)c.j * a.pitch + c.i;
    c.cd = (lj + 1) * MG_TD_W + li + MG_TD_X0;
    c.cf = (lj + 1) * MG_TF_W + li + MG_TF_X0;
    return c;
}

// calls f(column, kbeg, kend) for every run of this block
template <typename F>
__device__ __forceinline__ void mg_for_runs(const MgnArgs &a, const MgRuns &R, F f)
{
    for (long long pos = R.first; pos < R.last;) {
        const long long tile = pos / R.nku, ku = pos % R.nku;
        const long long nrun = min(R.last - pos, R.nku - ku);
        const int kbeg = R.klo + (int)ku * R.kunit, kend = min(R.khi, kbeg + (int)nrun * R.kunit);
        f(mg_col(a, R, tile), kbeg, kend);
        pos += nrun;
    }
}

// One run through the ring: planes kbeg-1 .. kend of the given vectors are staged (ring entries g0 .. g0 + nplanes - 1), and
// body(k, m, c, p) is called for k = kbeg .. kend-1 with the slot numbers holding planes k-1, k, k+1.
template <typename F>
__device__ __forceinline__ void mg_ring_run(MgRing &ring, const CUtensorMap *md, const CUtensorMap *mf, const CUtensorMap *mg,
                                            const CUtensorMap *mfd, const MgCol &col, int kbeg, int kend, F body)
{
    const int nplanes = kend - kbeg + 2;
    const unsigned g0 = ring.g;
    if (threadIdx.x == 0) {
        // the slots were last read through the generic proxy, and the vectors were last written by generic stores of other
        // blocks (ordered by the grid barrier before this pass): order both before the async-proxy copies
        asm volatile("fence.proxy.async;" ::: "memory");
        for (int q = 0; q < min(nplanes, MG_RING); q++) mg_ring_issue(ring, g0 + q, md, mf, mg, mfd, col.tx0, col.ty0, kbeg - 1 + q);
    }
    mg_ring_wait(ring, g0);
    mg_ring_wait(ring, g0 + 1);
    for (int k = kbeg; k < kend; k++) {
        const unsigned t = (unsigned)(k - kbeg);
        mg_ring_wait(ring, g0 + t + 2);
        body(k, (int)((g0 + t) % MG_RING), (int)((g0 + t + 1) % MG_RING), (int)((g0 + t + 2) % MG_RING));
        __syncthreads();                      // everybody is done with plane k-1: its slot takes the plane MG_RING further on
        // (reads through the generic proxy followed by an async-proxy write of the same slot need no proxy fence: the block
        // barrier orders them; a fence per plane cost more than the plane's arithmetic, ncu: membar + barrier stalls)
        if (threadIdx.x == 0 && (int)t + MG_RING < nplanes)
            mg_ring_issue(ring, g0 + t + MG_RING, md, mf, mg, mfd, col.tx0, col.ty0, kbeg - 1 + (int)t + MG_RING);
    }
    ring.g = g0 + (unsigned)nplanes;
}

// Passes A and B are the fine level of the V-cycle: FP32 throughout (see mg_down).  With x0 = w D^-1 r the pre-smoothed
// iterate, D x0 = w r, so neither pass needs the diagonal itself:
//   residual of x0:          r - K x0 = (1 - w) r + offdiag(x0)
//   post-smoothing of xu:    xu + w D^-1 (r - K xu) = (1 - w) xu + (w D^-1)(r + offdiag(xu))
// with offdiag(v) = sum over the six neighbours of g v.  Staged vectors: rf (FP32 copy of r) and winv = w / diag.

// Pass A (down, fine -> level 1): residual of the pre-smoothed iterate x0, summed over each aggregate -> b of level 1
template <class Own>
__device__ __forceinline__ void mg_fine_down(const Own &own, const MgnArgs &a, const MgRuns &R, MgRing &ring, bool to_all)
{
    const StencilC &s = a.s;
    const MgLevel &C = a.L[1];
    const int lane = threadIdx.x & 31;
    const mgf gx = (mgf)s.gdx2, gy = (mgf)s.gdy2, gz = (mgf)s.gdz2, W1 = (mgf)(1.0 - MG_OMEGA);
    const mgf *ringr = reinterpret_cast<const mgf *>(ring.sd);          // the double-sized slots carry the FP32 tile of rf
    mg_for_runs(a, R, [&](const MgCol col, const int kbeg, const int kend) {
        const bool writer = col.inmesh && (!C.fi || !(lane & 1)) && (!C.fj || !(lane & 16));
        const long long crow = (long long)(col.j >> C.fj) * C.ni + (col.i >> C.fi), cplane = (long long)C.ni * C.nj;
        mgf sum = 0;
        mg_ring_run(ring, nullptr, &a.tm_w, nullptr, &a.tm_rf, col, kbeg, kend, [&](int k, int m, int c, int p) {
            const mgf *rm = ringr + m * (2 * MG_TD_SLOT) + col.cf, *rc = ringr + c * (2 * MG_TD_SLOT) + col.cf, *rp = ringr + p * (2 * MG_TD_SLOT) + col.cf;
            const mgf *wm = ring.sf + m * MG_TF_SLOT + col.cf, *wc = ring.sf + c * MG_TF_SLOT + col.cf, *wp = ring.sf + p * MG_TF_SLOT + col.cf;
            const mgf off = gx * (rc[-1] * wc[-1] + rc[1] * wc[1]) + gy * (rc[-MG_TF_W] * wc[-MG_TF_W] + rc[MG_TF_W] * wc[MG_TF_W]) +
                            gz * (rm[0] * wm[0] + rp[0] * wp[0]);
            sum += wc[0] != (mgf)0 ? W1 * rc[0] + off : (mgf)0;
            if (((k + 1) & (R.kunit - 1)) == 0 || k + 1 == s.nk) {        // last plane of an aggregate: combine the children, store
                mgf t = sum;
                if (C.fi) t += __shfl_xor_sync(0xffffffffu, t, 1);
                if (C.fj) t += __shfl_xor_sync(0xffffffffu, t, 16);
                if (writer) {
                    const long long I = (long long)(k >> C.fk) * cplane + crow;
                    if (to_all) own.st_all(C.b, I, t);
                    else own.st(C.b, I, C, 1, t);
                }
                sum = 0;
            }
        });
    });
}

// Pass B (up, level 1 -> fine): z = post-smoothing of x0 + P e; returns the thread's share of r.z
template <class Own>
__device__ __forceinline__ double mg_fine_up(const Own &own, const MgnArgs &a, const MgRuns &R, MgRing &ring, const mgf *e)
{
    const StencilC &s = a.s;
    const MgLevel &C = a.L[1];
    const mgf gx = (mgf)s.gdx2, gy = (mgf)s.gdy2, gz = (mgf)s.gdz2, W1 = (mgf)(1.0 - MG_OMEGA);
    const mgf *ringr = reinterpret_cast<const mgf *>(ring.sd);
    double acc = 0;
    mg_for_runs(a, R, [&](const MgCol col, const int kbeg, const int kend) {
        // the correction of a neighbour is that of ITS aggregate: offsets of the neighbours' aggregates relative to this one's
        const int cplane = C.ni * C.nj;
        const int oxm = ((col.i - 1) >> C.fi) - (col.i >> C.fi), oxp = ((col.i + 1) >> C.fi) - (col.i >> C.fi);
        const int oym = (((col.j - 1) >> C.fj) - (col.j >> C.fj)) * C.ni, oyp = (((col.j + 1) >> C.fj) - (col.j >> C.fj)) * C.ni;
        const long long crow = (long long)(col.j >> C.fj) * C.ni + (col.i >> C.fi);
        long long up = col.pbase + (long long)kbeg * a.psk;
        mg_ring_run(ring, nullptr, &a.tm_w, nullptr, &a.tm_rf, col, kbeg, kend, [&](int k, int m, int c, int p) {
            // corrections first: these are the only global loads of the step, in flight while the tile is read
            mgf e0 = 0, exm = 0, exp_ = 0, eym = 0, eyp = 0, ezm = 0, ezp = 0;
            if (col.inner) {
                const mgf *ep = e + ((long long)(k >> C.fk) * cplane + crow);
                const int ozm = k >= 1 ? (((k - 1) >> C.fk) - (k >> C.fk)) * cplane : 0;
                const int ozp = k + 1 < s.nk ? (((k + 1) >> C.fk) - (k >> C.fk)) * cplane : 0;
                e0 = ep[0]; exm = ep[oxm]; exp_ = ep[oxp]; eym = ep[oym]; eyp = ep[oyp]; ezm = ep[ozm]; ezp = ep[ozp];
            }
            const mgf *rm = ringr + m * (2 * MG_TD_SLOT) + col.cf, *rc = ringr + c * (2 * MG_TD_SLOT) + col.cf, *rp = ringr + p * (2 * MG_TD_SLOT) + col.cf;
            const mgf *wm = ring.sf + m * MG_TF_SLOT + col.cf, *wc = ring.sf + c * MG_TF_SLOT + col.cf, *wp = ring.sf + p * MG_TF_SLOT + col.cf;
            // a neighbour that is not an unknown (winv = 0) carries 0, whatever its aggregate's correction is
            const mgf w_xm = wc[-1], w_xp = wc[1], w_ym = wc[-MG_TF_W], w_yp = wc[MG_TF_W], w_zm = wm[0], w_zp = wp[0];
            const mgf vxm = w_xm != (mgf)0 ? rc[-1] * w_xm + exm : (mgf)0, vxp = w_xp != (mgf)0 ? rc[1] * w_xp + exp_ : (mgf)0;
            const mgf vym = w_ym != (mgf)0 ? rc[-MG_TF_W] * w_ym + eym : (mgf)0, vyp = w_yp != (mgf)0 ? rc[MG_TF_W] * w_yp + eyp : (mgf)0;
            const mgf vzm = w_zm != (mgf)0 ? rm[0] * w_zm + ezm : (mgf)0, vzp = w_zp != (mgf)0 ? rp[0] * w_zp + ezp : (mgf)0;
            const mgf off = gx * (vxm + vxp) + gy * (vym + vyp) + gz * (vzm + vzp);
            const mgf r0 = rc[0], w0 = wc[0];
            const mgf xu = r0 * w0 + e0;
            const mgf zf = w0 != (mgf)0 ? W1 * xu + w0 * (r0 + off) : (mgf)0;
            acc += (double)r0 * (double)zf;
            if (col.inmesh) own.st_k(a.z, up, k, 0, zf);
            up += a.psk;
        });
    });
    return acc;
}

// Pass C: d = z + beta d_old (formed on the fly for the neighbours, written for this node); returns the share of d.K d
template <class Own>
__device__ __forceinline__ double mg_fine_dir(const Own &own, const MgnArgs &a, const MgRuns &R, MgRing &ring, double beta, bool d_is_d0)
{
    const StencilC &s = a.s;
    const CUtensorMap *md = d_is_d0 ? &a.tm_d0 : &a.tm_d1;       // d_old
    double *d_new = d_is_d0 ? a.d1 : a.d0;
    double acc = 0;
    mg_for_runs(a, R, [&](const MgCol col, const int kbeg, const int kend) {
        long long up = col.pbase + (long long)kbeg * a.psk;
        mg_ring_run(ring, md, &a.tm_z, &a.tm_g, nullptr, col, kbeg, kend, [&](int k, int m, int c, int p) {
            const double *dm = ring.sd + m * MG_TD_SLOT + col.cd, *dc = ring.sd + c * MG_TD_SLOT + col.cd, *dp = ring.sd + p * MG_TD_SLOT + col.cd;
            const mgf *zm = ring.sf + m * MG_TF_SLOT + col.cf, *zc = ring.sf + c * MG_TF_SLOT + col.cf, *zp = ring.sf + p * MG_TF_SLOT + col.cf;
            const double dg = ring.sg[c * MG_TF_SLOT + col.cf];
            // z and d are identically zero outside the REG set: no neighbour masks
            const double n_c = (double)zc[0] + beta * dc[0];
            const double n_xm = (double)zc[-1] + beta * dc[-1], n_xp = (double)zc[1] + beta * dc[1];
            const double n_ym = (double)zc[-MG_TF_W] + beta * dc[-MG_TD_W], n_yp = (double)zc[MG_TF_W] + beta * dc[MG_TD_W];
            const double n_zm = (double)zm[0] + beta * dm[0], n_zp = (double)zp[0] + beta * dp[0];
            const double off = s.gdx2 * (n_xm + n_xp) + s.gdy2 * (n_ym + n_yp) + s.gdz2 * (n_zm + n_zp);
            const double q = n_c * (dg * n_c - off);
            acc += dg != 0 ? q : 0.0;
            if (col.inmesh) own.st_k(d_new, up, k, 0, n_c);
            up += a.psk;
        });
    });
    return acc;
}

// Pass D: delta += alpha d ; r -= alpha K d ; returns the share of |r|^2
template <class Own>
__device__ __forceinline__ double mg_fine_res(const Own &own, const MgnArgs &a, const MgRuns &R, MgRing &ring, double alpha, bool d_is_d0)
{
    const StencilC &s = a.s;
    const CUtensorMap *md = d_is_d0 ? &a.tm_d0 : &a.tm_d1;
    double acc = 0;
    mg_for_runs(a, R, [&](const MgCol col, const int kbeg, const int kend) {
        long long up = col.pbase + (long long)kbeg * a.psk;
        // delta and r are read at the node only: straight from global memory, one plane ahead of their use
        double del_n = 0, r_n = 0;
        if (col.inmesh) { del_n = a.delta[up]; r_n = a.r[up]; }
        mg_ring_run(ring, md, &a.tm_g, nullptr, nullptr, col, kbeg, kend, [&](int k, int m, int c, int p) {
            const double del = del_n, rr = r_n;
            if (col.inmesh && k + 1 < kend) { del_n = a.delta[up + a.psk]; r_n = a.r[up + a.psk]; }
            const double *dm = ring.sd + m * MG_TD_SLOT + col.cd, *dc = ring.sd + c * MG_TD_SLOT + col.cd, *dp = ring.sd + p * MG_TD_SLOT + col.cd;
            const double dg = ring.sf[c * MG_TF_SLOT + col.cf];
            const double d0 = dc[0];
            const double q = dg * d0 - (s.gdx2 * (dc[-1] + dc[1]) + s.gdy2 * (dc[-MG_TD_W] + dc[MG_TD_W]) + s.gdz2 * (dm[0] + dp[0]));
            if (dg != 0) {
                a.delta[up] = del + alpha * d0;
                const double rn = rr - alpha * q;
                a.r[up] = rn;                                   // r is read at the node only; the V-cycle reads its FP32 copy
                own.st_k(a.rf, up, k, 0, (mgf)rn);
                acc += rn * rn;
            }
            up += a.psk;
        });
    });
    return acc;
}

// Newton linearisation at the current phi (the reference's GS residual with the Neumann face neighbours folded into the
// node itself, PotentialSolver.cpp:389-421), Jacobian diagonal; delta = 0, d0 = 0.  Returns the share of |R|^2.
// phi, rho and the node types are the caller's arrays (native layout), r / diagf / d0 / delta the solver's (padded rows).
template <class Own>
__device__ __forceinline__ double mg_linearise(const Own &own, const MgnArgs &a, const MgRuns &R)
{
    const StencilC &s = a.s;
    const double *phi = a.phi;
    double acc = 0;
    mg_for_runs(a, R, [&](const MgCol col, const int kbeg, const int kend) {
        if (!col.inmesh) return;
        long long u = col.nbase + (long long)kbeg * s.sk, up = col.pbase + (long long)kbeg * a.psk;
        for (int k = kbeg; k < kend; k++, u += s.sk, up += a.psk) {
            double r = 0;
            mgf dj = 0, wi = 0;
            // REG nodes are interior: on an inner column away from the first and last plane every neighbour address is valid
            // whatever the node is, so all loads are issued at once and the node type only selects
            if (col.inner && k >= 1 && k + 1 < s.nk) {
                const int ty = a.type[u];
                const bool fxm = a.type[u - 1] >= NT_I0, fxp = a.type[u + 1] >= NT_I0, fym = a.type[u - s.sj] >= NT_I0,
                           fyp = a.type[u + s.sj] >= NT_I0, fzm = a.type[u - s.sk] >= NT_I0, fzp = a.type[u + s.sk] >= NT_I0;
                const double p = phi[u], rho = a.rho[u];
                const double pxm = phi[u - 1], pxp = phi[u + 1], pym = phi[u - s.sj], pyp = phi[u + s.sj], pzm = phi[u - s.sk], pzp = phi[u + s.sk];
                if (ty == NT_REG) {
                    const double ex = exp((p - a.phi0) / a.Te0);
                    const double src = (rho - C_QE * (a.n0 * ex)) / C_EPS_0;
                    const double xm = fxm ? p : pxm, xp = fxp ? p : pxp;
                    const double ym = fym ? p : pym, yp = fyp ? p : pyp;
                    const double zm = fzm ? p : pzm, zp = fzp ? p : pzp;
                    r = -p * (2 * s.gdx2 + 2 * s.gdy2 + 2 * s.gdz2) + src + s.gdx2 * (xm + xp) + s.gdy2 * (ym + yp) + s.gdz2 * (zm + zp);
                    double d0 = 2 * s.gdx2 + 2 * s.gdy2 + 2 * s.gdz2;
                    if (fxm) d0 -= s.gdx2;
                    if (fxp) d0 -= s.gdx2;
                    if (fym) d0 -= s.gdy2;
                    if (fyp) d0 -= s.gdy2;
                    if (fzm) d0 -= s.gdz2;
                    if (fzp) d0 -= s.gdz2;
                    const double djd = d0 + a.n0 * C_QE / (C_EPS_0 * a.Te0) * ex;
                    dj = (mgf)djd;
                    wi = (mgf)(MG_OMEGA / djd);
                    acc += r * r;
                }
            }
            a.r[up] = r;
            own.st_k(a.rf, up, k, 0, (mgf)r);
            own.st_k(a.winv, up, k, 0, wi);
            a.diagf[up] = dj;                                   // read at the node only (passes C and D, the coarse diagonals)
            own.st_k(a.d0, up, k, 0, 0.0);
            a.delta[up] = 0;
        }
    });
    return acc;
}

// phi += delta on the unknowns; sum of delta^2 counted once per node that takes the value (the node + the face nodes
// mirroring it), i.e. the reference's sum over all nodes of y^2 (PotentialSolver.cpp:279-283)
template <class Own>
__device__ __forceinline__ double mg_update(const Own &own, const MgnArgs &a, const MgRuns &R)
{
    const StencilC &s = a.s;
    double acc = 0;
    mg_for_runs(a, R, [&](const MgCol col, const int kbeg, const int kend) {
        if (!col.inner) return;                                   // REG nodes are interior
        long long u = col.nbase + (long long)kbeg * s.sk, up = col.pbase + (long long)kbeg * a.psk;
        for (int k = kbeg; k < kend; k++, u += s.sk, up += a.psk) {
            if (k < 1 || k + 1 >= s.nk) continue;
            const int ty = a.type[u];
            const int cnt = 1 + (a.type[u - 1] >= NT_I0) + (a.type[u + 1] >= NT_I0) + (a.type[u - s.sj] >= NT_I0) + (a.type[u + s.sj] >= NT_I0) +
                            (a.type[u - s.sk] >= NT_I0) + (a.type[u + s.sk] >= NT_I0);
            const double dl = a.delta[up], p = a.phi[u];
            if (ty != NT_REG) continue;
            own.st_k(a.phi, u, k, 0, p + dl);
            acc += cnt * (dl * dl);
        }
    });
    return acc;
}

// Sum of per-block partials (of every rank in slab mode), computed redundantly by every block in the same fixed order
template <class Own>
__device__ __forceinline__ double mg_total(const Own &own, const double *part, int nb, double *sh, double *bcast)
{
    return grid_total(part, own.nparts(nb), sh, bcast);
}

// z = M^-1 r (one V-cycle); returns this thread's share of r.z.  Ends WITHOUT a barrier.  ok = false: a slab barrier gave up.
template <class Own>
__device__ __forceinline__ double mg_vcycle(cg::grid_group &grid, Own &own, const MgnArgs &a, const MgRuns &R, MgRing &ring, long long t0,
                                            long long stride, mgf *smem, unsigned long long &tick, bool &ok)
{
        if (a.nlev == 1) {           // degenerate hierarchy (tiny mesh): plain Jacobi preconditioner
        double acc = 0;
        mg_for_runs(a, R, [&](const MgCol col, const int kbeg, const int kend) {
            if (!col.inmesh) return;
            long long up = col.pbase + (long long)kbeg * a.psk;
            for (int k = kbeg; k < kend; k++, up += a.psk) {
                const mgf dg = a.diagf[up];
                const double rr = a.r[up];
                const mgf zf = dg != (mgf)0 ? (mgf)(rr * (double)__frcp_rn(dg)) : (mgf)0;
                own.st_k(a.z, up, k, 0, zf);
                acc += rr * (double)zf;
            }
        });
        return acc;
    }
    // The coarse levels are tiny: in slab mode the right-hand side of level `lr` (a.first_redundant; at the latest the
    // coarsest level) is stored to EVERY rank, and every rank runs the levels lr .. coarsest redundantly in full with
    // grid-local barriers only (identical arithmetic -> identical result everywhere).
    const int lc = a.nlev - 1;
    int lr = lc;
    if constexpr (Own::slab) lr = a.first_redundant;
    OwnAll whole;
    // a pass whose output is only read across the slab boundary (halo planes) synchronises with the two neighbours; one that
    // stores to every rank (the right-hand side of the first redundant level) or feeds a dot product needs all ranks
    mg_fine_down(own, a, R, ring, lr == 1);
    ok = (lr == 1 ? own.barrier(grid) : own.nbarrier(grid)) && ok;
    MG_TICK(0);
    for (int l = 1; l + 1 < a.nlev; l++) {
        if (Own::slab && l >= lr) {
            mg_down(whole, a.L[l], l, a.L[l + 1], t0, stride, false);
            ok = own.lbarrier(grid) && ok;
        } else {
            mg_down(own, a.L[l], l, a.L[l + 1], t0, stride, l + 1 == lr);
            ok = (l + 1 == lr ? own.barrier(grid) : own.nbarrier(grid)) && ok;
        }
    }
    MG_TICK(1);
    const mgf *e = mg_coarsest(a.L[lc], a.coarse_sweeps, smem);
    MG_TICK(2);
    for (int l = a.nlev - 2; l >= 1; l--) {
        if (Own::slab && l >= lr) {
            mg_up(whole, a.L[l], l, a.L[l + 1], e, t0, stride);
            ok = own.lbarrier(grid) && ok;
        } else {
            mg_up(own, a.L[l], l, a.L[l + 1], e, t0, stride);
            ok = own.nbarrier(grid) && ok;
        }
        e = a.L[l].xn;
    }
    MG_TICK(3);
    return mg_fine_up(own, a, R, ring, e);
}

template <class Own>
__device__ __forceinline__ void mgn_body(const MgnArgs &a, Own &own, unsigned char *smem_raw)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[32];
    __shared__ double bc;
    const StencilC &s = a.s;
    const MgLevel &L0 = a.L[0];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int nb = gridDim.x;
    const int np = own.nparts(nb);
    double *pA = a.part, *pB = a.part + np, *pC = a.part + 2 * np;
    const MgRuns R = mg_runs(own, a);
    const double nn = (double)s.nn;
    mgf *smem = reinterpret_cast<mgf *>(smem_raw);                // the coarsest level's arrays, then the TMA ring
    MgRing ring;
    mg_ring_init(ring, smem_raw + a.smem_ring_off);
    bool ok = true;
    unsigned long long tick = mg_now();

    if constexpr (Own::slab) {
        // working copy of phi inside the pool: this rank's planes plus one halo plane on each side (kept current afterwards
        // by the peer stores of the update pass)
        const long long plane = s.sk;
        const long long lo = max(0ll, (long long)own.klo(L0, 0) * plane - plane), hi = min(s.nn, (long long)own.khi(L0, 0) * plane + plane);
        for (long long u = lo + t0; u < hi; u += stride) a.phi[u] = a.phi_user[u];
        ok = own.lbarrier(grid) && ok;
    }

    int converged = 0, nit = 0, lin_fail = 0;
    bool last_full = false;       // the previous linear solve went all the way to tol/2 (not stopped early by the forcing term)
    long long lin_total = 0;
    double ynorm = 0, Rn = 0, Rprev = 0, R0 = 0, l2 = 0;
    for (nit = 0;; nit++) {
        // ---- linearise at the current phi
        double acc = mg_linearise(own, a, R);
        double t = block_sum(acc, sh);
        if (threadIdx.x == 0) own.put_partial(pC, nb, t);
        ok = own.barrier(grid) && ok;
        Rn = sqrt(mg_total(own, pC, nb, sh, &bc) / nn);
        MG_TICK(7);
        if (nit == 0) R0 = Rn;
        if (blockIdx.x == 0 && threadIdx.x == 0 && nit < 10) a.out[MGN_OUT_HIST + 2 * nit] = Rn;
        if (!ok) break;
        // Converged: the reference's test on the update (PotentialSolver.cpp:282-285) AND the nonlinear residual below tol.  Near
        // the rounding floor of the residual evaluation (|phi| diag eps, reachable only with tolerances far below production)
        // |R| stops falling: there a full-accuracy linear solve whose update passed the reference's test is accepted, as the
        // reference itself would.
        if (nit == 0 ? Rn < a.tol : (ynorm < a.nr_tol && (Rn < a.tol || last_full))) { converged = 1; break; }
        if (nit >= a.nr_max_it) break;
        // ---- Galerkin diagonals of the coarse levels.  Only the mass term e^(phi/Te) of the diagonal depends on phi; once a Newton
        // update is small against Te the coarse operators of the previous step are as good a preconditioner (same CG iteration
        // counts on the bench case, scripts/mg_variants_prototype.py) and their rebuild -- one barrier per level -- is skipped.
        // The fine-level diagonal, which defines the operator itself, is always current.  ynorm is the same number in every
        // block and on every rank, so the branch is uniform.
        const bool rebuild = nit == 0 || (a.n0 != 0 && ynorm > 0.1 * fabs(a.Te0));
        if (rebuild) {
            int lr = a.nlev - 1;
            if constexpr (Own::slab) lr = a.first_redundant;
            OwnAll whole;
            for (int l = 1; l < a.nlev; l++) {
                if (Own::slab && l > lr) {
                    mg_diagl(whole, a.L[l - 1], a.L[l], l, t0, stride, false);
                    ok = own.lbarrier(grid) && ok;
                } else {
                    const bool to_all = Own::slab && l == lr;
                    if (l == 1) mg_diag1(own, s, a.type, a.diagf, a.pitch, a.L[1], t0, stride, to_all);
                    else mg_diagl(own, a.L[l - 1], a.L[l], l, t0, stride, to_all);
                    ok = (to_all ? own.barrier(grid) : own.nbarrier(grid)) && ok;
                }
            }
        }
        if (rebuild && a.nlev > 1) mg_coarsest_stage(a.L[a.nlev - 1], smem);
        MG_TICK(8);
        // ---- CG on K delta = R down to the forcing level
        const double eta = nit == 0 ? a.eta0 : fmin(a.eta_max, a.gamma * pow(Rn / Rprev, a.eta_pow));
        const double stop = fmax(0.5 * a.tol, eta * Rn);
        Rprev = Rn;
        l2 = Rn;
        int it = 0;
        double rz = 0;
        bool d_is_d0 = true;          // which of d0 / d1 holds the previous search direction
        while (ok && l2 >= stop && it < a.max_it) {
            // z = M^-1 r ; rz' = r.z
            acc = mg_vcycle(grid, own, a, R, ring, t0, stride, smem, tick, ok);
            t = block_sum(acc, sh);
            if (threadIdx.x == 0) own.put_partial(pB, nb, t);
            ok = own.barrier(grid) && ok;
            MG_TICK(4);
            const double rz_new = mg_total(own, pB, nb, sh, &bc);
            const double beta = (it == 0) ? 0.0 : rz_new / rz;
            rz = rz_new;
            // d = z + beta d ; dq = d.K d
            acc = mg_fine_dir(own, a, R, ring, beta, d_is_d0);
            t = block_sum(acc, sh);
            if (threadIdx.x == 0) own.put_partial(pA, nb, t);
            ok = own.barrier(grid) && ok;
            MG_TICK(5);
            const double dq = mg_total(own, pA, nb, sh, &bc);
            if (!(dq > 0)) { lin_fail = 1; break; }          // breakdown (cannot happen for an SPD system short of overflow): give up cleanly
            const double alpha = rz / dq;
            // delta += alpha d ; r -= alpha K d ; |r|
            acc = mg_fine_res(own, a, R, ring, alpha, !d_is_d0);
            t = block_sum(acc, sh);
            if (threadIdx.x == 0) own.put_partial(pC, nb, t);
            ok = own.barrier(grid) && ok;
            MG_TICK(6);
            l2 = sqrt(mg_total(own, pC, nb, sh, &bc) / nn);
            it++;
            d_is_d0 = !d_is_d0;
        }
        if (l2 >= stop) lin_fail = 1;
        last_full = stop <= 0.5 * a.tol && l2 < stop;
        lin_total += it;
        if (blockIdx.x == 0 && threadIdx.x == 0 && nit < 10) a.out[MGN_OUT_HIST + 2 * nit + 1] = it;
        if (!ok) break;
        // ---- phi += delta
        acc = mg_update(own, a, R);
        t = block_sum(acc, sh);
        if (threadIdx.x == 0) own.put_partial(pA, nb, t);
        ok = own.barrier(grid) && ok;
        ynorm = sqrt(mg_total(own, pA, nb, sh, &bc) / nn);
        MG_TICK(9);
    }
    if constexpr (Own::slab) {
        for (long long u = (long long)own.klo(L0, 0) * s.sk + t0; u < (long long)own.khi(L0, 0) * s.sk; u += stride) a.phi_user[u] = a.phi[u];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.out[MGN_OUT_CONVERGED] = converged; a.out[MGN_OUT_NEWTON] = nit; a.out[MGN_OUT_LIN] = (double)lin_total;
        a.out[MGN_OUT_YNORM] = ynorm; a.out[MGN_OUT_RNORM] = Rn; a.out[MGN_OUT_RNORM0] = R0; a.out[MGN_OUT_LINFAIL] = lin_fail;
        a.out[MGN_OUT_LINL2] = l2; a.out[MGN_OUT_ABORT] = ok ? 0.0 : 1.0;
    }
}

extern __shared__ __align__(128) unsigned char mgn_smem[];

__global__ void __launch_bounds__(MG_THREADS, 3) k_mg_newton(const __grid_constant__ MgnArgs a)
{
    OwnAll own;
    mgn_body(a, own, mgn_smem);
}

// slab-decomposed variant: one of these kernels per rank, running concurrently, talking through peer memory only
__global__ void __launch_bounds__(MG_THREADS, 3) k_mg_newton_slab(const __grid_constant__ MgnArgs a, OwnSlab own, unsigned long long *epoch_io)
{
    own.epoch = epoch_io[0];
    own.lepoch = epoch_io[1];
    own.nepoch = epoch_io[2];
    own.lepoch_all = epoch_io[3];
    mgn_body(a, own, mgn_smem);
    // every block read the epochs before its first barrier, and nobody gets past that barrier before all have arrived
    if (blockIdx.x == 0 && threadIdx.x == 0) { epoch_io[0] = own.epoch; epoch_io[1] = own.lepoch; epoch_io[2] = own.nepoch; epoch_io[3] = own.lepoch_all; }
}

// ---- host side -------------------------------------------------------------------------------------------------------

static inline int mg_pitch(int ni) { return (ni + 3) & ~3; }

// number of levels, their dimensions and per-dimension coarsening shifts for an (ni,nj,nk) mesh with spacings dh
static int mg_level_dims(const StencilC &s, long long dims[MG_MAX_LEVELS][3], int shifts[MG_MAX_LEVELS][3])
{
    int n[3] = {s.ni, s.nj, s.nk}, nlev = 1;
    // spacing from the stencil coefficients: g = 1/dh^2
    double h[3] = {1.0 / sqrt(s.gdx2), 1.0 / sqrt(s.gdy2), 1.0 / sqrt(s.gdz2)};
    for (int a = 0; a < 3; a++) { dims[0][a] = n[a]; shifts[0][a] = 0; }
    static const bool semi = getenv("ESPIC_MG_FULL_COARSENING") == nullptr;
    // halve (rounding up) while every dimension stays > 4 and the level is worth a barrier
    while (nlev < MG_MAX_LEVELS && std::min(n[0], std::min(n[1], n[2])) > 4 && (long long)n[0] * n[1] * n[2] > MG_COARSEST_NODES) {
        const double hmin = std::min(h[0], std::min(h[1], h[2]));
        for (int a = 0; a < 3; a++) {
            const bool coarsen = !semi || h[a] <= 1.42 * hmin;
            shifts[nlev][a] = coarsen ? 1 : 0;
            if (coarsen) { n[a] = (n[a] + 1) / 2; h[a] *= 2; }
            dims[nlev][a] = n[a];
        }
        nlev++;
    }
    // the coarsest level is solved inside shared memory (3 FP32 vectors).  A mesh with one short dimension stops coarsening
    // early and can leave a last level that does not fit: such a hierarchy is cut back (in the limit to the plain Jacobi
    // preconditioner of one level).  Never the case for the meshes of this path.
    while (nlev > 1 && dims[nlev - 1][0] * dims[nlev - 1][1] * dims[nlev - 1][2] > 2 * MG_COARSEST_NODES) nlev--;
    return nlev;
}

static long long mg_coarse_elems(const StencilC &s)
{
    long long dims[MG_MAX_LEVELS][3];
    int shifts[MG_MAX_LEVELS][3];
    const int nlev = mg_level_dims(s, dims, shifts);
    long long total = 0;
    for (int l = 1; l < nlev; l++) total += MG_LEVEL_ARRAYS * ((dims[l][0] * dims[l][1] * dims[l][2] + 3) & ~3ll);
    return total;
}

// slab mode: first level that every rank solves in full -- the largest run of trailing levels whose node counts are all
// <= limit (the coarsest level always is; level 0 never)
static int mg_first_redundant(int nlev, const long long dims[MG_MAX_LEVELS][3], long long limit)
{
    int lr = nlev - 1;
    while (lr > 1 && dims[lr - 1][0] * dims[lr - 1][1] * dims[lr - 1][2] <= limit) lr--;
    return lr;
}

static long long mg_slab_redundant_limit()
{
    const char *ev = getenv("ESPIC_MG_SLAB_REDUNDANT_NODES");
    return ev ? atoll(ev) : MG_SLAB_REDUNDANT_NODES;
}

// fine planes per coarsest plane: slab boundaries must be multiples of it so that no aggregate straddles two ranks
static int mg_slab_plane_unit(int nlev, const int shifts[MG_MAX_LEVELS][3])
{
    int kshift = 0;
    for (int l = 1; l < nlev; l++) kshift += shifts[l][2];
    return 1 << kshift;
}

// Host-side planning only (no device work, callable without a GPU): the hierarchy ESPIC_SOLVE_PCG_MG(_SLAB) builds for a mesh.
extern "C" int espic_mg_plan(int ni, int nj, int nk, const double dh[3], int nranks, long long dims_out[8][3], int *first_redundant,
                             int *slab_plane_unit)
{
    if (ni < 2 || nj < 2 || nk < 2 || !dh || !(dh[0] > 0) || !(dh[1] > 0) || !(dh[2] > 0) || nranks < 1) {
        espic_set_error("espic_mg_plan: bad mesh");
        return -1;
    }
    StencilC s;
    memset(&s, 0, sizeof(s));
    s.ni = ni; s.nj = nj; s.nk = nk; s.nn = (long long)ni * nj * nk; s.sj = ni; s.sk = (long long)ni * nj;
    s.gdx2 = 1.0 / (dh[0] * dh[0]); s.gdy2 = 1.0 / (dh[1] * dh[1]); s.gdz2 = 1.0 / (dh[2] * dh[2]);
    long long dims[MG_MAX_LEVELS][3];
    int shifts[MG_MAX_LEVELS][3];
    const int nlev = mg_level_dims(s, dims, shifts);
    for (int l = 0; l < 8; l++)
        for (int a = 0; a < 3; a++)
            if (dims_out) dims_out[l][a] = l < nlev ? dims[l][a] : 0;
    if (first_redundant) *first_redundant = nranks > 1 ? mg_first_redundant(nlev, dims, mg_slab_redundant_limit()) : nlev - 1;
    if (slab_plane_unit) *slab_plane_unit = mg_slab_plane_unit(nlev, shifts);
    return nlev;
}

struct MgKnobs { int coarse_sweeps; double eta0, eta_max, gamma, eta_pow, link_scale; bool profile; };
// shared by the single-GPU and the slab solver, so that the two paths cannot drift apart
static const MgKnobs &mg_knobs()
{
    static MgKnobs k;
    static bool init = false;
    if (!init) {
        const char *ev = getenv("ESPIC_MG_COARSE_SWEEPS");
        k.coarse_sweeps = ev ? std::max(0, atoi(ev)) : MG_COARSE_SWEEPS;
        // ESPIC_MG_EXACT_NEWTON=1: every linear solve goes to tol/2 (the reference's exact Newton); else Eisenstat-Walker
        const bool exact = getenv("ESPIC_MG_EXACT_NEWTON") != nullptr;
        k.eta0 = exact ? 0.0 : (getenv("ESPIC_MG_ETA0") ? atof(getenv("ESPIC_MG_ETA0")) : 1e-2);
        k.eta_max = exact ? 0.0 : 1e-1;
        k.gamma = exact ? 0.0 : 0.9;
        // exponent of the Eisenstat-Walker term.  2 (the textbook value for a quadratically converging Newton iteration)
        // over-solves here: the remainder a Newton step leaves was measured at ~(|R_k|/|R_k-1|)^1.5 |R_k| on the bench case
        k.eta_pow = getenv("ESPIC_MG_ETA_POW") ? atof(getenv("ESPIC_MG_ETA_POW")) : 1.5;
        k.link_scale = getenv("ESPIC_MG_LINK_SCALE") ? atof(getenv("ESPIC_MG_LINK_SCALE")) : 0.6;
        k.profile = getenv("ESPIC_MG_PROFILE") != nullptr;
        init = true;
    }
    return k;
}

// (re)build the hierarchy H for the current geometry; coarse-level arrays go to `external` if given (slab mode: a pool
// that the other ranks map), else to an allocation owned by H
static int mg_setup(espic_ctx *c, const StencilC &s, MgHierarchy *H, mgf *external)
{
    if (H->geom_version == c->geom_version && H->nlev > 0) return 0;
    if (H->pool && H->own_pool) { CK(cudaStreamSynchronize(c->stream)); CK(cudaFree(H->pool)); H->pool = nullptr; }
    long long dims[MG_MAX_LEVELS][3];
    int shifts[MG_MAX_LEVELS][3];
    const int nlev = mg_level_dims(s, dims, shifts);
    const long long total = mg_coarse_elems(s);
    if (external) { H->pool = external; H->own_pool = false; }
    else if (total > 0) {
        CK(cudaMalloc(&H->pool, (size_t)total * sizeof(mgf)));
        CK(cudaMemsetAsync(H->pool, 0, (size_t)total * sizeof(mgf), c->stream));
        H->own_pool = true;
    }
    mgf *p = H->pool;
    for (int l = 0; l < nlev; l++) {
        MgLevel &L = H->L[l];
        L.ni = (int)dims[l][0]; L.nj = (int)dims[l][1]; L.nk = (int)dims[l][2];
        L.fi = shifts[l][0]; L.fj = shifts[l][1]; L.fk = shifts[l][2];
        L.nn = dims[l][0] * dims[l][1] * dims[l][2];
        L.pitch = l == 0 ? mg_pitch(L.ni) : L.ni;
        if (l == 0) { L.diag = L.minv = L.cx = L.cy = L.cz = L.x = L.xn = L.b = L.dlx = L.dly = L.dlz = L.mass = nullptr; continue; }
        const long long pad = (L.nn + 3) & ~3ll;           // keep every array 16-byte aligned
        L.diag = p; p += pad; L.minv = p; p += pad; L.cx = p; p += pad; L.cy = p; p += pad; L.cz = p; p += pad;
        L.x = p; p += pad; L.xn = p; p += pad; L.b = p; p += pad;
        L.dlx = p; p += pad; L.dly = p; p += pad; L.dlz = p; p += pad; L.mass = p; p += pad;
    }
    H->nlev = nlev;
    for (int l = 1; l < nlev; l++) {
        const double scale = mg_knobs().link_scale;
        if (l == 1) k_mg_setup_from_types<<<nblk(H->L[1].nn, 256), 256, 0, c->stream>>>(s, c->node_type, H->L[1], scale);
        else k_mg_setup_from_level<<<nblk(H->L[l].nn, 256), 256, 0, c->stream>>>(H->L[l - 1], H->L[l], scale);
        LAUNCH_CHECK(c);
    }
    H->geom_version = c->geom_version;
    return 0;
}

// dynamic shared memory: the coarsest level's arrays, then (128-byte aligned) the TMA ring
static size_t mg_smem_coarse_bytes(const MgHierarchy *H)
{
    size_t b = 0;
    if (H->nlev > 1) {
        const MgLevel &L = H->L[H->nlev - 1];
        b = L.nn <= MG_STAGE_NODES ? (size_t)8 * (L.nn + 2 * (long long)L.ni * L.nj) * sizeof(mgf) : (size_t)3 * L.nn * sizeof(mgf);
    }
    return (b + 127) & ~(size_t)127;
}
static size_t mg_smem_bytes(const MgHierarchy *H) { return mg_smem_coarse_bytes(H) + MG_RING_BYTES; }

// 3-D tensor map of one padded fine vector: box = one plane of a tile plus halo.  The encoder lives in libcuda; it is looked
// up through the runtime so that this library needs no link-time dependency on the driver.
static int mg_make_map(CUtensorMap *m, void *base, bool f64, int pitch, int nj, int nk)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
            espic_set_error("cuTensorMapEncodeTiled is not available from this driver (the multigrid solver stages its tiles with TMA)");
            return -1;
        }
        fn = (EncodeFn)p;
    }
    const cuuint64_t es = f64 ? 8 : 4;
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)nj, (cuuint64_t)nk};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * es, (cuuint64_t)pitch * nj * es};
    const cuuint32_t box[3] = {(cuuint32_t)(f64 ? MG_TD_W : MG_TF_W), MG_TY + 2, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(m, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { espic_set_error("cuTensorMapEncodeTiled failed (%d) for a %d x %d x %d vector", (int)r, pitch, nj, nk); return -1; }
    return 0;
}

// fine vectors in one block of 6 * nnp doubles: r | delta | d0 | d1 (FP64), then z | diagf | winv | rf (FP32)
static int mg_fill_fine(MgnArgs &a, const StencilC &s, const MgHierarchy *H, double *base, long long nnp)
{
    a.pitch = H->L[0].pitch;
    a.psk = (long long)a.pitch * s.nj;
    double *r = base, *delta = base + nnp, *d0 = base + 2 * nnp, *d1 = base + 3 * nnp;
    mgf *z = reinterpret_cast<mgf *>(base + 4 * nnp), *diagf = z + nnp, *winv = z + 2 * nnp, *rf = z + 3 * nnp;
    a.r = r; a.delta = delta; a.d0 = d0; a.d1 = d1; a.z = z; a.diagf = diagf; a.winv = winv; a.rf = rf;
    a.smem_ring_off = (int)mg_smem_coarse_bytes(H);
    int rc;
    if ((rc = mg_make_map(&a.tm_rf, rf, false, a.pitch, s.nj, s.nk))) return rc;
    if ((rc = mg_make_map(&a.tm_w, winv, false, a.pitch, s.nj, s.nk))) return rc;
    if ((rc = mg_make_map(&a.tm_d0, d0, true, a.pitch, s.nj, s.nk))) return rc;
    if ((rc = mg_make_map(&a.tm_d1, d1, true, a.pitch, s.nj, s.nk))) return rc;
    if ((rc = mg_make_map(&a.tm_z, z, false, a.pitch, s.nj, s.nk))) return rc;
    if ((rc = mg_make_map(&a.tm_g, diagf, false, a.pitch, s.nj, s.nk))) return rc;
    return 0;
}

template <typename K>
static int mg_grid(espic_ctx *c, K kernel, size_t smem, const StencilC &s, int planes, int kunit, int *grid)
{
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bps = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, MG_THREADS, smem));
    if (bps < 1) { espic_set_error("the multigrid Newton kernel cannot be made resident"); return -1; }
    static const int bps_env = getenv("ESPIC_MG_BLOCKS_PER_SM") ? atoi(getenv("ESPIC_MG_BLOCKS_PER_SM")) : 0;
    if (bps_env > 0 && bps_env < bps) bps = bps_env;
    // no more blocks than marching work items (small meshes: fewer blocks = cheaper barriers)
    const long long units = (long long)((s.ni + MG_TX - 1) / MG_TX) * ((s.nj + MG_TY - 1) / MG_TY) * ((planes + kunit - 1) / kunit);
    *grid = (int)std::min<long long>((long long)bps * c->sm_count, std::max<long long>(units, 1));
    if (*grid > 1024) *grid = 1024;
    return 0;
}

static void mg_report(espic_ctx *c, const double *h, const espic_solve_params *p, espic_solve_info *info, const char *tag)
{
    info->converged = h[MGN_OUT_CONVERGED] != 0.0;
    info->nr_iters = (int)h[MGN_OUT_NEWTON];
    info->lin_iters = (long long)h[MGN_OUT_LIN];
    info->residual = h[MGN_OUT_YNORM];
    if (h[MGN_OUT_LINFAIL] != 0.0) fprintf(stderr, "PCG failed to converge, norm(g) = %g\n", h[MGN_OUT_LINL2]);
    if (!info->converged) printf("NR+PCG failed to converge, norm = %g\n", h[MGN_OUT_YNORM]);
    if (mg_knobs().profile && c->rank == 0) {
        fprintf(stderr, "[%s] newton steps %d, CG its %lld, |R| %.3e -> %.3e, |y| %.3e;", tag, info->nr_iters, info->lin_iters,
                h[MGN_OUT_RNORM0], h[MGN_OUT_RNORM], h[MGN_OUT_YNORM]);
        for (int q = 0; q < std::min(info->nr_iters + 1, 10); q++) fprintf(stderr, "  |R%d| %.3e (%d its)", q, h[MGN_OUT_HIST + 2 * q], (int)h[MGN_OUT_HIST + 2 * q + 1]);
        fprintf(stderr, "\n");
    }
}

static int mg_print_profile(espic_ctx *c, long long lin_iters, const char *tag)
{
    unsigned long long hp[16];
    CK(cudaMemcpyAsync(hp, c->dscal + 40, sizeof(hp), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemsetAsync(c->dscal + 40, 0, sizeof(hp), c->stream));
    const double n = (double)std::max<long long>(lin_iters, 1);
    fprintf(stderr, "[%s profile] us per CG iteration: down0 %.1f  down %.1f  coarsest %.1f  up %.1f  up0+rz %.1f  d %.1f  r %.1f | per solve: "
                    "linearise %.1f  galerkin %.1f  update %.1f\n", tag, hp[0] * 1e-3 / n, hp[1] * 1e-3 / n, hp[2] * 1e-3 / n, hp[3] * 1e-3 / n,
            hp[4] * 1e-3 / n, hp[5] * 1e-3 / n, hp[6] * 1e-3 / n, hp[7] * 1e-3, hp[8] * 1e-3, hp[9] * 1e-3);
    return 0;
}

// Newton + multigrid-preconditioned CG, one cooperative launch
static int solve_nrpcg_mg(espic_ctx *c, const espic_solve_params *p, espic_solve_info *info)
{
    int r;
    if ((r = ensure_node_types(c, 0))) return r;
    StencilC s = make_stencil(c->m);
    MgHierarchy *H = g_mg_of(c);
    if ((r = mg_setup(c, s, H, nullptr))) return r;
    const MgKnobs &kn = mg_knobs();
    const size_t smem = mg_smem_bytes(H);
    int grid = 0;
    if ((r = mg_grid(c, k_mg_newton, smem, s, s.nk, H->nlev > 1 ? (1 << H->L[1].fk) : 1, &grid))) return r;
    if ((r = ensure_buf(&c->red, &c->red_cap, 3ll * grid + 64, c->stream))) return r;
    double *dout = reinterpret_cast<double *>(c->dscal + 64);        // 32 doubles: slots 64..95
    MgnArgs a;
    memset(&a, 0, sizeof(a));
    a.s = s; a.nlev = H->nlev; a.coarse_sweeps = kn.coarse_sweeps; a.first_redundant = H->nlev - 1;
    for (int l = 0; l < H->nlev; l++) a.L[l] = H->L[l];
    a.type = c->node_type; a.rho = c->rho; a.phi = c->phi; a.phi_user = c->phi;
    {
        const long long nnp = (long long)H->L[0].pitch * s.nj * s.nk;
        if (H->fine_nnp != nnp) {
            if (H->fine) { CK(cudaStreamSynchronize(c->stream)); CK(cudaFree(H->fine)); H->fine = nullptr; }
            CK(cudaMalloc(&H->fine, (size_t)nnp * 6 * sizeof(double)));                 // 4 double + 4 float vectors
            CK(cudaMemsetAsync(H->fine, 0, (size_t)nnp * 6 * sizeof(double), c->stream));  // the row pads must be (and stay) zero
            H->fine_nnp = nnp;
        }
        if ((r = mg_fill_fine(a, s, H, H->fine, nnp))) return r;
    }
    a.part = c->red;
    a.phi0 = p->phi0; a.Te0 = p->Te0; a.n0 = p->n0;
    a.max_it = p->max_it; a.nr_max_it = p->nr_max_it; a.tol = p->tol; a.nr_tol = p->nr_tol;
    a.eta0 = kn.eta0; a.eta_max = kn.eta_max; a.gamma = kn.gamma; a.eta_pow = kn.eta_pow;
    a.out = dout;
    a.prof = kn.profile ? c->dscal + 40 : nullptr;
    void *args[] = {&a};
    CK(cudaLaunchCooperativeKernel((void *)k_mg_newton, dim3(grid), dim3(MG_THREADS), args, smem, c->stream));
    LAUNCH_CHECK(c);
    for (int level = 0; level < 3; level++) {
        k_mirror<<<nblk(s.nn, 256), 256, 0, c->stream>>>(s, c->node_type, c->phi, level);
        LAUNCH_CHECK(c);
    }
    double *h = reinterpret_cast<double *>(c->hpin) + 64;
    CK(cudaMemcpyAsync(h, dout, 32 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    mg_report(c, h, p, info, "mg");
    if (kn.profile) return mg_print_profile(c, info->lin_iters, "mg");
    return 0;
}


// ====================================================================================================================
// Slab-decomposed variant (ESPIC_SOLVE_PCG_MG_SLAB): one k-slab per rank, all traffic through peer-mapped memory
// ====================================================================================================================

struct SlabState {
    bool ready = false;
    long long geom_version = -1;
    double *pool = nullptr;            // this rank's pool; every rank carves it identically
    size_t pool_doubles = 0;
    void *peer_base[MG_MAX_RANKS] = {nullptr};
    // carved arrays
    double *fine, *phi, *part;
    long long nnp;
    mgf *coarse;
    unsigned long long *flags, *nflags, *epoch, *arrive, *release, *abort_word, *larrive, *lrelease;
    MgHierarchy H;
    OwnSlab own;
};

static void slab_destroy(espic_ctx *c)
{
    if (!c->slab) return;
    SlabState *S = static_cast<SlabState *>(c->slab);
    for (int p = 0; p < MG_MAX_RANKS; p++)
        if (S->peer_base[p] && p != c->rank) cudaIpcCloseMemHandle(S->peer_base[p]);
    cudaFree(S->pool);
    delete S;
    c->slab = nullptr;
}

static int slab_setup(espic_ctx *c, const StencilC &s)
{
    if (c->nranks < 2 || !c->nccl) { espic_set_error("ESPIC_SOLVE_PCG_MG_SLAB needs espic_comm_init with at least 2 ranks"); return -1; }
    if (c->nranks > MG_MAX_RANKS) { espic_set_error("ESPIC_SOLVE_PCG_MG_SLAB supports at most %d ranks", MG_MAX_RANKS); return -1; }
    if (!c->slab) c->slab = new SlabState();
    SlabState *S = static_cast<SlabState *>(c->slab);
    if (S->ready && S->geom_version == c->geom_version) return 0;
    if (S->ready) { espic_set_error("ESPIC_SOLVE_PCG_MG_SLAB: the geometry changed after the slab solver was set up"); return -1; }
    long long dims[MG_MAX_LEVELS][3];
    int shifts[MG_MAX_LEVELS][3];
    const int nlev = mg_level_dims(s, dims, shifts);
    const int unit = mg_slab_plane_unit(nlev, shifts);        // fine planes per coarsest plane
    if (s.nk % (c->nranks * unit) != 0) {
        espic_set_error("ESPIC_SOLVE_PCG_MG_SLAB: nk=%d must be a multiple of nranks*2^(k-coarsenings) = %d", s.nk, c->nranks * unit);
        return -1;
    }
    const int planes = s.nk / c->nranks;
    // ---- pool: identical carving on every rank (counted in doubles; FP32 arrays take half)
    const long long nnp = (long long)mg_pitch(s.ni) * s.nj * s.nk;          // padded rows (a multiple of 4 elements)
    const long long coarse = (mg_coarse_elems(s) + 1) / 2;
    const long long nparts = 3ll * c->nranks * 1024;
    S->pool_doubles = (size_t)(6 * nnp + nnp + coarse + nparts + 256);
    CK(cudaMalloc(&S->pool, S->pool_doubles * sizeof(double)));
    CK(cudaMemsetAsync(S->pool, 0, S->pool_doubles * sizeof(double), c->stream));
    double *p = S->pool;
    S->fine = p; p += 6 * nnp; S->phi = p; p += nnp;          // fine vectors (mg_fill_fine's layout), working copy of phi
    S->nnp = nnp;
    S->coarse = reinterpret_cast<mgf *>(p); p += coarse;
    S->part = p; p += nparts;
    S->flags = reinterpret_cast<unsigned long long *>(p); p += 16;
    S->nflags = reinterpret_cast<unsigned long long *>(p); p += 16;
    S->epoch = reinterpret_cast<unsigned long long *>(p); p += 16;
    S->arrive = reinterpret_cast<unsigned long long *>(p); p += 16;
    S->release = reinterpret_cast<unsigned long long *>(p); p += 16;
    S->abort_word = reinterpret_cast<unsigned long long *>(p); p += 16;
    S->larrive = reinterpret_cast<unsigned long long *>(p); p += 16;
    S->lrelease = reinterpret_cast<unsigned long long *>(p); p += 16;
    // ---- exchange IPC handles through the NCCL communicator and map the peers
    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, S->pool));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    char *dh = nullptr;
    CK(cudaMalloc(&dh, 64 * MG_MAX_RANKS));
    CK(cudaMemcpyAsync(dh + 64 * c->rank, &mine, 64, cudaMemcpyHostToDevice, c->stream));
    int r;
    if ((r = espic_comm_allgather_bytes(c, dh, 64))) return r;
    cudaIpcMemHandle_t all[MG_MAX_RANKS];
    CK(cudaMemcpyAsync(all, dh, 64 * c->nranks, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaFree(dh));
    OwnSlab &own = S->own;
    own.rank = c->rank; own.nranks = c->nranks; own.flags = S->flags; own.epoch = 0;
    own.nflags = S->nflags; own.nepoch = 0; own.lepoch_all = 0;
    own.neighbour_sync = getenv("ESPIC_MG_SLAB_NEIGHBOUR_SYNC") ? atoi(getenv("ESPIC_MG_SLAB_NEIGHBOUR_SYNC")) : 1;
    own.arrive = S->arrive; own.release = S->release; own.abort_word = S->abort_word;
    own.larrive = S->larrive; own.lrelease = S->lrelease; own.lepoch = 0;
    for (int q = 0; q < MG_MAX_RANKS; q++) own.peer[q] = 0;
    for (int q = 0; q < c->nranks; q++) {
        if (q == c->rank) { S->peer_base[q] = S->pool; continue; }
        CK(cudaIpcOpenMemHandle(&S->peer_base[q], all[q], cudaIpcMemLazyEnablePeerAccess));
        own.peer[q] = (long long)((char *)S->peer_base[q] - (char *)S->pool);
    }
    for (int l = 0; l < MG_MAX_LEVELS; l++) { own.k0[l] = 0; own.k1[l] = 0; }
    for (int l = 0, sh = 0; l < nlev; l++) {
        sh += shifts[l][2];
        own.k0[l] = (c->rank * planes) >> sh;
        own.k1[l] = (c->rank + 1 == c->nranks) ? (int)dims[l][2] : (((c->rank + 1) * planes) >> sh);
    }
    // ---- hierarchy inside the pool
    if ((r = mg_setup(c, s, &S->H, S->coarse))) return r;
    // every rank's pool must be zeroed and mapped before anybody stores into it: a collective on the stream + sync
    if ((r = espic_comm_allgather_doubles(c, S->part, 1))) return r;
    CK(cudaStreamSynchronize(c->stream));
    S->geom_version = c->geom_version;
    S->ready = true;
    return 0;
}

static int solve_nrpcg_mg_slab(espic_ctx *c, const espic_solve_params *p, espic_solve_info *info)
{
    int r;
    if ((r = ensure_node_types(c, 0))) return r;
    StencilC s = make_stencil(c->m);
    if ((r = slab_setup(c, s))) return r;
    SlabState *S = static_cast<SlabState *>(c->slab);
    MgHierarchy *H = &S->H;
    const MgKnobs &kn = mg_knobs();
    const long long slab_nn = (long long)(s.nk / c->nranks) * s.sk;
    const size_t smem = mg_smem_bytes(H);
    // the same grid on every rank (the partial-sum layout depends on it)
    int grid = 0;
    if ((r = mg_grid(c, k_mg_newton_slab, smem, s, s.nk / c->nranks, H->nlev > 1 ? (1 << H->L[1].fk) : 1, &grid))) return r;
    double *dout = reinterpret_cast<double *>(c->dscal + 64);
    MgnArgs a;
    memset(&a, 0, sizeof(a));
    a.s = s; a.nlev = H->nlev; a.coarse_sweeps = kn.coarse_sweeps;
    // Levels with at most MG_SLAB_REDUNDANT_NODES nodes (override: ESPIC_MG_SLAB_REDUNDANT_NODES, 0 = only the coarsest)
    // are solved by every rank in full instead of by slabs: 2 inter-GPU barriers less per level and V-cycle for redundant
    // work on a small level (profiles/r1_slab_redundant_levels_n4.txt).
    {
        long long dims[MG_MAX_LEVELS][3];
        for (int l = 0; l < H->nlev; l++) { dims[l][0] = H->L[l].ni; dims[l][1] = H->L[l].nj; dims[l][2] = H->L[l].nk; }
        a.first_redundant = mg_first_redundant(H->nlev, dims, mg_slab_redundant_limit());
    }
    for (int l = 0; l < H->nlev; l++) a.L[l] = H->L[l];
    a.type = c->node_type; a.rho = c->rho; a.phi = S->phi; a.phi_user = c->phi;
    if ((r = mg_fill_fine(a, s, H, S->fine, S->nnp))) return r;
    a.part = S->part;
    a.phi0 = p->phi0; a.Te0 = p->Te0; a.n0 = p->n0;
    a.max_it = p->max_it; a.nr_max_it = p->nr_max_it; a.tol = p->tol; a.nr_tol = p->nr_tol;
    a.eta0 = kn.eta0; a.eta_max = kn.eta_max; a.gamma = kn.gamma; a.eta_pow = kn.eta_pow;
    a.out = dout;
    a.prof = kn.profile ? c->dscal + 40 : nullptr;
    OwnSlab own = S->own;
    void *args[] = {&a, &own, &S->epoch};
    CK(cudaLaunchCooperativeKernel((void *)k_mg_newton_slab, dim3(grid), dim3(MG_THREADS), args, smem, c->stream));
    LAUNCH_CHECK(c);
    // every rank needs the whole potential: gather the slabs.  The collective is also the fence between solves: no rank can
    // start the next solve's peer stores before every rank has left this kernel.
    if ((r = espic_comm_allgather_doubles(c, c->phi, (size_t)slab_nn))) return r;
    for (int level = 0; level < 3; level++) {
        k_mirror<<<nblk(s.nn, 256), 256, 0, c->stream>>>(s, c->node_type, c->phi, level);
        LAUNCH_CHECK(c);
    }
    double *h = reinterpret_cast<double *>(c->hpin) + 64;
    CK(cudaMemcpyAsync(h, dout, 32 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (h[MGN_OUT_ABORT] != 0.0) {
        espic_set_error("ESPIC_SOLVE_PCG_MG_SLAB: an inter-GPU barrier timed out (a peer rank did not arrive); the solve was abandoned");
        return -1;
    }
    mg_report(c, h, p, info, "mg slab");
    if (kn.profile && c->rank == 0) return mg_print_profile(c, info->lin_iters, "mg slab");
    return 0;
}
