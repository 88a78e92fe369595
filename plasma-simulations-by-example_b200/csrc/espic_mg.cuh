// espic_mg.cuh -- aggregation-multigrid preconditioned CG for the Newton systems of the Boltzmann-electron Poisson solve.
// Textually included by espic_fields.cu (it uses StencilC, the node types and the SPD Newton kernels defined there).
//
// The reference preconditions CG with the matrix diagonal (PotentialSolver::solvePCGLinear, PotentialSolver.cpp:299-331);
// on the 128^3 mesh that takes ~300 iterations per Newton step and is >60 % of a PIC step.  ESPIC_SOLVE_PCG_MG keeps the
// Newton iteration, the SPD system K = -(L - diag P) on the REG nodes, the stopping tests (sqrt(sum r^2 / n) < tol for
// CG, sqrt(sum y^2 / n) < nr_tol for Newton) and replaces only the preconditioner by one multigrid V(1,1) cycle:
//   * coarsening by 2x2x2 aggregation of nodes, piecewise-constant prolongation P, restriction P^T, Galerkin coarse
//     operators P^T K P -- for a 7-point stencil these are again 7-point stencils (diagonal + three link arrays);
//   * damped-Jacobi smoothing (w = 0.9), one sweep before and one after the coarse correction; both are fused with the
//     grid transfer (down pass: smooth from zero + residual + restriction; up pass: prolongation + smooth), so a
//     level costs two passes and two grid-wide barriers per V-cycle;
//   * a few Jacobi sweeps on the coarsest level.
// The cycle is a symmetric positive definite operator (self-adjoint smoother, R = P^T), which CG requires.
// Everything runs in ONE persistent cooperative kernel per linear solve; convergence is tested on the device.
#pragma once

#define MG_MAX_LEVELS 8
#define MG_OMEGA 0.9
#define MG_COARSE_SWEEPS 6
#define MG_COARSEST_NODES 4096      // stop coarsening once a level is this small
#define MG_SLAB_REDUNDANT_NODES 65536   // slab mode: levels up to this size are solved by every rank in full

struct MgLevel {
    int ni, nj, nk;
    int fi, fj, fk;           // log2 of the coarsening factor from the next finer level to this one, per dimension (0 or 1):
                              // a direction whose spacing is already > sqrt(2) x the smallest one is not coarsened (semi-
                              // coarsening: the reference mesh has dz = 2 dx, a 4:4:1 anisotropic stencil on which point
                              // smoothers with full coarsening converge ~1.4x slower)
    long long nn;
    double *diag, *minv;      // Galerkin diagonal and its inverse (0 on nodes without unknowns)
    double *cx, *cy, *cz;     // link to the +x / +y / +z neighbour (>= 0; K = diag - sum links); level 0: not stored
    double *x, *xn, *b;       // pre-smoothed iterate, post-smoothed iterate, right-hand side
};

struct MgHierarchy {
    int nlev = 0;
    MgLevel L[MG_MAX_LEVELS];
    long long geom_version = -1;
    double *pool = nullptr;   // one allocation for all coarse-level arrays
    uint8_t *nbmask = nullptr; // per fine node: bit b set iff neighbour b (-x,+x,-y,+y,-z,+z) is an unknown (REG)
    // |R after the first Newton update| / |R before it| seen in the previous solve: the first linear solve of the next
    // solve is stopped a decade below the nonlinear residual it cannot remove anyway (inexact Newton forcing term)
    double newton_ratio = 0;
};

static MgHierarchy *g_mg_of(espic_ctx *c);   // stored in the context (espic_internal.cuh: void *mg)

// ---- setup kernels -------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_mg_nbmask(StencilC s, const uint8_t *__restrict__ type, uint8_t *__restrict__ nbmask)
{
    long long u = blockIdx.x * 256ll + threadIdx.x;
    if (u >= s.nn) return;
    unsigned m = 0;
    if (type[u] == NT_REG) {         // REG nodes are interior: all six neighbours exist
        m |= (type[u - 1] == NT_REG) << 0;    m |= (type[u + 1] == NT_REG) << 1;
        m |= (type[u - s.sj] == NT_REG) << 2; m |= (type[u + s.sj] == NT_REG) << 3;
        m |= (type[u - s.sk] == NT_REG) << 4; m |= (type[u + s.sk] == NT_REG) << 5;
        m |= 64;                              // bit 6: the node itself is an unknown
    }
    nbmask[u] = (uint8_t)m;
}

// links of level 1 from the fine node types: a fine link (u, u+e) exists iff both ends are REG; its weight is g = 1/dh^2
__global__ void __launch_bounds__(256) k_mg_links_from_types(StencilC s, const uint8_t *__restrict__ type, MgLevel C)
{
    long long I = blockIdx.x * 256ll + threadIdx.x;
    if (I >= C.nn) return;
    const int ci = (int)(I % C.ni), cj = (int)((I / C.ni) % C.nj), ck = (int)(I / ((long long)C.ni * C.nj));
    double lx = 0, ly = 0, lz = 0;
    for (int dk = 0; dk <= C.fk; dk++)
        for (int dj = 0; dj <= C.fj; dj++)
            for (int di = 0; di <= C.fi; di++) {
                const int i = (ci << C.fi) + di, j = (cj << C.fj) + dj, k = (ck << C.fk) + dk;
                if (i >= s.ni || j >= s.nj || k >= s.nk) continue;
                const long long u = (long long)k * s.sk + (long long)j * s.sj + i;
                if (type[u] != NT_REG) continue;
                if (di == C.fi && i + 1 < s.ni && type[u + 1] == NT_REG) lx += s.gdx2;
                if (dj == C.fj && j + 1 < s.nj && type[u + s.sj] == NT_REG) ly += s.gdy2;
                if (dk == C.fk && k + 1 < s.nk && type[u + s.sk] == NT_REG) lz += s.gdz2;
            }
    C.cx[I] = lx; C.cy[I] = ly; C.cz[I] = lz;
}

// links of level l+1 from the links of level l (l >= 1)
__global__ void __launch_bounds__(256) k_mg_links_from_links(MgLevel F, MgLevel C)
{
    long long I = blockIdx.x * 256ll + threadIdx.x;
    if (I >= C.nn) return;
    const int ci = (int)(I % C.ni), cj = (int)((I / C.ni) % C.nj), ck = (int)(I / ((long long)C.ni * C.nj));
    double lx = 0, ly = 0, lz = 0;
    for (int dk = 0; dk <= C.fk; dk++)
        for (int dj = 0; dj <= C.fj; dj++)
            for (int di = 0; di <= C.fi; di++) {
                const int i = (ci << C.fi) + di, j = (cj << C.fj) + dj, k = (ck << C.fk) + dk;
                if (i >= F.ni || j >= F.nj || k >= F.nk) continue;
                const long long u = ((long long)k * F.nj + j) * F.ni + i;
                if (di == C.fi) lx += F.cx[u];
                if (dj == C.fj) ly += F.cy[u];
                if (dk == C.fk) lz += F.cz[u];
            }
    C.cx[I] = lx; C.cy[I] = ly; C.cz[I] = lz;
}

// Galerkin diagonal of level 1: sum of the fine diagonals minus twice the fine links inside the aggregate
__global__ void __launch_bounds__(256) k_mg_diag_from_fine(StencilC s, const uint8_t *__restrict__ type, const double *__restrict__ diagJ, MgLevel C)
{
    long long I = blockIdx.x * 256ll + threadIdx.x;
    if (I >= C.nn) return;
    const int ci = (int)(I % C.ni), cj = (int)((I / C.ni) % C.nj), ck = (int)(I / ((long long)C.ni * C.nj));
    double d = 0;
    for (int dk = 0; dk <= C.fk; dk++)
        for (int dj = 0; dj <= C.fj; dj++)
            for (int di = 0; di <= C.fi; di++) {
                const int i = (ci << C.fi) + di, j = (cj << C.fj) + dj, k = (ck << C.fk) + dk;
                if (i >= s.ni || j >= s.nj || k >= s.nk) continue;
                const long long u = (long long)k * s.sk + (long long)j * s.sj + i;
                if (type[u] != NT_REG) continue;
                d += diagJ[u];
                if (di < C.fi && i + 1 < s.ni && type[u + 1] == NT_REG) d -= 2 * s.gdx2;
                if (dj < C.fj && j + 1 < s.nj && type[u + s.sj] == NT_REG) d -= 2 * s.gdy2;
                if (dk < C.fk && k + 1 < s.nk && type[u + s.sk] == NT_REG) d -= 2 * s.gdz2;
            }
    C.diag[I] = d;
    C.minv[I] = d > 0 ? 1.0 / d : 0.0;
}

__global__ void __launch_bounds__(256) k_mg_diag_from_level(MgLevel F, MgLevel C)
{
    long long I = blockIdx.x * 256ll + threadIdx.x;
    if (I >= C.nn) return;
    const int ci = (int)(I % C.ni), cj = (int)((I / C.ni) % C.nj), ck = (int)(I / ((long long)C.ni * C.nj));
    double d = 0;
    for (int dk = 0; dk <= C.fk; dk++)
        for (int dj = 0; dj <= C.fj; dj++)
            for (int di = 0; di <= C.fi; di++) {
                const int i = (ci << C.fi) + di, j = (cj << C.fj) + dj, k = (ck << C.fk) + dk;
                if (i >= F.ni || j >= F.nj || k >= F.nk) continue;
                const long long u = ((long long)k * F.nj + j) * F.ni + i;
                d += F.diag[u];
                if (di < C.fi && i + 1 < F.ni) d -= 2 * F.cx[u];
                if (dj < C.fj && j + 1 < F.nj) d -= 2 * F.cy[u];
                if (dk < C.fk && k + 1 < F.nk) d -= 2 * F.cz[u];
            }
    C.diag[I] = d;
    C.minv[I] = d > 0 ? 1.0 / d : 0.0;
}

// ---- who owns what: single GPU, or one k-slab per rank with peer-mapped pools -----------------------------------------
// Every pass below is written once and instantiated twice.  `Own` tells a pass which flat index range of a level this
// rank computes and what a store has to do besides writing locally:
//   OwnAll   one GPU: the whole level, plain stores, grid-wide barrier.
//   OwnSlab  rank r of R owns planes [k0,k1) of every level (boundaries are multiples of 2^(levels-1) fine planes, so an
//            aggregate never straddles two ranks).  All solver vectors live at the same offset of a pool that every rank
//            maps from every other rank (CUDA IPC over NVLink).  A value written on the first / last plane of the slab is
//            ALSO stored straight into the lower / upper neighbour's copy (peer store fused into the producing pass: the
//            halo exchange costs no extra pass and no NCCL call); dot-product partials are stored to every rank so all
//            ranks add the same numbers in the same order; the barrier is a grid barrier plus one flag per peer.
struct OwnAll {
    static constexpr bool slab = false;
    __device__ __forceinline__ long long lo(const MgLevel &L, int) const { return 0; }
    __device__ __forceinline__ long long hi(const MgLevel &L, int) const { return L.nn; }
    __device__ __forceinline__ void st(double *A, long long u, const MgLevel &, int, double v) const { A[u] = v; }
    __device__ __forceinline__ void st_all(double *A, long long u, double v) const { A[u] = v; }
    __device__ __forceinline__ int nparts(int nb) const { return nb; }
    __device__ __forceinline__ void put_partial(double *base, int nb, double v) const { base[blockIdx.x] = v; }
    __device__ __forceinline__ void barrier(cg::grid_group &grid) { grid.sync(); }
};

#define MG_MAX_RANKS 8
struct OwnSlab {
    static constexpr bool slab = true;
    int rank, nranks;
    int k0[MG_MAX_LEVELS], k1[MG_MAX_LEVELS];
    long long peer[MG_MAX_RANKS];        // byte distance from an address in this rank's pool to the same address in rank p's pool
    unsigned long long *flags;           // in the pool: flags[p] = last barrier epoch rank p has reached (written by rank p)
    unsigned long long epoch;            // barriers passed so far (identical on every rank)
    __device__ __forceinline__ long long lo(const MgLevel &L, int l) const { return (long long)k0[l] * L.ni * L.nj; }
    __device__ __forceinline__ long long hi(const MgLevel &L, int l) const { return (long long)k1[l] * L.ni * L.nj; }
    __device__ __forceinline__ void st(double *A, long long u, const MgLevel &L, int l, double v) const
    {
        A[u] = v;
        const long long plane = (long long)L.ni * L.nj;
        if (rank > 0 && u < lo(L, l) + plane) *reinterpret_cast<double *>(reinterpret_cast<char *>(A + u) + peer[rank - 1]) = v;
        if (rank + 1 < nranks && u >= hi(L, l) - plane) *reinterpret_cast<double *>(reinterpret_cast<char *>(A + u) + peer[rank + 1]) = v;
    }
    // store into every rank's copy (the right-hand side of the coarsest level: that level is solved by every rank in full)
    __device__ __forceinline__ void st_all(double *A, long long u, double v) const
    {
        for (int p = 0; p < nranks; p++) *reinterpret_cast<double *>(reinterpret_cast<char *>(A + u) + peer[p]) = v;
    }
    __device__ __forceinline__ int nparts(int nb) const { return nb * nranks; }
    __device__ __forceinline__ void put_partial(double *base, int nb, double v) const
    {
        for (int p = 0; p < nranks; p++)
            *reinterpret_cast<double *>(reinterpret_cast<char *>(base + rank * nb + blockIdx.x) + peer[p]) = v;
    }
    // All blocks of all ranks.  Local arrival (one atomic per block on a cumulative counter), then block 0 exchanges one
    // flag with every peer over NVLink, then it releases the local blocks: one local round trip plus one remote one.
    // Ordering: every block's thread 0 issues a system-scope fence after the block barrier and before arriving, block 0
    // fences again (system scope) between seeing all arrivals and signalling the peers, so a peer that sees the flag also
    // sees every halo value and partial sum stored before this barrier.
    unsigned long long *arrive, *release;     // in the local pool
    __device__ __forceinline__ void barrier(cg::grid_group &)
    {
        __syncthreads();
        epoch++;
        if (blockIdx.x == 0) {
            if (threadIdx.x < 32) {          // first warp: lane 0 collects the local arrivals, lane p talks to peer p
                const int lane = threadIdx.x;
                if (lane == 0) {
                    __threadfence_system();
                    volatile unsigned long long *arr = arrive;
                    const unsigned long long want = (unsigned long long)(gridDim.x - 1) * epoch;
                    while (*arr < want) { }
                    __threadfence_system();
                }
                __syncwarp();
                if (lane < nranks) {
                    *reinterpret_cast<volatile unsigned long long *>(reinterpret_cast<char *>(flags + rank) + peer[lane]) = epoch;
                    volatile unsigned long long *mine = flags + lane;
                    while (*mine < epoch) { }
                    __threadfence_system();
                }
                __syncwarp();
                if (lane == 0) {
                    volatile unsigned long long *rel = release;
                    *rel = epoch;
                    __threadfence();
                }
            }
        } else if (threadIdx.x == 0) {
            __threadfence_system();
            atomicAdd(arrive, 1ull);
            volatile unsigned long long *rel = release;
            while (*rel < epoch) { }
            __threadfence();
        }
        __syncthreads();
    }
};

// ---- device pieces of the V-cycle (grid-stride; the caller separates them with barriers) ---------------------------------

// off-diagonal part of K v at fine node u (level 0): neighbours outside the REG set carry v == 0
__device__ __forceinline__ double mg_offdiag0(const StencilC &s, const double *__restrict__ v, long long u)
{
    return s.gdx2 * (v[u - 1] + v[u + 1]) + s.gdy2 * (v[u - s.sj] + v[u + s.sj]) + s.gdz2 * (v[u - s.sk] + v[u + s.sk]);
}

// sum over neighbours of link * f(neighbour) on a coarse level, f given as a functor of the neighbour's flat index
template <typename F>
__device__ __forceinline__ double mg_offdiag(const MgLevel &L, int i, int j, int k, long long u, F f)
{
    const long long sj = L.ni, sk = (long long)L.ni * L.nj;
    double a = 0;
    if (i > 0) a += L.cx[u - 1] * f(u - 1);
    if (i + 1 < L.ni) a += L.cx[u] * f(u + 1);
    if (j > 0) a += L.cy[u - sj] * f(u - sj);
    if (j + 1 < L.nj) a += L.cy[u] * f(u + sj);
    if (k > 0) a += L.cz[u - sk] * f(u - sk);
    if (k + 1 < L.nk) a += L.cz[u] * f(u + sk);
    return a;
}

// Down pass from the fine level: x0 = w D^-1 r is already stored (written together with r); the residual r - K x0 is
// summed over each aggregate -> b of level 1.  One thread per COARSE node (it owns the 8 children).
template <class Own>
__device__ __forceinline__ void mg_down0(const Own &own, const StencilC &s, const double *__restrict__ r,
                                         const double *__restrict__ diag, const double *__restrict__ x0, const MgLevel &C,
                                         long long t0, long long stride, bool to_all)
{
    for (long long I = own.lo(C, 1) + t0; I < own.hi(C, 1); I += stride) {
        const int ci = (int)(I % C.ni), cj = (int)((I / C.ni) % C.nj), ck = (int)(I / ((long long)C.ni * C.nj));
        double sum = 0;
        if (C.minv[I] != 0) {
#pragma unroll
            for (int dk = 0; dk < 2; dk++)
#pragma unroll
                for (int dj = 0; dj < 2; dj++)
#pragma unroll
                    for (int di = 0; di < 2; di++) {
                        if (di > C.fi || dj > C.fj || dk > C.fk) continue;
                        const int i = (ci << C.fi) + di, j = (cj << C.fj) + dj, k = (ck << C.fk) + dk;
                        if (i >= s.ni || j >= s.nj || k >= s.nk) continue;
                        const long long u = (long long)k * s.sk + (long long)j * s.sj + i;
                        const double dg = diag[u];
                        if (dg == 0) continue;                 // not an unknown
                        sum += r[u] - (dg * x0[u] - mg_offdiag0(s, x0, u));
                    }
        }
        if (to_all) own.st_all(C.b, I, sum);
        else own.st(C.b, I, C, 1, sum);
    }
}

// The coarse levels are small: what matters there is the length of the dependent-load chain of a thread, not bandwidth.
// All coarse passes therefore spread one node over 8 consecutive lanes (the 8 children of an aggregate, or the 7 stencil
// terms of a node) and combine with three shuffles; the sum order is fixed, so results are reproducible.
__device__ __forceinline__ double mg_sum8(double v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// Down pass between coarse levels F (level lf) -> C: lane c of a group handles child c of coarse node I
template <class Own>
__device__ __forceinline__ void mg_down(const Own &own, const MgLevel &F, int lf, const MgLevel &C, long long t0, long long stride,
                                        bool to_all)
{
    const long long first = own.lo(C, lf + 1), total = (own.hi(C, lf + 1) - first) * 8;
    for (long long w = t0; (w & ~31LL) < total; w += stride) {
        const long long I = first + (w >> 3);
        const int c = (int)(w & 7);
        const bool live = w < total;
        double res = 0;
        if (live) {
            const int ci = (int)(I % C.ni), cj = (int)((I / C.ni) % C.nj), ck = (int)(I / ((long long)C.ni * C.nj));
            const int di = c & 1, dj = (c >> 1) & 1, dk = c >> 2;
            const int i = (ci << C.fi) + di, j = (cj << C.fj) + dj, k = (ck << C.fk) + dk;
            if (di <= C.fi && dj <= C.fj && dk <= C.fk && i < F.ni && j < F.nj && k < F.nk) {
                const long long u = ((long long)k * F.nj + j) * F.ni + i;
                const double mi = F.minv[u];
                double xu = 0;
                if (mi != 0) {
                    xu = MG_OMEGA * F.b[u] * mi;
                    const double off = MG_OMEGA * mg_offdiag(F, i, j, k, u, [&](long long v) { return F.b[v] * F.minv[v]; });
                    res = F.b[u] - (F.diag[u] * xu - off);
                }
                own.st(F.x, u, F, lf, xu);
            }
        }
        res = mg_sum8(res);
        if (c == 0 && live) {
            if (to_all) own.st_all(C.b, I, res);
            else own.st(C.b, I, C, lf + 1, res);
        }
    }
}

// stencil term c of node u on a coarse level: c = 0 centre (b - diag*v), c = 1..6 the six links, c = 7 nothing
template <typename F>
__device__ __forceinline__ double mg_term(const MgLevel &L, int c, int i, int j, int k, long long u, F val, double &centre)
{
    const long long sj = L.ni, sk = (long long)L.ni * L.nj;
    switch (c) {
        case 0: centre = val(u, i, j, k); return L.b[u] - L.diag[u] * centre;
        case 1: return i > 0 ? L.cx[u - 1] * val(u - 1, i - 1, j, k) : 0.0;
        case 2: return i + 1 < L.ni ? L.cx[u] * val(u + 1, i + 1, j, k) : 0.0;
        case 3: return j > 0 ? L.cy[u - sj] * val(u - sj, i, j - 1, k) : 0.0;
        case 4: return j + 1 < L.nj ? L.cy[u] * val(u + sj, i, j + 1, k) : 0.0;
        case 5: return k > 0 ? L.cz[u - sk] * val(u - sk, i, j, k - 1) : 0.0;
        case 6: return k + 1 < L.nk ? L.cz[u] * val(u + sk, i, j, k + 1) : 0.0;
        default: return 0.0;
    }
}

// one damped-Jacobi sweep on level L (index l): out = in + w D^-1 (b - K in); 8 lanes per node
template <class Own>
__device__ __forceinline__ void mg_jacobi(const Own &own, const MgLevel &L, int l, const double *__restrict__ in,
                                          double *__restrict__ out, long long t0, long long stride)
{
    const long long first = own.lo(L, l), total = (own.hi(L, l) - first) * 8;
    for (long long w = t0; (w & ~31LL) < total; w += stride) {
        const long long u = first + (w >> 3);
        const int c = (int)(w & 7);
        const bool live = w < total;
        double term = 0, centre = 0, mi = 0;
        if (live) {
            mi = L.minv[u];
            if (mi != 0) {
                const int i = (int)(u % L.ni), j = (int)((u / L.ni) % L.nj), k = (int)(u / ((long long)L.ni * L.nj));
                term = mg_term(L, c, i, j, k, u, [&](long long v, int, int, int) { return in[v]; }, centre);
            }
        }
        const double tot = mg_sum8(term);
        centre = __shfl_sync(0xffffffffu, centre, (threadIdx.x & 31) & ~7);
        if (c == 0 && live) own.st(out, u, L, l, (mi != 0) ? centre + MG_OMEGA * mi * tot : 0.0);
    }
}

// Up pass on a coarse level F (index lf) with the correction e of the next coarser level C:
//   xn = (x + P e) + w D^-1 (b - K (x + P e))
template <class Own>
__device__ __forceinline__ void mg_up(const Own &own, const MgLevel &F, int lf, const MgLevel &C, const double *__restrict__ e,
                                      long long t0, long long stride)
{
    const long long first = own.lo(F, lf), count = own.hi(F, lf) - first;
    auto val = [&](long long v, int vi, int vj, int vk) {
        // links to nodes without unknowns are zero on coarse levels, so no mask is needed on the neighbours
        return F.x[v] + e[((long long)(vk >> C.fk) * C.nj + (vj >> C.fj)) * C.ni + (vi >> C.fi)];
    };
    if (count * 2 > stride) {          // a big level: one thread per node keeps every lane busy
        for (long long u = first + t0; u < first + count; u += stride) {
            const double mi = F.minv[u];
            double out = 0;
            if (mi != 0) {
                const int i = (int)(u % F.ni), j = (int)((u / F.ni) % F.nj), k = (int)(u / ((long long)F.ni * F.nj));
                double centre = 0, tot = 0;
#pragma unroll
                for (int c = 0; c < 7; c++) tot += mg_term(F, c, i, j, k, u, val, centre);
                out = centre + MG_OMEGA * mi * tot;
            }
            own.st(F.xn, u, F, lf, out);
        }
        return;
    }
    const long long total = count * 8;          // a small level: 8 lanes per node, one stencil term each
    for (long long w = t0; (w & ~31LL) < total; w += stride) {
        const long long u = first + (w >> 3);
        const int c = (int)(w & 7);
        const bool live = w < total;
        double term = 0, centre = 0, mi = 0;
        if (live) {
            mi = F.minv[u];
            if (mi != 0) {
                const int i = (int)(u % F.ni), j = (int)((u / F.ni) % F.nj), k = (int)(u / ((long long)F.ni * F.nj));
                term = mg_term(F, c, i, j, k, u, val, centre);
            }
        }
        const double tot = mg_sum8(term);
        centre = __shfl_sync(0xffffffffu, centre, (threadIdx.x & 31) & ~7);
        if (c == 0 && live) own.st(F.xn, u, F, lf, (mi != 0) ? centre + MG_OMEGA * mi * tot : 0.0);
    }
}

// Up pass on the fine level: z = (x0 + P e) + w D^-1 (r - K (x0 + P e)).  One thread per level-1 node: it holds the
// correction of its own aggregate and of the six neighbouring aggregates in registers and walks its 8 children, so the
// prolongated iterate never touches memory; nbmask replaces six mask loads per node.  Returns the thread's share of r.z
template <class Own>
__device__ __forceinline__ double mg_up0(const Own &own, const StencilC &s, const MgLevel &L0, const double *__restrict__ r,
                                         const double *__restrict__ diag, const double *__restrict__ minv,
                                         const double *__restrict__ x0, const uint8_t *__restrict__ nbmask, const MgLevel &C,
                                         const double *__restrict__ e, double *__restrict__ z, long long t0, long long stride)
{
    double acc = 0;
    const long long csj = C.ni, csk = (long long)C.ni * C.nj;
    for (long long I = own.lo(C, 1) + t0; I < own.hi(C, 1); I += stride) {
        const int ci = (int)(I % C.ni), cj = (int)((I / C.ni) % C.nj), ck = (int)(I / csk);
        const double e0 = e[I];
        const double exm = ci > 0 ? e[I - 1] : 0.0, exp_ = ci + 1 < C.ni ? e[I + 1] : 0.0;
        const double eym = cj > 0 ? e[I - csj] : 0.0, eyp = cj + 1 < C.nj ? e[I + csj] : 0.0;
        const double ezm = ck > 0 ? e[I - csk] : 0.0, ezp = ck + 1 < C.nk ? e[I + csk] : 0.0;
#pragma unroll
        for (int dk = 0; dk < 2; dk++)
#pragma unroll
            for (int dj = 0; dj < 2; dj++)
#pragma unroll
                for (int di = 0; di < 2; di++) {
                    if (di > C.fi || dj > C.fj || dk > C.fk) continue;
                    const int i = (ci << C.fi) + di, j = (cj << C.fj) + dj, k = (ck << C.fk) + dk;
                    if (i >= s.ni || j >= s.nj || k >= s.nk) continue;
                    const long long u = (long long)k * s.sk + (long long)j * s.sj + i;
                    const unsigned m = nbmask[u];
                    double zu = 0;
                    if (m & 64u) {
                        // the neighbour on the inner side of the aggregate shares e0, the outer one takes the next aggregate's
                        const double vxm = (m & 1u) ? x0[u - 1] + (di > 0 ? e0 : exm) : 0.0;
                        const double vxp = (m & 2u) ? x0[u + 1] + (di < C.fi ? e0 : exp_) : 0.0;
                        const double vym = (m & 4u) ? x0[u - s.sj] + (dj > 0 ? e0 : eym) : 0.0;
                        const double vyp = (m & 8u) ? x0[u + s.sj] + (dj < C.fj ? e0 : eyp) : 0.0;
                        const double vzm = (m & 16u) ? x0[u - s.sk] + (dk > 0 ? e0 : ezm) : 0.0;
                        const double vzp = (m & 32u) ? x0[u + s.sk] + (dk < C.fk ? e0 : ezp) : 0.0;
                        const double off = s.gdx2 * (vxm + vxp) + s.gdy2 * (vym + vyp) + s.gdz2 * (vzm + vzp);
                        const double xu = x0[u] + e0;
                        zu = xu + MG_OMEGA * minv[u] * (r[u] - (diag[u] * xu - off));
                        acc += r[u] * zu;
                    }
                    own.st(z, u, L0, 0, zu);
                }
    }
    return acc;
}

struct MgPcgArgs {
    StencilC s;
    int nlev;
    int coarse_sweeps;            // Jacobi sweeps on the coarsest level (even)
    int first_redundant;          // slab mode: first level that every rank solves in full (<= nlev-1, >= 1); unused on one GPU
    MgLevel L[MG_MAX_LEVELS];     // L[0]: dims, diag = diagJ, minv, x = x0 (= w D^-1 r, kept current with r); links unused
    const uint8_t *nbmask;
    double *delta, *r, *z, *d0, *d1, *q;     // r enters holding the right-hand side; d0/d1 ping-pong search directions
    double *part;
    int max_it;
    double tol;
    double rel_tol;               // stop at l2 < max(tol, rel_tol * l2_start)
    double *out;                  // converged, iterations, l2, l2 at the start
    unsigned long long *prof;     // optional: nanoseconds per phase as seen by block 0 (ESPIC_MG_PROFILE=1), 8 slots
};

__device__ __forceinline__ unsigned long long mg_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// phase ids: 0 down0, 1 coarser down passes, 2 coarsest sweeps, 3 coarse up passes, 4 up0 + r.z, 5 d/q pass, 6 r pass
#define MG_TICK(id) do { if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long n_ = mg_now(); a.prof[id] += n_ - tick; tick = n_; } } while (0)

// z = M^-1 r (one V-cycle); returns this thread's share of r.z.  Ends WITHOUT a barrier.
template <class Own>
__device__ __forceinline__ double mg_vcycle(cg::grid_group &grid, Own &own, const MgPcgArgs &a, long long t0, long long stride,
                                            unsigned long long &tick)
{
    const MgLevel &L0 = a.L[0];
    if (a.nlev == 1) {           // degenerate hierarchy (tiny mesh): plain Jacobi preconditioner
        double acc = 0;
        for (long long u = own.lo(L0, 0) + t0; u < own.hi(L0, 0); u += stride) {
            double zu = L0.minv[u] * a.r[u];
            own.st(a.z, u, L0, 0, zu);
            acc += a.r[u] * zu;
        }
        return acc;
    }
    // The coarse levels are tiny: in slab mode the right-hand side of level `lr` (a.first_redundant; by default the coarsest
    // level) is stored to EVERY rank, and every rank runs the levels lr .. coarsest redundantly in full with grid-local
    // barriers only (identical arithmetic -> identical result everywhere).  That removes 2 inter-GPU barriers per redundant
    // level and coarse_sweeps+1 for the coarsest from every V-cycle.
    const int lc = a.nlev - 1;
    const MgLevel &Lc = a.L[lc];
    int lr = lc;
    if constexpr (Own::slab) lr = a.first_redundant;
    OwnAll whole;
    mg_down0(own, a.s, a.r, L0.diag, L0.x, a.L[1], t0, stride, lr == 1);
    own.barrier(grid);
    MG_TICK(0);
    for (int l = 1; l + 1 < a.nlev; l++) {
        if (Own::slab && l >= lr) {
            mg_down(whole, a.L[l], l, a.L[l + 1], t0, stride, false);
            grid.sync();
        } else {
            mg_down(own, a.L[l], l, a.L[l + 1], t0, stride, l + 1 == lr);
            own.barrier(grid);
        }
    }
    MG_TICK(1);
    {
        for (long long u = t0; u < Lc.nn; u += stride) Lc.x[u] = MG_OMEGA * Lc.b[u] * Lc.minv[u];
        grid.sync();
        for (int sweep = 0; sweep < a.coarse_sweeps; sweep += 2) {
            mg_jacobi(whole, Lc, lc, Lc.x, Lc.xn, t0, stride);
            grid.sync();
            mg_jacobi(whole, Lc, lc, Lc.xn, Lc.x, t0, stride);
            grid.sync();
        }
    }
    MG_TICK(2);
    const double *e = Lc.x;
    for (int l = a.nlev - 2; l >= 1; l--) {
        if (Own::slab && l >= lr) {
            mg_up(whole, a.L[l], l, a.L[l + 1], e, t0, stride);
            grid.sync();
        } else {
            mg_up(own, a.L[l], l, a.L[l + 1], e, t0, stride);
            own.barrier(grid);
        }
        e = a.L[l].xn;
    }
    MG_TICK(3);
    return mg_up0(own, a.s, L0, a.r, L0.diag, L0.minv, L0.x, a.nbmask, a.L[1], e, a.z, t0, stride);
}

// Sum of per-block partials (of every rank in slab mode), computed redundantly by every block in the same fixed order
template <class Own>
__device__ __forceinline__ double mg_total(const Own &own, const double *part, int nb, double *sh, double *bcast)
{
    return grid_total(part, own.nparts(nb), sh, bcast);
}

template <class Own>
__device__ __forceinline__ void mg_pcg_body(MgPcgArgs &a, Own &own)
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
    const double *diag = L0.diag;
    const double *minv = L0.minv;
    double *x0 = L0.x;
    const long long ulo = own.lo(L0, 0), uhi = own.hi(L0, 0);

    // delta = 0: r = R; x0 = w D^-1 r; |r|
    double acc = 0;
    for (long long u = ulo + t0; u < uhi; u += stride) {
        const double r = a.r[u];
        own.st(x0, u, L0, 0, MG_OMEGA * r * minv[u]);
        acc += r * r;
    }
    double t = block_sum(acc, sh);
    if (threadIdx.x == 0) own.put_partial(pC, nb, t);
    own.barrier(grid);
    double l2 = sqrt(mg_total(own, pC, nb, sh, &bc) / (double)s.nn);
    const double l2_start = l2;
    const double stop = fmax(a.tol, a.rel_tol * l2_start);
    int it = 0, converged = l2 < stop;
    double rz = 0, beta = 0;
    double *d_old = a.d0, *d_new = a.d1;
    unsigned long long tick = mg_now();
    while (!converged && it < a.max_it) {
        // z = M^-1 r ; rz' = r.z
        acc = mg_vcycle(grid, own, a, t0, stride, tick);
        t = block_sum(acc, sh);
        if (threadIdx.x == 0) own.put_partial(pB, nb, t);
        own.barrier(grid);
        MG_TICK(4);
        const double rz_new = mg_total(own, pB, nb, sh, &bc);
        beta = (it == 0) ? 0.0 : rz_new / rz;
        rz = rz_new;
        // d = z + beta d (formed on the fly for the neighbours, written for this node) ; q = K d ; dq = d.q
        acc = 0;
        for (long long u = ulo + t0; u < uhi; u += stride) {
            const double dj = diag[u];
            double du = 0, qu = 0;
            if (dj != 0) {
                du = a.z[u] + beta * d_old[u];
                // z and d are identically zero outside the REG set: no neighbour masks
                const double off = s.gdx2 * ((a.z[u - 1] + beta * d_old[u - 1]) + (a.z[u + 1] + beta * d_old[u + 1])) +
                                   s.gdy2 * ((a.z[u - s.sj] + beta * d_old[u - s.sj]) + (a.z[u + s.sj] + beta * d_old[u + s.sj])) +
                                   s.gdz2 * ((a.z[u - s.sk] + beta * d_old[u - s.sk]) + (a.z[u + s.sk] + beta * d_old[u + s.sk]));
                qu = dj * du - off;
                acc += du * qu;
            }
            own.st(d_new, u, L0, 0, du);
            a.q[u] = qu;
        }
        t = block_sum(acc, sh);
        if (threadIdx.x == 0) own.put_partial(pA, nb, t);
        own.barrier(grid);
        MG_TICK(5);
        const double alpha = rz / mg_total(own, pA, nb, sh, &bc);
        // delta += alpha d ; r -= alpha q ; |r|
        acc = 0;
        for (long long u = ulo + t0; u < uhi; u += stride) {
            a.delta[u] = a.delta[u] + alpha * d_new[u];
            const double r = a.r[u] - alpha * a.q[u];
            a.r[u] = r;
            own.st(x0, u, L0, 0, MG_OMEGA * r * minv[u]);          // pre-smoothed iterate of the next V-cycle
            acc += r * r;
        }
        t = block_sum(acc, sh);
        if (threadIdx.x == 0) own.put_partial(pC, nb, t);
        own.barrier(grid);
        MG_TICK(6);
        l2 = sqrt(mg_total(own, pC, nb, sh, &bc) / (double)s.nn);
        it++;
        double *tmp = d_old; d_old = d_new; d_new = tmp;
        if (l2 < stop) converged = 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { a.out[0] = converged; a.out[1] = it; a.out[2] = l2; a.out[3] = l2_start; }
}

#ifndef MG_BLOCK
#define MG_BLOCK 512
#endif
__global__ void __launch_bounds__(MG_BLOCK, 1024 / MG_BLOCK) k_mg_pcg(MgPcgArgs a)
{
    OwnAll own;
    mg_pcg_body(a, own);
}

// slab-decomposed variant: one of these kernels per rank, running concurrently, talking through peer memory only
__global__ void __launch_bounds__(MG_BLOCK, 1024 / MG_BLOCK) k_mg_pcg_slab(MgPcgArgs a, OwnSlab own, unsigned long long *epoch_io)
{
    own.epoch = *epoch_io;
    mg_pcg_body(a, own);
    if (blockIdx.x == 0 && threadIdx.x == 0) *epoch_io = own.epoch;
}

// ---- host side -------------------------------------------------------------------------------------------------------

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
    return nlev;
}

static long long mg_coarse_doubles(const StencilC &s)
{
    long long dims[MG_MAX_LEVELS][3];
    int shifts[MG_MAX_LEVELS][3];
    const int nlev = mg_level_dims(s, dims, shifts);
    long long total = 0;
    for (int l = 1; l < nlev; l++) total += 8 * dims[l][0] * dims[l][1] * dims[l][2];
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

// (re)build the hierarchy H for the current geometry; coarse-level arrays go to `external` if given (slab mode: a pool
// that the other ranks map), else to an allocation owned by H
static int mg_setup(espic_ctx *c, const StencilC &s, MgHierarchy *H, double *external)
{
    if (H->geom_version == c->geom_version && H->nlev > 0) return 0;
    if (H->pool && !external) { CK(cudaStreamSynchronize(c->stream)); CK(cudaFree(H->pool)); H->pool = nullptr; }
    if (!H->nbmask) CK(cudaMalloc(&H->nbmask, (size_t)s.nn));
    k_mg_nbmask<<<nblk(s.nn, 256), 256, 0, c->stream>>>(s, c->node_type, H->nbmask);
    LAUNCH_CHECK(c);
    long long dims[MG_MAX_LEVELS][3];
    int shifts[MG_MAX_LEVELS][3];
    const int nlev = mg_level_dims(s, dims, shifts);
    const long long total = mg_coarse_doubles(s);
    if (external) H->pool = external;
    else if (total > 0) CK(cudaMalloc(&H->pool, (size_t)total * sizeof(double)));
    double *p = H->pool;
    for (int l = 0; l < nlev; l++) {
        MgLevel &L = H->L[l];
        L.ni = (int)dims[l][0]; L.nj = (int)dims[l][1]; L.nk = (int)dims[l][2];
        L.fi = shifts[l][0]; L.fj = shifts[l][1]; L.fk = shifts[l][2];
        L.nn = dims[l][0] * dims[l][1] * dims[l][2];
        if (l == 0) { L.diag = L.minv = L.cx = L.cy = L.cz = L.x = L.xn = L.b = nullptr; continue; }
        L.diag = p; p += L.nn; L.minv = p; p += L.nn; L.cx = p; p += L.nn; L.cy = p; p += L.nn; L.cz = p; p += L.nn;
        L.x = p; p += L.nn; L.xn = p; p += L.nn; L.b = p; p += L.nn;
    }
    H->nlev = nlev;
    for (int l = 1; l < nlev; l++) {
        if (l == 1) k_mg_links_from_types<<<nblk(H->L[1].nn, 256), 256, 0, c->stream>>>(s, c->node_type, H->L[1]);
        else k_mg_links_from_links<<<nblk(H->L[l].nn, 256), 256, 0, c->stream>>>(H->L[l - 1], H->L[l]);
        LAUNCH_CHECK(c);
    }
    H->geom_version = c->geom_version;
    return 0;
}

// Newton + multigrid-preconditioned CG (same outer iteration as solve_nrpcg_spd)
static int solve_nrpcg_mg(espic_ctx *c, const espic_solve_params *p, espic_solve_info *info)
{
    int r;
    if ((r = ensure_node_types(c, 0))) return r;
    if ((r = ensure_sv(c, 8))) return r;
    StencilC s = make_stencil(c->m);
    MgHierarchy *H = g_mg_of(c);
    if ((r = mg_setup(c, s, H, nullptr))) return r;
    double *diag0 = c->sv[0], *R = c->sv[1], *diagJ = c->sv[2], *minv = c->sv[3];
    double *delta = c->sv[4], *z = c->sv[5], *d0 = c->sv[6], *d1 = c->sv[7];
    // two more fine vectors (q, x0) + the partial sums live in the reduction scratch
    if ((r = ensure_buf(&c->red, &c->red_cap, 2 * s.nn + 8192, c->stream))) return r;
    double *q = c->red, *x0 = c->red + s.nn, *part = c->red + 2 * s.nn;
    static int coarse_sweeps = -1;
    if (coarse_sweeps < 0) {
        const char *ev = getenv("ESPIC_MG_COARSE_SWEEPS");
        coarse_sweeps = ev ? std::max(0, atoi(ev)) : MG_COARSE_SWEEPS;
        coarse_sweeps += coarse_sweeps & 1;
    }
    if (c->diag0_version != c->geom_version) {
        k_spd_diag0<<<nblk(s.nn, 256), 256, 0, c->stream>>>(s, c->node_type, diag0);
        LAUNCH_CHECK(c);
        c->diag0_version = c->geom_version;
    }
    int bps = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_mg_pcg, MG_BLOCK, 0));
    if (bps < 1) { espic_set_error("k_mg_pcg cannot be made resident"); return -1; }
    long long want = (s.nn + MG_BLOCK - 1) / MG_BLOCK;
    int grid = (int)std::min<long long>((long long)bps * c->sm_count, std::max<long long>(want, 1));
    if (3 * grid > 4096) grid = 4096 / 3;
    const int nb_res = std::min<long long>(nblk(s.nn, 256), 1024);
    double *dout = reinterpret_cast<double *>(c->dscal + 24);
    double *dres = reinterpret_cast<double *>(c->dscal + 16);
    double norm = 0, r0_norm = 0, lin_stop = 0;
    bool converged = false;
    static const bool inexact = getenv("ESPIC_MG_EXACT_NEWTON") == nullptr;
    static const double forcing = getenv("ESPIC_MG_FORCING") ? atof(getenv("ESPIC_MG_FORCING")) : 1.0;
    for (int it = 0; it < p->nr_max_it; it++) {
        info->nr_iters++;
        k_spd_linearise<<<nblk(s.nn, 256), 256, 0, c->stream>>>(s, c->node_type, c->rho, c->phi, diag0, p->phi0, p->Te0, p->n0,
                                                                R, diagJ, minv, delta);
        LAUNCH_CHECK(c);
        // Galerkin diagonals: the Boltzmann term changes with phi, the links do not
        for (int l = 1; l < H->nlev; l++) {
            if (l == 1) k_mg_diag_from_fine<<<nblk(H->L[1].nn, 256), 256, 0, c->stream>>>(s, c->node_type, diagJ, H->L[1]);
            else k_mg_diag_from_level<<<nblk(H->L[l].nn, 256), 256, 0, c->stream>>>(H->L[l - 1], H->L[l]);
            LAUNCH_CHECK(c);
        }
        MgPcgArgs a;
        a.s = s; a.nlev = H->nlev; a.coarse_sweeps = coarse_sweeps; a.nbmask = H->nbmask; a.first_redundant = H->nlev - 1;
        a.prof = getenv("ESPIC_MG_PROFILE") ? c->dscal + 40 : nullptr;
        for (int l = 0; l < H->nlev; l++) a.L[l] = H->L[l];
        a.L[0].diag = diagJ; a.L[0].minv = minv; a.L[0].x = x0;
        a.delta = delta; a.r = R; a.z = z; a.d0 = d0; a.d1 = d1; a.q = q;
        a.part = part; a.max_it = p->max_it; a.tol = p->tol; a.out = dout;
        // Newton step 0 leaves a nonlinear residual of about newton_ratio * |R0| whatever the accuracy of its linear solve:
        // its linear solve stops at that level (measured on the bench case: factor 0.1 / 0.3 / 0.6 / 1.0 -> 58.5 / 56.4 /
        // 54.6 / 53.7 CG iterations per step, Newton count unchanged); every later step, and a first step without
        // history, is solved to tol, and convergence is only declared after a solve that went to tol.
        a.rel_tol = (it == 0 && inexact) ? forcing * std::min(std::max(H->newton_ratio, 0.0), 1e-2) : 0.0;
        CK(cudaMemsetAsync(d0, 0, (size_t)s.nn * sizeof(double), c->stream));     // beta = 0 in the first iteration must meet finite numbers
        void *args[] = {&a};
        CK(cudaLaunchCooperativeKernel((void *)k_mg_pcg, dim3(grid), dim3(MG_BLOCK), args, 0, c->stream));
        LAUNCH_CHECK(c);
        k_spd_update<<<nb_res, 256, 0, c->stream>>>(s, c->node_type, delta, c->phi, part);
        LAUNCH_CHECK(c);
        k_sum_final<<<1, 256, 0, c->stream>>>(part, nb_res, dres);
        LAUNCH_CHECK(c);
        double *h = reinterpret_cast<double *>(c->hpin) + 24;
        CK(cudaMemcpyAsync(h, dout, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        double sum;
        if ((r = read_scalar(c, dres, &sum))) return r;
        info->lin_iters += (long long)h[1];
        if (h[0] == 0.0) fprintf(stderr, "PCG failed to converge, norm(g) = %g\n", h[2]);
        norm = sqrt(sum / (double)s.nn);
        if (it == 0) r0_norm = h[3];
        if (it == 1 && r0_norm > 0) H->newton_ratio = h[3] / r0_norm;
        lin_stop = (a.rel_tol > 0) ? std::max(p->tol, a.rel_tol * h[3]) : p->tol;
        if (getenv("ESPIC_MG_PROFILE"))
            fprintf(stderr, "[mg newton %d] |R| %.3e -> %.3e in %d its, |y| = %.3e\n", it, h[3], h[2], (int)h[1], norm);
        // converged as the reference defines it (update below nr_tol) -- but only after a linear solve that went to tol
        if (norm < p->nr_tol && lin_stop <= p->tol) { converged = true; break; }
    }
    for (int level = 0; level < 3; level++) {
        k_mirror<<<nblk(s.nn, 256), 256, 0, c->stream>>>(s, c->node_type, c->phi, level);
        LAUNCH_CHECK(c);
    }
    if (!converged) printf("NR+PCG failed to converge, norm = %g\n", norm);
    info->converged = converged;
    info->residual = norm;
    if (getenv("ESPIC_MG_PROFILE")) {
        unsigned long long hp[8];
        CK(cudaMemcpyAsync(hp, c->dscal + 40, sizeof(hp), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaMemsetAsync(c->dscal + 40, 0, sizeof(hp), c->stream));
        fprintf(stderr, "[mg profile] its=%lld  us/it: down0 %.1f  down %.1f  coarsest %.1f  up %.1f  up0+rz %.1f  dq %.1f  r %.1f\n",
                info->lin_iters, hp[0] * 1e-3 / info->lin_iters, hp[1] * 1e-3 / info->lin_iters, hp[2] * 1e-3 / info->lin_iters,
                hp[3] * 1e-3 / info->lin_iters, hp[4] * 1e-3 / info->lin_iters, hp[5] * 1e-3 / info->lin_iters, hp[6] * 1e-3 / info->lin_iters);
    }
    return 0;
}


// ====================================================================================================================
// Slab-decomposed variant (ESPIC_SOLVE_PCG_MG_SLAB): one k-slab per rank, all traffic through peer-mapped memory
// ====================================================================================================================

struct SlabState {
    bool ready = false, diag0_done = false;
    long long geom_version = -1;
    double *pool = nullptr;            // this rank's pool; every rank carves it identically
    size_t pool_doubles = 0;
    void *peer_base[MG_MAX_RANKS] = {nullptr};
    // carved arrays (fine vectors), the coarse-level arrays, partial sums, flags
    double *diag0, *R, *diagJ, *minv, *delta, *z, *d0, *d1, *q, *x0, *coarse, *part;
    unsigned long long *flags, *epoch, *arrive, *release;
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
    cudaFree(S->H.nbmask);
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
    // ---- pool: identical carving on every rank
    const long long coarse = mg_coarse_doubles(s);
    const long long nparts = 3ll * c->nranks * 4096;
    S->pool_doubles = (size_t)(10 * s.nn + coarse + nparts + 128);
    CK(cudaMalloc(&S->pool, S->pool_doubles * sizeof(double)));
    CK(cudaMemsetAsync(S->pool, 0, S->pool_doubles * sizeof(double), c->stream));
    double *p = S->pool;
    S->diag0 = p; p += s.nn; S->R = p; p += s.nn; S->diagJ = p; p += s.nn; S->minv = p; p += s.nn; S->delta = p; p += s.nn;
    S->z = p; p += s.nn; S->d0 = p; p += s.nn; S->d1 = p; p += s.nn; S->q = p; p += s.nn; S->x0 = p; p += s.nn;
    S->coarse = p; p += coarse; S->part = p; p += nparts;
    S->flags = reinterpret_cast<unsigned long long *>(p); p += 16;
    S->epoch = reinterpret_cast<unsigned long long *>(p); p += 16;
    S->arrive = reinterpret_cast<unsigned long long *>(p); p += 16;
    S->release = reinterpret_cast<unsigned long long *>(p); p += 16;
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
    own.arrive = S->arrive; own.release = S->release;
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
    const long long slab_nn = (long long)(s.nk / c->nranks) * s.sk;
    if (!S->diag0_done) {
        k_spd_diag0<<<nblk(s.nn, 256), 256, 0, c->stream>>>(s, c->node_type, S->diag0);
        LAUNCH_CHECK(c);
        S->diag0_done = true;
    }
    int bps = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_mg_pcg_slab, MG_BLOCK, 0));
    if (bps < 1) { espic_set_error("k_mg_pcg_slab cannot be made resident"); return -1; }
    // the same grid on every rank (the partial-sum layout depends on it): all blocks the device can hold
    int grid = bps * c->sm_count;
    if (grid > 4096) grid = 4096;
    const int nb_res = std::min<long long>(nblk(s.nn, 256), 1024);
    if ((r = ensure_buf(&c->red, &c->red_cap, 8192, c->stream))) return r;
    double *dout = reinterpret_cast<double *>(c->dscal + 24);
    double *dres = reinterpret_cast<double *>(c->dscal + 16);
    double norm = 0, r0_norm = 0;
    bool converged = false;
    for (int it = 0; it < p->nr_max_it; it++) {
        info->nr_iters++;
        // the Newton linearisation and the Galerkin diagonals are computed by every rank for the whole mesh (cheap, pointwise)
        k_spd_linearise<<<nblk(s.nn, 256), 256, 0, c->stream>>>(s, c->node_type, c->rho, c->phi, S->diag0, p->phi0, p->Te0, p->n0,
                                                                S->R, S->diagJ, S->minv, S->delta);
        LAUNCH_CHECK(c);
        for (int l = 1; l < H->nlev; l++) {
            if (l == 1) k_mg_diag_from_fine<<<nblk(H->L[1].nn, 256), 256, 0, c->stream>>>(s, c->node_type, S->diagJ, H->L[1]);
            else k_mg_diag_from_level<<<nblk(H->L[l].nn, 256), 256, 0, c->stream>>>(H->L[l - 1], H->L[l]);
            LAUNCH_CHECK(c);
        }
        CK(cudaMemsetAsync(S->d0, 0, (size_t)s.nn * sizeof(double), c->stream));
        MgPcgArgs a;
        a.s = s; a.nlev = H->nlev; a.coarse_sweeps = MG_COARSE_SWEEPS; a.nbmask = H->nbmask;
        // Levels with at most MG_SLAB_REDUNDANT_NODES nodes (override: ESPIC_MG_SLAB_REDUNDANT_NODES, 0 = only the coarsest)
        // are solved by every rank in full instead of by slabs: 2 inter-GPU barriers less per level and V-cycle for redundant
        // work on a small level.  Measured on 4 B200, 256^3 (profiles/r1_slab_redundant_levels_n4.txt): 530 us per CG
        // iteration with only the coarsest level redundant, 508 with <= 8192 nodes, 493 with <= 65536, 492 with <= 524288.
        {
            long long dims[MG_MAX_LEVELS][3];
            for (int l = 0; l < H->nlev; l++) { dims[l][0] = H->L[l].ni; dims[l][1] = H->L[l].nj; dims[l][2] = H->L[l].nk; }
            a.first_redundant = mg_first_redundant(H->nlev, dims, mg_slab_redundant_limit());
        }
        for (int l = 0; l < H->nlev; l++) a.L[l] = H->L[l];
        a.L[0].diag = S->diagJ; a.L[0].minv = S->minv; a.L[0].x = S->x0;
        a.delta = S->delta; a.r = S->R; a.z = S->z; a.d0 = S->d0; a.d1 = S->d1; a.q = S->q;
        a.part = S->part; a.max_it = p->max_it; a.tol = p->tol; a.out = dout;
        a.prof = getenv("ESPIC_MG_PROFILE") ? c->dscal + 40 : nullptr;
        // same inexact-Newton forcing as the single-GPU solver (identical on every rank: it derives from all-reduced norms)
        a.rel_tol = (it == 0 && getenv("ESPIC_MG_EXACT_NEWTON") == nullptr) ? std::min(std::max(H->newton_ratio, 0.0), 1e-2) : 0.0;
        // nobody may store into a neighbour's pool before that neighbour has finished preparing this Newton step
        if ((r = espic_comm_allgather_doubles(c, S->part, 1))) return r;
        OwnSlab own = S->own;
        void *args[] = {&a, &own, &S->epoch};
        CK(cudaLaunchCooperativeKernel((void *)k_mg_pcg_slab, dim3(grid), dim3(MG_BLOCK), args, 0, c->stream));
        LAUNCH_CHECK(c);
        // every rank needs the whole update: gather the slabs of delta
        if ((r = espic_comm_allgather_doubles(c, S->delta, (size_t)slab_nn))) return r;
        k_spd_update<<<nb_res, 256, 0, c->stream>>>(s, c->node_type, S->delta, c->phi, c->red);
        LAUNCH_CHECK(c);
        k_sum_final<<<1, 256, 0, c->stream>>>(c->red, nb_res, dres);
        LAUNCH_CHECK(c);
        double *h = reinterpret_cast<double *>(c->hpin) + 24;
        CK(cudaMemcpyAsync(h, dout, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        double sum;
        if ((r = read_scalar(c, dres, &sum))) return r;
        info->lin_iters += (long long)h[1];
        if (h[0] == 0.0) fprintf(stderr, "PCG failed to converge, norm(g) = %g\n", h[2]);
        norm = sqrt(sum / (double)s.nn);
        if (it == 0) r0_norm = h[3];
        if (it == 1 && r0_norm > 0) H->newton_ratio = h[3] / r0_norm;
        const double lin_stop = (a.rel_tol > 0) ? std::max(p->tol, a.rel_tol * h[3]) : p->tol;
        if (norm < p->nr_tol && lin_stop <= p->tol) { converged = true; break; }
    }
    for (int level = 0; level < 3; level++) {
        k_mirror<<<nblk(s.nn, 256), 256, 0, c->stream>>>(s, c->node_type, c->phi, level);
        LAUNCH_CHECK(c);
    }
    if (!converged) printf("NR+PCG failed to converge, norm = %g\n", norm);
    info->converged = converged;
    info->residual = norm;
    if (getenv("ESPIC_MG_PROFILE") && c->rank == 0) {
        unsigned long long hp[8];
        CK(cudaMemcpyAsync(hp, c->dscal + 40, sizeof(hp), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaMemsetAsync(c->dscal + 40, 0, sizeof(hp), c->stream));
        fprintf(stderr, "[mg slab profile] its=%lld  us/it: down0 %.1f  down %.1f  coarsest %.1f  up %.1f  up0+rz %.1f  dq %.1f  r %.1f\n",
                info->lin_iters, hp[0] * 1e-3 / info->lin_iters, hp[1] * 1e-3 / info->lin_iters, hp[2] * 1e-3 / info->lin_iters,
                hp[3] * 1e-3 / info->lin_iters, hp[4] * 1e-3 / info->lin_iters, hp[5] * 1e-3 / info->lin_iters, hp[6] * 1e-3 / info->lin_iters);
    }
    return 0;
}
