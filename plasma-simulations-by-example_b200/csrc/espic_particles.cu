// espic_particles.cu -- Species::advance / computeNumberDensity / addParticle / ColdBeamSource::sample /
// diagnostics as sm_100a kernels over SoA particle arrays (include/espic.h).
#include "espic_internal.cuh"
#include <algorithm>
#include <chrono>
#include <stdlib.h>
#include <math.h>

#define SP_CHECK(c, sp) do { if ((sp) < 0 || (sp) >= (c)->nsp) { espic_set_error("bad species id %d", (sp)); return -1; } } while (0)
static inline unsigned nblk(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// =====================================================================================================
// exclusive scan of uint32 (two level, chunk = 8192 elements)
// =====================================================================================================

#define SCAN_CHUNK (1 << SCAN_CHUNK_LOG2)

__global__ void __launch_bounds__(256) k_scan_l1(const uint32_t *__restrict__ in, long long n,
                                                 uint32_t *__restrict__ pre, uint32_t *__restrict__ ctot)
{
    __shared__ uint32_t wsum[8];
    const long long base = (long long)blockIdx.x * SCAN_CHUNK + (long long)threadIdx.x * 32;
    uint32_t v[32];
    uint32_t tsum = 0;
#pragma unroll
    for (int q = 0; q < 32; q++) {
        long long i = base + q;
        v[q] = (i < n) ? in[i] : 0u;
        tsum += v[q];
    }
    // block exclusive scan of tsum
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int q = 0; q < w; q++) woff += wsum[q];
    uint32_t run = woff + inc - tsum;
#pragma unroll
    for (int q = 0; q < 32; q++) {
        long long i = base + q;
        if (i < n) pre[i] = run;
        run += v[q];
    }
    if (threadIdx.x == 255) ctot[blockIdx.x] = woff + inc;
}

__global__ void __launch_bounds__(1024) k_scan_l2(uint32_t *__restrict__ ctot, long long nchunks, unsigned long long *total)
{
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (long long b = 0; b < nchunks; b += 1024) {
        long long i = b + threadIdx.x;
        unsigned long long v = (i < nchunks) ? ctot[i] : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        unsigned long long woff = 0;
        for (int q = 0; q < w; q++) woff += wsum[q];
        unsigned long long carry = carry_s;
        if (i < nchunks) ctot[i] = (uint32_t)(carry + woff + inc - v);
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + woff + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

int espic_scan_u32(espic_ctx *c, const uint32_t *in, long long n, unsigned long long *d_total)
{
    long long nchunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (nchunks < 1) nchunks = 1;
    int r;
    if ((r = ensure_buf(&c->scan_pre, &c->scan_cap, n, c->stream))) return r;
    if ((r = ensure_buf(&c->scan_coff, &c->scan_coff_cap, nchunks, c->stream))) return r;
    k_scan_l1<<<(unsigned)nchunks, 256, 0, c->stream>>>(in, n, c->scan_pre, c->scan_coff);
    LAUNCH_CHECK(c);
    k_scan_l2<<<1, 1024, 0, c->stream>>>(c->scan_coff, nchunks, d_total);
    LAUNCH_CHECK(c);
    return 0;
}

__device__ __forceinline__ unsigned long long scan_at(const uint32_t *__restrict__ pre, const uint32_t *__restrict__ coff, long long w)
{
    return (unsigned long long)pre[w] + coff[w >> SCAN_CHUNK_LOG2];
}

// =====================================================================================================
// scatter (Field::scatter, Field.h:167-186) into an accumulator
//
// Particles are kept (approximately) sorted by cell, so neighbouring lanes mostly deposit on the same eight nodes.
// Instead of 8 atomics per particle the warp first adds up, in registers and with shuffles, every run of consecutive
// lanes that sit in the same cell (segmented reduction) and only the first lane of each run issues the 8 atomics.
// FP64 mode: the additions happen in a different order than the reference's particle loop (rounding only).
// Fixed-point mode: every particle's eight weights are rounded to int64 multiples of 2^-shift FIRST, all later
// additions are integer and therefore exact: the result does not depend on particle order, run lengths or GPU count.
// =====================================================================================================

template <int MODE> struct AccVal;
template <> struct AccVal<ESPIC_DEPOSIT_FP64> {
    typedef double T;
    static __device__ __forceinline__ T quant(double w, double) { return w; }
    static __device__ __forceinline__ void red(double *acc, long long u, T v) { atomicAdd(acc + u, v); }
};
template <> struct AccVal<ESPIC_DEPOSIT_FIXED> {
    typedef long long T;
    static __device__ __forceinline__ T quant(double w, double scale) { return __double2ll_rn(w * scale); }
    static __device__ __forceinline__ void red(double *acc, long long u, T v)
    {
        atomicAdd(reinterpret_cast<unsigned long long *>(acc) + u, (unsigned long long)v);
    }
};

__device__ __forceinline__ void cell3(const MeshC &m, double x, double y, double z, int &i, int &j, int &k,
                                      double &di, double &dj, double &dk)
{
    cell_frac(x, m.x0[0], m.dh[0], m.rdh[0], m.ni, i, di);
    cell_frac(y, m.x0[1], m.dh[1], m.rdh[1], m.nj, j, dj);
    cell_frac(z, m.x0[2], m.dh[2], m.rdh[2], m.nk, k, dk);
}

// the eight node weights in the reference's node order, each mpw*w_i*w_j*w_k multiplied left to right (Field.h:177-184),
// from the cell fractions
template <int MODE>
__device__ __forceinline__ void weights_from_fractions(double di, double dj, double dk, double mpw, double scale,
                                                       typename AccVal<MODE>::T w[8])
{
    const double ai = 1 - di, aj = 1 - dj, ak = 1 - dk;
    w[0] = AccVal<MODE>::quant(mpw * ai * aj * ak, scale);
    w[1] = AccVal<MODE>::quant(mpw * di * aj * ak, scale);
    w[2] = AccVal<MODE>::quant(mpw * di * dj * ak, scale);
    w[3] = AccVal<MODE>::quant(mpw * ai * dj * ak, scale);
    w[4] = AccVal<MODE>::quant(mpw * ai * aj * dk, scale);
    w[5] = AccVal<MODE>::quant(mpw * di * aj * dk, scale);
    w[6] = AccVal<MODE>::quant(mpw * di * dj * dk, scale);
    w[7] = AccVal<MODE>::quant(mpw * ai * dj * dk, scale);
}

template <int MODE>
__device__ __forceinline__ long long particle_weights(const MeshC &m, double x, double y, double z, double mpw, double scale,
                                                      typename AccVal<MODE>::T w[8])
{
    int i, j, k; double di, dj, dk;
    cell3(m, x, y, z, i, j, k, di, dj, dk);
    if (i < 0 || j < 0 || k < 0) return -1;     // never for in-bounds particles; keeps stray input from writing out of range
    weights_from_fractions<MODE>(di, dj, dk, mpw, scale, w);
    return node_u(m, i, j, k);
}

// STRIDE: distance (in elements) between consecutive nodes of the target array (3 for one component of an interleaved
// vector field such as nv_sum, whose component offset is already added to acc)
template <int MODE, int STRIDE = 1>
__device__ __forceinline__ void red8(const MeshC &m, double *acc, long long u, const typename AccVal<MODE>::T w[8])
{
    // four row bases, the +1 neighbours are immediate offsets
    double *p00 = acc + u * STRIDE, *p10 = p00 + (long long)m.ni * STRIDE, *p01 = p00 + (long long)m.ni * m.nj * STRIDE,
           *p11 = p01 + (long long)m.ni * STRIDE;
    AccVal<MODE>::red(p00, 0, w[0]);
    AccVal<MODE>::red(p00, STRIDE, w[1]);
    AccVal<MODE>::red(p10, STRIDE, w[2]);
    AccVal<MODE>::red(p10, 0, w[3]);
    AccVal<MODE>::red(p01, 0, w[4]);
    AccVal<MODE>::red(p01, STRIDE, w[5]);
    AccVal<MODE>::red(p11, STRIDE, w[6]);
    AccVal<MODE>::red(p11, 0, w[7]);
}

// Whole-warp call.  u = lower node of the lane's cell (-1: nothing to deposit), w = its eight weights.
template <int MODE, int STRIDE = 1>
__device__ __forceinline__ void warp_deposit(const MeshC &m, double *acc, long long u, typename AccVal<MODE>::T w[8])
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long up = __shfl_up_sync(FULL, u, 1);
    const unsigned heads = __ballot_sync(FULL, lane == 0 || up != u);
    if (heads == FULL) {                   // no two neighbours share a cell: nothing to combine
        if (u >= 0) red8<MODE, STRIDE>(m, acc, u, w);
        return;
    }
    // last lane of this lane's run = lane before the next head
    const unsigned above = heads & ~((2u << lane) - 1u);       // heads strictly above this lane (lane 31: 2u<<31 == 0 -> mask 0xffffffff)
    const int run_end = (lane == 31 || above == 0) ? 31 : (__ffs(above) - 2);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const bool take = lane + o <= run_end;
        if (!__any_sync(FULL, take)) break;        // no run of this warp is longer than o lanes: the remaining rounds add nothing
#pragma unroll
        for (int q = 0; q < 8; q++) {
            typename AccVal<MODE>::T t = __shfl_down_sync(FULL, w[q], o);
            if (take) w[q] += t;
        }
    }
    if (((heads >> lane) & 1u) && u >= 0) red8<MODE, STRIDE>(m, acc, u, w);
}

// two particles per thread (adjacent in memory: one 128-bit load per array), merged in registers when they share a cell
template <int MODE>
__device__ __forceinline__ void deposit_pair(const MeshC &m, double *acc, double scale, bool live0, double x0, double y0, double z0,
                                             double w0, bool live1, double x1, double y1, double z1, double w1)
{
    typename AccVal<MODE>::T a[8], b[8];
    long long ua = -1, ub = -1;
    if (live0) ua = particle_weights<MODE>(m, x0, y0, z0, w0, scale, a);
    if (live1) ub = particle_weights<MODE>(m, x1, y1, z1, w1, scale, b);
    if (ub >= 0) {
        if (ua == ub) {
#pragma unroll
            for (int q = 0; q < 8; q++) a[q] += b[q];
        } else if (ua < 0) {
            ua = ub;
#pragma unroll
            for (int q = 0; q < 8; q++) a[q] = b[q];
        } else {
            red8<MODE>(m, acc, ub, b);     // the pair straddles two cells: the second one goes out on its own
        }
    }
    if (ua < 0) {
#pragma unroll
        for (int q = 0; q < 8; q++) a[q] = 0;
    }
    warp_deposit<MODE>(m, acc, ua, a);
}

// L2 prefetch (UBLKPF.L2, sm_90+) of `bytes` (multiple of 16) starting at p: issued for the tile a block `ahead` blocks
// later will stream, so that block's loads hit L2 instead of paying the full HBM latency (k_push: 72 % -> 86 % of the
// measured HBM peak at a distance of 1-3 blocks per SM; 0.5 and 27 blocks per SM are both worse)
__device__ __forceinline__ void l2_prefetch(const void *p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ double2 ld2(const double *p) { return __ldcs(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ void st2(double *p, double a, double b) { __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b)); }

template <int MODE>
__global__ void __launch_bounds__(256) k_deposit(MeshC m, const double *__restrict__ x, const double *__restrict__ y,
                                                 const double *__restrict__ z, const double *__restrict__ mpw,
                                                 long long n, double *acc, double scale)
{
    const long long i0 = 2 * (blockIdx.x * 256ll + threadIdx.x);
    if (i0 - 2 * (threadIdx.x & 31) >= n) return;            // whole warp past the end
    const bool v0 = i0 < n, v1 = i0 + 1 < n;
    double2 X = make_double2(0, 0), Y = X, Z = X, W = X;
    if (v0) { X = ld2(x + i0); Y = ld2(y + i0); Z = ld2(z + i0); W = ld2(mpw + i0); }   // capacity is even: slot i0+1 is allocated
    deposit_pair<MODE>(m, acc, scale, v0, X.x, Y.x, Z.x, W.x, v1, X.y, Y.y, Z.y, W.y);
}

// ---- tiled deposit: group the particles of a tile by cell in shared memory first ------------------------------------
// A few steps after a cell sort the particle stream is only roughly ordered: runs of equal cells shrink to 2-4 particles
// and the warp-level run merge above degenerates into one RED per particle and node (measured: 2.0 ms right after a sort,
// 5.7 ms three steps later at 2e8 particles; MIO-queue bound).  This kernel restores the grouping locally, per tile of
// DT_TILE consecutive particles, without moving any particle in global memory:
//   1. coalesced load of x,y,z,mpw into shared memory; each particle's cell is entered into a small open-addressing hash
//      table (32-bit shared atomics) which hands out a slot per distinct cell; slots are counted;
//   2. exclusive scan of the slot counts, then every particle id is written to its place in slot order (counting sort);
//   3. each thread takes DT_PER consecutive entries of that order -- now runs of one cell -- forms the weights from the
//      cell fractions step 1 left in shared memory (the cell itself is the slot's key), merges
//      them in registers, the warp merges runs across lanes with shuffles, run heads issue the REDs.
// REDs per tile drop from (#runs x 8) to about (#distinct cells x 8), independent of how scrambled the stream is.
#define DT_THREADS 256
#define DT_PER 4
#define DT_TILE (DT_THREADS * DT_PER)
#define DT_SLOTS 1024                     // hash table entries (a tile of 1024 particles cannot touch more cells)
#define DT_SLOT_BITS 10
#define DT_EMPTY 0xffffffffu
static_assert(DT_SLOTS >= DT_TILE && DT_SLOTS == (1 << DT_SLOT_BITS), "one hash slot per particle of the tile");

// VAL selects what is scattered: 0 the weight mpw (number density), 1 mpw*v, 2 (mpw*v)*v with v = vcomp[] (velocity moments,
// ch4 Species::sampleMoments: the same kernel runs once per sampled quantity)
// GROUP = false skips steps 1b and 2 (hash, scan, placement): each thread then merges DT_PER CONSECUTIVE particles of the
// stream.  That is the right kernel while the stream is still in cell order (runs of ~100 particles per cell): it needs a
// quarter of the warp shuffles of k_deposit (one segmented reduction per 4 particles instead of per 2) and no hashing.

// shared-memory working set of one tile (45 KB)
struct DepTile {
    double sx[DT_TILE], sy[DT_TILE], sz[DT_TILE], sw[DT_TILE];      // cell fractions and weight of every particle of the tile
    uint32_t hkey[DT_SLOTS], hcnt[DT_SLOTS];                        // hash table: lower node of the cell; slot counts, then offsets
    uint16_t pslot[DT_TILE], order[DT_TILE];
    uint32_t wsum[DT_THREADS / 32];
    uint32_t n_sorted;
};

__device__ __forceinline__ void dt_clear(DepTile &t)
{
#pragma unroll
    for (int q = threadIdx.x; q < DT_SLOTS; q += DT_THREADS) { t.hkey[q] = DT_EMPTY; t.hcnt[q] = 0; }
}

// Step 1 for one particle per lane (whole-warp call): entry p of the tile is the particle at (px,py,pz) with weight pw
// (0 = nothing to deposit).  Lanes with the same cell elect one leader that talks to the hash table (a freshly sorted tile
// would otherwise send 32 CAS + 32 adds to the same shared-memory word).
template <bool GROUP>
__device__ __forceinline__ void dt_insert(const MeshC &m, DepTile &t, int p, double px, double py, double pz, double pw)
{
    const int lane = threadIdx.x & 31;
    uint32_t key = DT_EMPTY;
    double d0 = 0, d1 = 0, d2 = 0;
    if (pw != 0) {
        int ci, cj, ck;
        cell3(m, px, py, pz, ci, cj, ck, d0, d1, d2);
        if (ci >= 0 && cj >= 0 && ck >= 0) key = (uint32_t)node_u(m, ci, cj, ck);
    }
    // only the fractions are needed from here on; the cell is recovered from the hash slot (hkey[slot] = lower node of the cell)
    t.sx[p] = d0; t.sy[p] = d1; t.sz[p] = d2; t.sw[p] = pw;
    if (!GROUP) {           // no grouping: remember the cell's lower node per particle (hkey doubles as that array)
        t.hkey[p] = key;
        return;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    uint32_t slot = 0xffffu;          // 0xffff: nothing to deposit (the table has a slot per particle of the tile, it cannot fill up)
    if (key != DT_EMPTY && lane == leader) {
        uint32_t h = (key * 2654435761u) >> (32 - DT_SLOT_BITS);
        for (int probe = 0; probe < DT_SLOTS; probe++) {
            const uint32_t old = atomicCAS(&t.hkey[h], DT_EMPTY, key);
            if (old == DT_EMPTY || old == key) { slot = h; break; }
            h = (h + 1) & (DT_SLOTS - 1);
        }
        if (slot < DT_SLOTS) atomicAdd(&t.hcnt[slot], (uint32_t)__popc(peers));
    }
    slot = __shfl_sync(0xffffffffu, slot, leader);
    t.pslot[p] = (uint16_t)slot;
}

// Steps 2 and 3 (whole-block call, after a __syncthreads behind the last dt_insert): counting sort of the particle ids by hash
// slot, then every thread merges DT_PER consecutive entries of that order, the warp merges runs, run heads issue the REDs.
template <int MODE, int STRIDE, bool GROUP>
__device__ __forceinline__ void dt_finish(const MeshC &m, DepTile &t, double *acc, double scale)
{
    typedef typename AccVal<MODE>::T T;
    const int tid = threadIdx.x, lane = tid & 31;
    if (GROUP) {
        {
            constexpr int PER = DT_SLOTS / DT_THREADS;
            uint32_t cnt[PER], tsum = 0;
#pragma unroll
            for (int q = 0; q < PER; q++) { cnt[q] = t.hcnt[tid * PER + q]; tsum += cnt[q]; }
            uint32_t inc = tsum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane == 31) t.wsum[tid >> 5] = inc;
            __syncthreads();
            uint32_t run = inc - tsum;
            for (int w = 0; w < (tid >> 5); w++) run += t.wsum[w];
#pragma unroll
            for (int q = 0; q < PER; q++) { t.hcnt[tid * PER + q] = run; run += cnt[q]; }
            if (tid == DT_THREADS - 1) t.n_sorted = run;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < DT_PER; j++) {
            const int p = j * DT_THREADS + tid;
            const uint32_t slot = t.pslot[p];
            const unsigned peers = __match_any_sync(0xffffffffu, slot);
            const int leader = __ffs(peers) - 1;
            uint32_t at = 0;
            if (slot < DT_SLOTS && lane == leader) at = atomicAdd(&t.hcnt[slot], (uint32_t)__popc(peers));
            at = __shfl_sync(0xffffffffu, at, leader);
            if (slot < DT_SLOTS) t.order[at + __popc(peers & ((1u << lane) - 1u))] = (uint16_t)p;
        }
        __syncthreads();
    }
    const int M = GROUP ? (int)t.n_sorted : DT_TILE;
    T a[8];
#pragma unroll
    for (int q = 0; q < 8; q++) a[q] = 0;
    long long ua = -1;
#pragma unroll
    for (int j = 0; j < DT_PER; j++) {
        const int q = tid * DT_PER + j;
        if (q >= M) break;
        const int p = GROUP ? t.order[q] : q;
        T v[8];
        const uint32_t cellu = GROUP ? t.hkey[t.pslot[p]] : t.hkey[p];
        if (!GROUP && cellu == DT_EMPTY) continue;                // nothing to deposit (dead slot, tail of the last tile)
        const long long u = (long long)cellu;
        weights_from_fractions<MODE>(t.sx[p], t.sy[p], t.sz[p], t.sw[p], scale, v);
        if (u != ua) {
            if (ua >= 0) red8<MODE, STRIDE>(m, acc, ua, a);       // rare: a cell boundary inside this thread's four entries
            ua = u;
#pragma unroll
            for (int c = 0; c < 8; c++) a[c] = v[c];
        } else {
#pragma unroll
            for (int c = 0; c < 8; c++) a[c] += v[c];
        }
    }
    warp_deposit<MODE, STRIDE>(m, acc, ua, a);
}

template <int MODE, int VAL = 0, int STRIDE = 1, bool GROUP = true>
__global__ void __launch_bounds__(DT_THREADS) k_deposit_tile(MeshC m, const double *__restrict__ x, const double *__restrict__ y,
                                                              const double *__restrict__ z, const double *__restrict__ mpw,
                                                              long long n, double *acc, double scale, int ahead,
                                                              const double *__restrict__ vcomp = nullptr)
{
    __shared__ DepTile t;
    const int tid = threadIdx.x;
    const long long base = blockIdx.x * (long long)DT_TILE;
    if (tid < 4) {
        const long long pf = base + (long long)ahead * DT_TILE;
        if (ahead > 0 && pf + DT_TILE <= n) l2_prefetch((tid == 0 ? x : tid == 1 ? y : tid == 2 ? z : mpw) + pf, DT_TILE * 8);
    }
    dt_clear(t);
    __syncthreads();
    // ---- 1. load, hash, count
#pragma unroll
    for (int j = 0; j < DT_PER / 2; j++) {
        const int p = 2 * (j * DT_THREADS + tid);
        const long long g = base + p;
        double2 X = make_double2(0, 0), Y = X, Z = X, W = X;
        if (g < n) { X = ld2(x + g); Y = ld2(y + g); Z = ld2(z + g); W = ld2(mpw + g); }     // capacity is even: g+1 is allocated
        if (VAL > 0 && g < n) {
            const double2 V = ld2(vcomp + g);
            W.x = W.x * V.x; W.y = W.y * V.y;                         // mpw*v
            if (VAL == 2) { W.x = W.x * V.x; W.y = W.y * V.y; }       // (mpw*v)*v
        }
        if (g + 1 >= n) W.y = 0;
        dt_insert<GROUP>(m, t, p, X.x, Y.x, Z.x, W.x);
        dt_insert<GROUP>(m, t, p + 1, X.y, Y.y, Z.y, W.y);
    }
    __syncthreads();
    dt_finish<MODE, STRIDE, GROUP>(m, t, acc, scale);
}

// ---- grouped deposit, second version ---------------------------------------------------------------------------------
// k_deposit_tile above spends its time in the L1/shared-memory pipe (ncu, profiles/r2_ncu_full_push_deposit_2e8.txt: 89 % busy;
// 2900 shared-memory wavefronts per tile, a third of them bank conflicts: the staged fractions are written with a two-particle
// stride and read back through the sort permutation).  Same algorithm, half the shared-memory traffic:
//   1. every thread loads FOUR CONSECUTIVE particles (one 256-bit load per array) and keeps their cell fractions in REGISTERS;
//      the hash insertion returns the particle's rank inside its cell with the count it adds (no second pass over the table);
//   2. one scan turns the per-cell counts into offsets;
//   3. every particle's record goes straight to its place in cell order (the only copy shared memory ever holds), in a layout
//      padded by one element per sixteen so that the blocked read of step 4 has no bank conflicts;
//   4. every thread merges four consecutive records of that order, the warp merges runs, run heads issue the REDs.
// (Measured and dropped: merging the four particles of a thread BEFORE the grouping -- after a few pushes a tile of 1024 particles
// still holds 430-640 such entries in 36-60 cells, profiles/r2_deposit_kernel_history.txt -- and replacing the warp's shuffle
// reduction by one thread per (run of equal cells, node) over staged entry sums: 2.55 against 2.42 ms.)
#define DG_THREADS 256
#define DG_TILE (4 * DG_THREADS)
#define DG_IDX(q) ((q) + ((q) >> 4))
#define DG_REC (DG_TILE + DG_TILE / 16)          // padded length of one record array
static_assert(DG_TILE <= DT_SLOTS, "one hash slot per particle of the tile");

struct DepGroup {
    double rec[4][DG_REC];                            // cell fractions and weight of every record, in cell order
    uint32_t hkey[DT_SLOTS], hcnt[DT_SLOTS];          // hash table: lower node of the cell; particle counts, then offsets
    uint32_t skey[DG_TILE];                           // lower node of the cell of every record, in cell order
    alignas(16) uint32_t wsum[DG_THREADS / 32];
    uint32_t n_sorted;
};
static_assert(DG_THREADS == 256, "the scan of k_deposit_group adds eight warp sums");

__device__ __forceinline__ void ld4(const double *p, double v[4])
{
    asm("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}

// four consecutive particles of the stream per thread: positions -> cell keys and fractions (in place), weights to deposit
template <int VAL>
__device__ __forceinline__ void load_four(const MeshC &m, const double *__restrict__ x, const double *__restrict__ y,
                                          const double *__restrict__ z, const double *__restrict__ mpw,
                                          const double *__restrict__ vcomp, long long g, long long n,
                                          double X[4], double Y[4], double Z[4], double W[4], uint32_t key[4])
{
#pragma unroll
    for (int j = 0; j < 4; j++) { X[j] = Y[j] = Z[j] = W[j] = 0; key[j] = DT_EMPTY; }
    if (g >= n) return;
    ld4(x + g, X); ld4(y + g, Y); ld4(z + g, Z); ld4(mpw + g, W);          // capacity is a multiple of 1024: g+3 is allocated
    if (VAL > 0) {
        double V[4];
        ld4(vcomp + g, V);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            W[j] = W[j] * V[j];                         // mpw*v
            if (VAL == 2) W[j] = W[j] * V[j];           // (mpw*v)*v
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (g + j >= n) W[j] = 0;
        if (W[j] != 0) {
            int ci, cj, ck; double d0, d1, d2;
            cell3(m, X[j], Y[j], Z[j], ci, cj, ck, d0, d1, d2);
            if (ci >= 0 && cj >= 0 && ck >= 0) { key[j] = (uint32_t)node_u(m, ci, cj, ck); X[j] = d0; Y[j] = d1; Z[j] = d2; }
        }
    }
}

template <int MODE, int VAL = 0, int STRIDE = 1>
__global__ void __launch_bounds__(DG_THREADS, 4) k_deposit_group(MeshC m, const double *__restrict__ x, const double *__restrict__ y,
                                                                  const double *__restrict__ z, const double *__restrict__ mpw,
                                                                  long long n, double *acc, double scale, int ahead,
                                                                  const double *__restrict__ vcomp = nullptr)
{
    typedef typename AccVal<MODE>::T T;
    __shared__ DepGroup t;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31;
    const long long base = blockIdx.x * (long long)DG_TILE;
    if (tid < 4) {
        const long long pf = base + (long long)ahead * DG_TILE;
        if (ahead > 0 && pf + DG_TILE <= n) l2_prefetch((tid == 0 ? x : tid == 1 ? y : tid == 2 ? z : mpw) + pf, DG_TILE * 8);
    }
    reinterpret_cast<uint4 *>(t.hkey)[tid] = make_uint4(DT_EMPTY, DT_EMPTY, DT_EMPTY, DT_EMPTY);
    reinterpret_cast<uint4 *>(t.hcnt)[tid] = make_uint4(0, 0, 0, 0);
    // ---- 1. load, cells; every particle enters the hash table (lanes of one cell through one leader) and learns its rank
    double X[4], Y[4], Z[4], W[4];
    uint32_t key[4], sr[4];
    load_four<VAL>(m, x, y, z, mpw, vcomp, base + 4 * tid, n, X, Y, Z, W, key);
    __syncthreads();
    // (the four rounds in three sweeps -- all matches, all table updates, all broadcasts -- so that their latencies overlap: a round's
    // match, its leader's CAS + atomic and the shuffle behind them form one dependent chain)
    unsigned peers[4];
#pragma unroll
    for (int j = 0; j < 4; j++) peers[j] = __match_any_sync(FULL, key[j]);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        sr[j] = 0;
        if (key[j] != DT_EMPTY && lane == __ffs(peers[j]) - 1) {
            uint32_t h = (key[j] * 2654435761u) >> (32 - DT_SLOT_BITS);
            for (int probe = 0; probe < DT_SLOTS; probe++) {          // the table has a slot per particle of the tile: it cannot fill up
                const uint32_t old = atomicCAS(&t.hkey[h], DT_EMPTY, key[j]);
                if (old == DT_EMPTY || old == key[j]) break;
                h = (h + 1) & (DT_SLOTS - 1);
            }
            sr[j] = h | (atomicAdd(&t.hcnt[h], (uint32_t)__popc(peers[j])) << 16);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t v = __shfl_sync(FULL, sr[j], __ffs(peers[j]) - 1);
        sr[j] = v + ((uint32_t)__popc(peers[j] & ((1u << lane) - 1u)) << 16);      // slot | rank of this particle in its cell
    }
    __syncthreads();
    // ---- 2. counts -> offsets
    {
        const uint4 c4 = reinterpret_cast<const uint4 *>(t.hcnt)[tid];
        const uint32_t tsum = c4.x + c4.y + c4.z + c4.w;
        uint32_t inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += up;
        }
        if (lane == 31) t.wsum[tid >> 5] = inc;
        __syncthreads();
        uint32_t run = inc - tsum;
        {       // sums of the warps below this one: two broadcast loads and predicated adds instead of a chain of up to seven loads
            const uint4 wa = reinterpret_cast<const uint4 *>(t.wsum)[0], wb = reinterpret_cast<const uint4 *>(t.wsum)[1];
            const int w = tid >> 5;
            run += (w > 0 ? wa.x : 0u) + (w > 1 ? wa.y : 0u) + (w > 2 ? wa.z : 0u) + (w > 3 ? wa.w : 0u) +
                   (w > 4 ? wb.x : 0u) + (w > 5 ? wb.y : 0u) + (w > 6 ? wb.z : 0u);
        }
        reinterpret_cast<uint4 *>(t.hcnt)[tid] = make_uint4(run, run + c4.x, run + c4.x + c4.y, run + c4.x + c4.y + c4.z);
        if (tid == DG_THREADS - 1) t.n_sorted = run + tsum;
    }
    __syncthreads();
    // ---- 3. records to their place in cell order
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (key[j] == DT_EMPTY) continue;
        const uint32_t pos = t.hcnt[sr[j] & 0xffffu] + (sr[j] >> 16);
        const uint32_t at = DG_IDX(pos);
        t.rec[0][at] = X[j]; t.rec[1][at] = Y[j]; t.rec[2][at] = Z[j]; t.rec[3][at] = W[j];
        t.skey[pos] = key[j];
    }
    __syncthreads();
    // ---- 4. four consecutive records per thread, merged in registers; runs across lanes merged by the warp
    const int M = (int)t.n_sorted;
    const uint4 k4 = reinterpret_cast<const uint4 *>(t.skey)[tid];
    const uint32_t kq[4] = {k4.x, k4.y, k4.z, k4.w};
    T a[8];
#pragma unroll
    for (int q = 0; q < 8; q++) a[q] = 0;
    long long ua = -1;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int q = tid * 4 + j;
        if (q >= M) break;
        const long long u = (long long)kq[j];
        const int at = DG_IDX(q);
        T v[8];
        weights_from_fractions<MODE>(t.rec[0][at], t.rec[1][at], t.rec[2][at], t.rec[3][at], scale, v);
        if (u != ua) {
            if (ua >= 0) red8<MODE, STRIDE>(m, acc, ua, a);       // a cell boundary inside this thread's four records
            ua = u;
#pragma unroll
            for (int c = 0; c < 8; c++) a[c] = v[c];
        } else {
#pragma unroll
            for (int c = 0; c < 8; c++) a[c] += v[c];
        }
    }
    warp_deposit<MODE, STRIDE>(m, acc, ua, a);
}

// The stream is in cell order (directly after a sort): no grouping, no shared memory.  Four consecutive particles per thread are
// merged in registers, the warp merges runs across lanes (warp_deposit), run heads issue the REDs.
template <int MODE, int VAL = 0, int STRIDE = 1>
__global__ void __launch_bounds__(DG_THREADS) k_deposit_runs(MeshC m, const double *__restrict__ x, const double *__restrict__ y,
                                                              const double *__restrict__ z, const double *__restrict__ mpw,
                                                              long long n, double *acc, double scale, int ahead,
                                                              const double *__restrict__ vcomp = nullptr)
{
    typedef typename AccVal<MODE>::T T;
    const int tid = threadIdx.x;
    const long long base = blockIdx.x * (long long)DG_TILE;
    if (tid < 4) {
        const long long pf = base + (long long)ahead * DG_TILE;
        if (ahead > 0 && pf + DG_TILE <= n) l2_prefetch((tid == 0 ? x : tid == 1 ? y : tid == 2 ? z : mpw) + pf, DG_TILE * 8);
    }
    double X[4], Y[4], Z[4], W[4];
    uint32_t key[4];
    load_four<VAL>(m, x, y, z, mpw, vcomp, base + 4 * tid, n, X, Y, Z, W, key);
    T a[8];
#pragma unroll
    for (int q = 0; q < 8; q++) a[q] = 0;
    long long ua = -1;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (key[j] == DT_EMPTY) continue;
        T v[8];
        weights_from_fractions<MODE>(X[j], Y[j], Z[j], W[j], scale, v);
        if ((long long)key[j] != ua) {
            if (ua >= 0) red8<MODE, STRIDE>(m, acc, ua, a);           // rare: a cell boundary inside this thread's four particles
            ua = (long long)key[j];
#pragma unroll
            for (int q = 0; q < 8; q++) a[q] = v[q];
        } else {
#pragma unroll
            for (int q = 0; q < 8; q++) a[q] += v[q];
        }
    }
    warp_deposit<MODE, STRIDE>(m, acc, ua, a);
}

// den = acc / node_vol (0 where node_vol == 0): Field::operator/= (Field.h:125-134)
template <int MODE>
__global__ void k_den_finalize(long long nn, const double *__restrict__ acc, const double *__restrict__ node_vol,
                               double *__restrict__ den, double inv_scale)
{
    long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (u >= nn) return;
    double a;
    if (MODE == ESPIC_DEPOSIT_FP64) a = acc[u];
    else a = (double)reinterpret_cast<const long long *>(acc)[u] * inv_scale;
    double v = node_vol[u];
    den[u] = (v != 0) ? a / v : 0.0;
}

// =====================================================================================================
// push (Species::advance): two particles per thread, 128-bit streaming loads/stores of the seven SoA arrays,
// E gathered with one 256-bit load per node from the padded copy of ef, kill bits collected per warp, the density
// scatter of the survivors fused in (the new position is still in registers).
// =====================================================================================================

// spread the low 16 bits of v to the even bit positions
__device__ __forceinline__ uint32_t spread16(uint32_t v)
{
    v &= 0xffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

template <int WALL>
__device__ __forceinline__ bool push_one(const MeshC &m, const double *__restrict__ ef4, double s, double dt,
                                         double &x, double &y, double &z, double &vx, double &vy, double &vz, double &mpw)
{
    int i, j, k; double di, dj, dk;
    cell3(m, x, y, z, i, j, k, di, dj, dk);
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    if (k < 0) k = 0;
    double e[3];
    gather_ef(m, ef4, i, j, k, di, dj, dk, e);
    // part.vel += ef_part*(dt*charge/mass);  part.pos += part.vel*dt;   (Species.cpp:22-25)
    vx = vx + e[0] * s; vy = vy + e[1] * s; vz = vz + e[2] * s;
    x = x + vx * dt; y = y + vy * dt; z = z + vz * dt;
    if (WALL == ESPIC_WALL_ABSORB) {
        // Species.cpp:28-32, removal test of Species.cpp:39
        if (in_sphere(m, x, y, z) || !in_bounds(m, x, y, z)) mpw = 0;
        return !(mpw > 0);
    } else {
        // ch2/Species.cpp:32-36
        if (x < m.x0[0]) { x = 2 * m.x0[0] - x; vx *= -1.0; } else if (x >= m.xm[0]) { x = 2 * m.xm[0] - x; vx *= -1.0; }
        if (y < m.x0[1]) { y = 2 * m.x0[1] - y; vy *= -1.0; } else if (y >= m.xm[1]) { y = 2 * m.xm[1] - y; vy *= -1.0; }
        if (z < m.x0[2]) { z = 2 * m.x0[2] - z; vz *= -1.0; } else if (z >= m.xm[2]) { z = 2 * m.xm[2] - z; vz *= -1.0; }
        return false;
    }
}

// MIG (spatial decomposition, espic_migrate.cuh): a survivor whose new cell plane k lies outside [klo, khi) leaves this part;
// its bit goes to leave_words AND to dead_words (the removal after the exchange closes both kinds of holes in one pass, as
// ch9/MPI does: moveKernel clears `alive` for both, ch9/MPI/src/Species.cpp:66,185-188).
// DIAG (ESPIC_PUSH_DIAG): the survivors' contributions to Species::getRealCount / getMomentum / getKE (Species.cpp:84-108) are
// summed while velocity and weight are in registers -- one partial per block and quantity, folded by k_diag_fold -- so that
// the diagnostics every Main.cpp prints each step cost no second pass over the particles (k_diag: 1 ms at 2e8).
template <int WALL, bool FUSE, int MODE, bool MIG = false, bool DIAG = false>
__global__ void __launch_bounds__(256, FUSE ? 2 : 4) k_push(MeshC m, const double *__restrict__ ef4,
                                              double *__restrict__ px, double *__restrict__ py, double *__restrict__ pz,
                                              double *__restrict__ pvx, double *__restrict__ pvy, double *__restrict__ pvz,
                                              double *__restrict__ pmpw, long long n, double s, double dt,
                                              uint32_t *__restrict__ dead_words, double *acc, double scale, int ahead,
                                              uint32_t *__restrict__ leave_words, int klo, int khi, double *__restrict__ diag_part = nullptr)
{
    const int lane = threadIdx.x & 31;
    const long long i0 = 2 * (blockIdx.x * 256ll + threadIdx.x);
    const long long wbase = i0 - 2 * lane;                   // first particle of this warp (multiple of 64)
    // L2 prefetch of the tile a block `ahead` blocks later will stream (one 4 KB bulk prefetch per array, issued by seven
    // lanes of warp 0): by the time that block is scheduled its loads hit L2 instead of paying the full HBM latency
    if (ahead > 0 && threadIdx.x < 7) {
        const long long pf = (blockIdx.x + (long long)ahead) * 512;
        if (pf + 512 <= n) {
            const double *src = threadIdx.x == 0 ? px : threadIdx.x == 1 ? py : threadIdx.x == 2 ? pz : threadIdx.x == 3 ? pvx
                              : threadIdx.x == 4 ? pvy : threadIdx.x == 5 ? pvz : pmpw;
            l2_prefetch(src + pf, 4096);
        }
    }
    const bool warp_live = wbase < n;
    if (!warp_live) return;
    const bool v0 = i0 < n, v1 = i0 + 1 < n;
    double2 X = make_double2(m.x0[0], m.x0[0]), Y = make_double2(m.x0[1], m.x0[1]), Z = make_double2(m.x0[2], m.x0[2]);
    double2 VX = make_double2(0, 0), VY = VX, VZ = VX, W = VX;
    if (v0) {      // capacity is a multiple of 1024: slot i0+1 is always allocated
        X = ld2(px + i0); Y = ld2(py + i0); Z = ld2(pz + i0);
        VX = ld2(pvx + i0); VY = ld2(pvy + i0); VZ = ld2(pvz + i0); W = ld2(pmpw + i0);
    }
    if (!v1) { X.y = m.x0[0]; Y.y = m.x0[1]; Z.y = m.x0[2]; VX.y = 0; VY.y = 0; VZ.y = 0; W.y = 0; }   // unowned slot: keep the arithmetic tame
    const double w0_in = W.x, w1_in = W.y;
    bool dead0 = push_one<WALL>(m, ef4, s, dt, X.x, Y.x, Z.x, VX.x, VY.x, VZ.x, W.x);
    bool dead1 = push_one<WALL>(m, ef4, s, dt, X.y, Y.y, Z.y, VX.y, VY.y, VZ.y, W.y);
    dead0 = dead0 && v0;
    dead1 = dead1 && v1;
    if (v1) {
        st2(px + i0, X.x, X.y); st2(py + i0, Y.x, Y.y); st2(pz + i0, Z.x, Z.y);
        st2(pvx + i0, VX.x, VX.y); st2(pvy + i0, VY.x, VY.y); st2(pvz + i0, VZ.x, VZ.y);
    } else if (v0) {
        px[i0] = X.x; py[i0] = Y.x; pz[i0] = Z.x; pvx[i0] = VX.x; pvy[i0] = VY.x; pvz[i0] = VZ.x;
    }
    if (WALL == ESPIC_WALL_ABSORB) {
        if (dead0 && w0_in != 0) pmpw[i0] = 0;               // part.mpw = 0 (Species.cpp:31)
        if (dead1 && w1_in != 0) pmpw[i0 + 1] = 0;
        bool gone0 = dead0, gone1 = dead1;
        if (MIG) {
            int k; double dk;
            cell_frac(Z.x, m.x0[2], m.dh[2], m.rdh[2], m.nk, k, dk);
            const bool lv0 = v0 && !dead0 && (k < klo || k >= khi);
            cell_frac(Z.y, m.x0[2], m.dh[2], m.rdh[2], m.nk, k, dk);
            const bool lv1 = v1 && !dead1 && (k < klo || k >= khi);
            const uint32_t l0 = __ballot_sync(0xffffffffu, lv0), l1 = __ballot_sync(0xffffffffu, lv1);
            if (lane == 0 && warp_live) {
                leave_words[wbase >> 5] = spread16(l0) | (spread16(l1) << 1);
                if (wbase + 32 < n) leave_words[(wbase >> 5) + 1] = spread16(l0 >> 16) | (spread16(l1 >> 16) << 1);
            }
            gone0 = dead0 || lv0; gone1 = dead1 || lv1;
        }
        // kill bit of particle wbase + b is bit (b & 31) of word (wbase + b) >> 5: interleave the two ballots
        const uint32_t b0 = __ballot_sync(0xffffffffu, gone0), b1 = __ballot_sync(0xffffffffu, gone1);
        if (lane == 0 && warp_live) {
            const uint32_t lo = spread16(b0) | (spread16(b1) << 1);
            const uint32_t hi = spread16(b0 >> 16) | (spread16(b1 >> 16) << 1);
            dead_words[wbase >> 5] = lo;
            if (wbase + 32 < n) dead_words[(wbase >> 5) + 1] = hi;
        }
    }
    if (FUSE) deposit_pair<MODE>(m, acc, scale, v0 && !dead0, X.x, Y.x, Z.x, W.x, v1 && !dead1, X.y, Y.y, Z.y, W.y);
    if (DIAG) {
        double b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (v0 && !dead0) {          // the expressions of k_diag
            b[0] += W.x; b[1] += VX.x * W.x; b[2] += VY.x * W.x; b[3] += VZ.x * W.x;
            b[4] += W.x * (VX.x * VX.x + VY.x * VY.x + VZ.x * VZ.x);
        }
        if (v1 && !dead1) {
            b[0] += W.y; b[1] += VX.y * W.y; b[2] += VY.y * W.y; b[3] += VZ.y * W.y;
            b[4] += W.y * (VX.y * VX.y + VY.y * VY.y + VZ.y * VZ.y);
        }
        // Five warp sums in nine shuffles instead of twenty-five: every exchange halves the number of quantities a lane carries
        // (lanes 16-31 take over b[4..7], then bit 3 and bit 2 of the lane split what is left), two plain butterfly steps finish.
        // Lane 4q ends up with the warp's sum of quantity q.  One partial per WARP and quantity, no block barrier: the warps of
        // this memory-bound kernel retire independently (five block_sum calls cost it as much as the separate k_diag pass).
        const unsigned FULL = 0xffffffffu;
        double c4[4], d2[2], e1;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const double r = __shfl_xor_sync(FULL, (lane & 16) ? b[i] : b[i + 4], 16);
            c4[i] = ((lane & 16) ? b[i + 4] : b[i]) + r;
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const double r = __shfl_xor_sync(FULL, (lane & 8) ? c4[i] : c4[i + 2], 8);
            d2[i] = ((lane & 8) ? c4[i + 2] : c4[i]) + r;
        }
        {
            const double r = __shfl_xor_sync(FULL, (lane & 4) ? d2[0] : d2[1], 4);
            e1 = ((lane & 4) ? d2[1] : d2[0]) + r;
        }
        e1 += __shfl_xor_sync(FULL, e1, 2);
        e1 += __shfl_xor_sync(FULL, e1, 1);
        if ((lane & 3) == 0) diag_part[(wbase >> 6) * 8 + (lane >> 2)] = e1;      // one 64-byte record per warp (slots 5-7 are zero)
    }
}

// first fold of the per-warp partials of a DIAG push: nparts records of 8 doubles (nq <= 8 used) -> nq x gridDim (fixed order:
// reproducible).  Every thread walks whole records, so a warp reads 2 KB of consecutive memory per step.
__global__ void __launch_bounds__(256) k_diag_fold(const double *__restrict__ part, long long nparts, int nq, double *__restrict__ out)
{
    __shared__ double sh[32];
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < nparts; i += (long long)gridDim.x * 256) {
        const double4 lo = *reinterpret_cast<const double4 *>(part + i * 8);
        const double4 hi = *reinterpret_cast<const double4 *>(part + i * 8 + 4);
        a[0] += lo.x; a[1] += lo.y; a[2] += lo.z; a[3] += lo.w; a[4] += hi.x; a[5] += hi.y; a[6] += hi.z; a[7] += hi.w;
    }
    for (int q = 0; q < nq; q++) {
        const double t = block_sum(a[q], sh);
        if (threadIdx.x == 0) out[(size_t)q * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) k_reduce_final(const double *__restrict__ part, int nparts, int nq, double *__restrict__ out);

// Push with the tile-grouping scatter fused in (ESPIC_PUSH_FUSE_DEPOSIT): a block owns DT_TILE = 1024 consecutive particles,
// pushes them two per thread in two rounds exactly as k_push does (same loads, same arithmetic, same kill words), and hands
// every survivor's NEW position to the shared-memory grouping of k_deposit_tile while it is still in registers.  Against
// k_push + k_deposit_tile this saves the second read of x, y, z, mpw (32 of 136 bytes per particle) and a launch, and the
// hashing overlaps the push's memory time (the push alone runs at 84 % of the HBM peak with most issue slots idle).
template <int WALL, int MODE>
__global__ void __launch_bounds__(DT_THREADS, 4) k_push_tile(MeshC m, const double *__restrict__ ef4,
                                                             double *__restrict__ px, double *__restrict__ py, double *__restrict__ pz,
                                                             double *__restrict__ pvx, double *__restrict__ pvy, double *__restrict__ pvz,
                                                             double *__restrict__ pmpw, long long n, double s, double dt,
                                                             uint32_t *__restrict__ dead_words, double *acc, double scale, int ahead)
{
    __shared__ DepTile t;
    const int tid = threadIdx.x, lane = tid & 31;
    const long long base = blockIdx.x * (long long)DT_TILE;
    if (ahead > 0 && tid < 7) {          // L2 prefetch of the tile a block `ahead` blocks later will stream (see k_push)
        const long long pf = base + (long long)ahead * DT_TILE;
        if (pf + DT_TILE <= n) {
            const double *src = tid == 0 ? px : tid == 1 ? py : tid == 2 ? pz : tid == 3 ? pvx : tid == 4 ? pvy : tid == 5 ? pvz : pmpw;
            l2_prefetch(src + pf, DT_TILE * 8);
        }
    }
    dt_clear(t);
    __syncthreads();
#pragma unroll 1
    for (int round = 0; round < DT_PER / 2; round++) {
        const int p = 2 * (round * DT_THREADS + tid);
        const long long i0 = base + p;
        const long long wbase = i0 - 2 * lane;                   // first particle of this warp in this round (multiple of 64)
        const bool v0 = i0 < n, v1 = i0 + 1 < n;
        double2 X = make_double2(m.x0[0], m.x0[0]), Y = make_double2(m.x0[1], m.x0[1]), Z = make_double2(m.x0[2], m.x0[2]);
        double2 VX = make_double2(0, 0), VY = VX, VZ = VX, W = VX;
        if (v0) {      // capacity is a multiple of 1024: slot i0+1 is always allocated
            X = ld2(px + i0); Y = ld2(py + i0); Z = ld2(pz + i0);
            VX = ld2(pvx + i0); VY = ld2(pvy + i0); VZ = ld2(pvz + i0); W = ld2(pmpw + i0);
        }
        if (!v1) { X.y = m.x0[0]; Y.y = m.x0[1]; Z.y = m.x0[2]; VX.y = 0; VY.y = 0; VZ.y = 0; W.y = 0; }
        if (!v0) W.x = 0;
        const double w0_in = W.x, w1_in = W.y;
        bool dead0 = push_one<WALL>(m, ef4, s, dt, X.x, Y.x, Z.x, VX.x, VY.x, VZ.x, W.x);
        bool dead1 = push_one<WALL>(m, ef4, s, dt, X.y, Y.y, Z.y, VX.y, VY.y, VZ.y, W.y);
        dead0 = dead0 && v0;
        dead1 = dead1 && v1;
        if (v1) {
            st2(px + i0, X.x, X.y); st2(py + i0, Y.x, Y.y); st2(pz + i0, Z.x, Z.y);
            st2(pvx + i0, VX.x, VX.y); st2(pvy + i0, VY.x, VY.y); st2(pvz + i0, VZ.x, VZ.y);
        } else if (v0) {
            px[i0] = X.x; py[i0] = Y.x; pz[i0] = Z.x; pvx[i0] = VX.x; pvy[i0] = VY.x; pvz[i0] = VZ.x;
        }
        if (WALL == ESPIC_WALL_ABSORB) {
            if (dead0 && w0_in != 0) pmpw[i0] = 0;               // part.mpw = 0 (Species.cpp:31)
            if (dead1 && w1_in != 0) pmpw[i0 + 1] = 0;
            if (wbase < n) {                                     // warp-uniform: the kill words of this warp's 64 particles
                const uint32_t b0 = __ballot_sync(0xffffffffu, dead0), b1 = __ballot_sync(0xffffffffu, dead1);
                if (lane == 0) {
                    dead_words[wbase >> 5] = spread16(b0) | (spread16(b1) << 1);
                    if (wbase + 32 < n) dead_words[(wbase >> 5) + 1] = spread16(b0 >> 16) | (spread16(b1 >> 16) << 1);
                }
            }
        }
        // survivors go into the tile's grouping with their new position (the reference scatters after the removal, Species.cpp:51-62)
        dt_insert<true>(m, t, p, X.x, Y.x, Z.x, (v0 && !dead0) ? W.x : 0.0);
        dt_insert<true>(m, t, p + 1, X.y, Y.y, Z.y, (v1 && !dead1) ? W.y : 0.0);
    }
    __syncthreads();
    dt_finish<MODE, 1, true>(m, t, acc, scale);
}

// ---- removal in the reference's swap-with-last order (Species.cpp:36-46) ------------------------------
// With D dead among n, L = n-D survivors.  The sequential loop fills the holes (dead, idx < L) in ascending
// order with the live particles of the tail (idx >= L) taken from the end: hole rank r <- tail-live rank r.
// e(idx) = #dead before idx (exclusive scan of the per-warp dead words).

__global__ void __launch_bounds__(256) k_dead_popc(const uint32_t *__restrict__ words, long long nw, uint32_t *__restrict__ cnt)
{
    long long w = blockIdx.x * 256ll + threadIdx.x;
    if (w < nw) cnt[w] = __popc(words[w]);
}

__global__ void __launch_bounds__(256) k_fill_lists(const uint32_t *__restrict__ words, long long nw, long long n,
                                                    const uint32_t *__restrict__ pre, const uint32_t *__restrict__ coff,
                                                    const unsigned long long *__restrict__ d_total,
                                                    long long *__restrict__ holes, long long *__restrict__ fillers)
{
    long long w = blockIdx.x * 256ll + threadIdx.x;
    if (w >= nw) return;
    const long long D = (long long)*d_total;
    const long long L = n - D;
    const uint32_t bitsw = words[w];
    const long long first = w * 32;
    if (first + 32 <= L && bitsw == 0) return;          // fully live head word: nothing to record
    const long long e0 = (long long)scan_at(pre, coff, w);
    for (int b = 0; b < 32; b++) {
        long long idx = first + b;
        if (idx >= n) break;
        bool dead = (bitsw >> b) & 1u;
        long long e = e0 + __popc(bitsw & ((1u << b) - 1u));
        if (dead) { if (idx < L) holes[e] = idx; }
        else if (idx >= L) fillers[(n - 1 - idx) - D + e] = idx;
    }
}

__global__ void __launch_bounds__(256) k_move(long long dmax, const long long *__restrict__ holes, const long long *__restrict__ fillers,
                                              double *p0, double *p1, double *p2, double *p3, double *p4, double *p5, double *p6)
{
    long long r = blockIdx.x * 256ll + threadIdx.x;
    if (r >= dmax) return;
    long long h = holes[r];
    if (h < 0) return;
    long long f = fillers[r];
    p0[h] = p0[f]; p1[h] = p1[f]; p2[h] = p2[f]; p3[h] = p3[f]; p4[h] = p4[f]; p5[h] = p5[f]; p6[h] = p6[f];
}


// Fixed point: weights are accumulated as round(w * 2^shift) in int64.  2^shift is chosen so that the sum of every
// weight in the whole system (all ranks) stays below 2^62; all ranks agree on it through a max-allreduce.
static int prepare_acc(espic_ctx *c, Species &s, int mode)
{
    CK(cudaMemsetAsync(s.acc, 0, (size_t)c->m.nn * sizeof(double), c->stream));
    s.acc_mode = mode;
    s.acc_shift = 0;
    if (mode == ESPIC_DEPOSIT_FIXED) {
        double bound = (double)std::max<long long>(s.np, 1) * (s.mpw_max > 0 ? s.mpw_max : 1.0);
        int r = espic_comm_max_double(c, &bound);
        if (r) return r;
        bound *= c->nranks;
        int e;
        frexp(bound, &e);         // bound < 2^e
        s.acc_shift = std::min(62 - e, 62);
    }
    return 0;
}

// count the dead of the species' kill words (one bit per particle of [0,n)), then remove them in the reference's order
static int compact_dead(espic_ctx *c, Species &s, long long n)
{
    s.diag_valid = false;
    const long long nw = (n + 31) / 32;
    int r;
    static const bool trace = getenv("ESPIC_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_a = trace ? now() : 0;
    if ((r = ensure_buf(&c->cell_cnt, &c->cell_cap, nw, c->stream))) return r;
    k_dead_popc<<<nblk(nw, 256), 256, 0, c->stream>>>(s.kill_words, nw, c->cell_cnt);
    LAUNCH_CHECK(c);
    if ((r = espic_scan_u32(c, c->cell_cnt, nw, c->dscal))) return r;
    unsigned long long *h = (unsigned long long *)c->hpin;
    CK(cudaMemcpyAsync(h, c->dscal, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    const double t_b = trace ? now() : 0;
    CK(cudaStreamSynchronize(c->stream));
    const double t_c = trace ? now() : 0;
    const long long D = (long long)h[0];
    if (D == 0) return 0;
    if (D < n) {
        if ((r = ensure_buf(&c->lists, &c->lists_cap, std::max(2 * D, n / 64), c->stream))) return r;
        CK(cudaMemsetAsync(c->lists, 0xff, (size_t)D * sizeof(long long), c->stream));
        k_fill_lists<<<nblk(nw, 256), 256, 0, c->stream>>>(s.kill_words, nw, n, c->scan_pre, c->scan_coff, c->dscal,
                                                           c->lists, c->lists + D);
        LAUNCH_CHECK(c);
        k_move<<<nblk(D, 256), 256, 0, c->stream>>>(D, c->lists, c->lists + D, s.p[0], s.p[1], s.p[2], s.p[3], s.p[4], s.p[5], s.p[6]);
        LAUNCH_CHECK(c);
    }
    s.np = n - D;
    if (trace) fprintf(stderr, "[espic_push] n=%lld D=%lld host ms: enqueue %.3f  wait-for-count %.3f  removal-enqueue %.3f\n",
                       n, D, t_b - t_a, t_c - t_b, now() - t_c);
    return 0;
}

extern "C" int espic_push(espic_ctx *c, int sp, double dt, int wall_mode, int flags)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    Species &s = c->sp[sp];
    const long long n = s.np;
    const bool fuse = (flags & ESPIC_PUSH_FUSE_DEPOSIT) != 0;
    const int mode = (flags & ESPIC_PUSH_FIXED_POINT) ? ESPIC_DEPOSIT_FIXED : ESPIC_DEPOSIT_FP64;
    MIG_GUARD(c, s, "espic_push");
    const bool mig = (flags & ESPIC_PUSH_MIGRATE) != 0;
    if (mig && (!c->mig || wall_mode != ESPIC_WALL_ABSORB || fuse || (flags & ESPIC_PUSH_NO_COMPACT))) {
        espic_set_error("espic_push: ESPIC_PUSH_MIGRATE needs espic_domain_set, ESPIC_WALL_ABSORB and neither FUSE_DEPOSIT nor NO_COMPACT");
        return -1;
    }
    s.acc_fresh = false;
    s.diag_valid = false;
    // diagnostics ride along only on the plain absorbing push (the removal's synchronisation then also delivers them)
    const bool diag = (flags & ESPIC_PUSH_DIAG) && !fuse && !mig && wall_mode == ESPIC_WALL_ABSORB && !(flags & ESPIC_PUSH_NO_COMPACT);
    if (fuse) { int r = prepare_acc(c, s, mode); if (r) return r; }
    if (n == 0) { if (fuse) s.acc_fresh = true; if (mig) { s.mig_stage = 1; s.mig_n = 0; } return 0; }
    const double sfac = dt * s.charge / s.mass;     // Species.cpp:22, evaluated as the reference does
    const double scale = ldexp(1.0, s.acc_shift);
    const long long nw = (n + 31) / 32;
    int r;
    // with MIGRATE the words must also cover the arrivals appended before the removal: leave headroom so they rarely regrow
    if (wall_mode == ESPIC_WALL_ABSORB) { if ((r = ensure_buf(&s.kill_words, &s.kill_cap, mig ? nw + nw / 16 + 1024 : nw, c->stream))) return r; }
    if (mig) { if ((r = ensure_buf(&s.leave_words, &s.leave_cap, nw, c->stream))) return r; }
    const unsigned grid = nblk((n + 1) / 2, 256);
    const long long nwarp = (n + 63) / 64;           // one partial record (8 doubles, 5 used) per warp of the push
    const int fold_blocks = 4 * c->sm_count;
    if (diag && (r = ensure_buf(&c->red, &c->red_cap, 8 * nwarp + 5 * fold_blocks, c->stream))) return r;
    if (!c->push_ev0) { CK(cudaEventCreate(&c->push_ev0)); CK(cudaEventCreate(&c->push_ev1)); }
    CK(cudaEventRecord(c->push_ev0, c->stream));
    static const int ahead_env = getenv("ESPIC_PUSH_PREFETCH") ? atoi(getenv("ESPIC_PUSH_PREFETCH")) : -1;
    // default distance: two blocks per SM (measured plateau: 1-3 blocks per SM)
    const int ahead = ahead_env >= 0 ? ahead_env : 2 * c->sm_count;
#define PUSH_ARGS c->m, c->ef4, s.p[0], s.p[1], s.p[2], s.p[3], s.p[4], s.p[5], s.p[6], n, sfac, dt, s.kill_words, s.acc, scale, ahead, s.leave_words, c->dom_klo, c->dom_khi
    // the fused scatter groups by cell inside a 1024-particle tile (k_push_tile); ESPIC_FUSE_WARP_MERGE=1 selects the round-1
    // variant that only merges runs of equal cells inside a warp (2x slower once the cell order has decayed)
    static const bool warp_fuse = getenv("ESPIC_FUSE_WARP_MERGE") != nullptr;
    const unsigned tgrid = nblk(n, DT_TILE);
#define TILE_ARGS c->m, c->ef4, s.p[0], s.p[1], s.p[2], s.p[3], s.p[4], s.p[5], s.p[6], n, sfac, dt, s.kill_words, s.acc, scale, ahead / 2
    if (wall_mode == ESPIC_WALL_ABSORB) {
        if (mig) k_push<ESPIC_WALL_ABSORB, false, ESPIC_DEPOSIT_FP64, true><<<grid, 256, 0, c->stream>>>(PUSH_ARGS);
        else if (diag) k_push<ESPIC_WALL_ABSORB, false, ESPIC_DEPOSIT_FP64, false, true><<<grid, 256, 0, c->stream>>>(PUSH_ARGS, c->red);
        else if (!fuse) k_push<ESPIC_WALL_ABSORB, false, ESPIC_DEPOSIT_FP64><<<grid, 256, 0, c->stream>>>(PUSH_ARGS);
        else if (!warp_fuse && mode == ESPIC_DEPOSIT_FP64) k_push_tile<ESPIC_WALL_ABSORB, ESPIC_DEPOSIT_FP64><<<tgrid, DT_THREADS, 0, c->stream>>>(TILE_ARGS);
        else if (!warp_fuse) k_push_tile<ESPIC_WALL_ABSORB, ESPIC_DEPOSIT_FIXED><<<tgrid, DT_THREADS, 0, c->stream>>>(TILE_ARGS);
        else if (mode == ESPIC_DEPOSIT_FP64) k_push<ESPIC_WALL_ABSORB, true, ESPIC_DEPOSIT_FP64><<<grid, 256, 0, c->stream>>>(PUSH_ARGS);
        else k_push<ESPIC_WALL_ABSORB, true, ESPIC_DEPOSIT_FIXED><<<grid, 256, 0, c->stream>>>(PUSH_ARGS);
    } else if (wall_mode == ESPIC_WALL_REFLECT) {
        if (!fuse) k_push<ESPIC_WALL_REFLECT, false, ESPIC_DEPOSIT_FP64><<<grid, 256, 0, c->stream>>>(PUSH_ARGS);
        else if (!warp_fuse && mode == ESPIC_DEPOSIT_FP64) k_push_tile<ESPIC_WALL_REFLECT, ESPIC_DEPOSIT_FP64><<<tgrid, DT_THREADS, 0, c->stream>>>(TILE_ARGS);
        else if (!warp_fuse) k_push_tile<ESPIC_WALL_REFLECT, ESPIC_DEPOSIT_FIXED><<<tgrid, DT_THREADS, 0, c->stream>>>(TILE_ARGS);
        else if (mode == ESPIC_DEPOSIT_FP64) k_push<ESPIC_WALL_REFLECT, true, ESPIC_DEPOSIT_FP64><<<grid, 256, 0, c->stream>>>(PUSH_ARGS);
        else k_push<ESPIC_WALL_REFLECT, true, ESPIC_DEPOSIT_FIXED><<<grid, 256, 0, c->stream>>>(PUSH_ARGS);
    } else { espic_set_error("espic_push: bad wall mode %d", wall_mode); return -1; }
#undef PUSH_ARGS
#undef TILE_ARGS
    LAUNCH_CHECK(c);
    CK(cudaEventRecord(c->push_ev1, c->stream));
    c->push_timed = true;
    if (fuse) s.acc_fresh = true;
    if (s.pushes_since_sort < (1 << 20)) s.pushes_since_sort++;
    if (wall_mode != ESPIC_WALL_ABSORB || (flags & ESPIC_PUSH_NO_COMPACT)) { s.n_settled = s.np; return 0; }
    if (mig) { s.mig_stage = 1; s.mig_n = n; return 0; }     // removal happens in espic_migrate, after the exchange

    double *hd = reinterpret_cast<double *>(c->hpin) + 100;
    if (diag) {
        double *fold = c->red + 8 * nwarp, *dout = reinterpret_cast<double *>(c->dscal + 100);
        k_diag_fold<<<fold_blocks, 256, 0, c->stream>>>(c->red, nwarp, 5, fold);
        LAUNCH_CHECK(c);
        k_reduce_final<<<1, 256, 0, c->stream>>>(fold, fold_blocks, 5, dout);
        LAUNCH_CHECK(c);
        CK(cudaMemcpyAsync(hd, dout, 5 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));     // arrives with the removal's count
    }
    int rr = compact_dead(c, s, n);
    s.n_settled = s.np;
    if (diag && rr == 0) {
        for (int q = 0; q < 5; q++) s.diag_sums[q] = hd[q];
        s.diag_valid = true;
    }
    return rr;
}

// device time of the most recent k_push launch alone (no removal bookkeeping), from events on the launching stream
extern "C" int espic_last_push_ms(espic_ctx *c, double *ms)
{
    *ms = 0;
    if (!c->push_timed) return 0;
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->push_ev1));
    float f = 0;
    CK(cudaEventElapsedTime(&f, c->push_ev0, c->push_ev1));
    *ms = f;
    return 0;
}

extern "C" int espic_deposit(espic_ctx *c, int sp, int mode)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    Species &s = c->sp[sp];
    if (mode != ESPIC_DEPOSIT_FP64 && mode != ESPIC_DEPOSIT_FIXED) { espic_set_error("espic_deposit: bad mode %d", mode); return -1; }
    MIG_GUARD(c, s, "espic_deposit");
    if (!(s.acc_fresh && s.acc_mode == mode)) {
        int r = prepare_acc(c, s, mode);
        if (r) return r;
        if (s.np > 0) {
            const double scale = ldexp(1.0, s.acc_shift);
            // Directly after a sort (no push in between) the ungrouped tile kernel (merge 4 consecutive particles per thread, then
            // the warp) is the cheapest.  A few pushes later the thermal x/y motion (2 % of the particles change cell per step)
            // has cut the runs to pieces -- in either sort order, measured: profiles/r2_sort_order_sweep.txt -- and the grouping
            // kernel wins (its time does not depend on the disorder).  ESPIC_DEPOSIT_PLAIN_STEPS moves the switch.
            static const int plain_steps = getenv("ESPIC_DEPOSIT_PLAIN_STEPS") ? atoi(getenv("ESPIC_DEPOSIT_PLAIN_STEPS")) : 0;
            const bool ordered = s.pushes_since_sort <= plain_steps;
            static const int ahead_env = getenv("ESPIC_DEPOSIT_PREFETCH") ? atoi(getenv("ESPIC_DEPOSIT_PREFETCH")) : -1;
            const int ahead = ahead_env >= 0 ? ahead_env : 2 * c->sm_count;
            const unsigned tgrid = nblk(s.np, DT_TILE);
            static const bool v1 = getenv("ESPIC_DEPOSIT_V1") != nullptr;      // the round-1 tile kernel, for A/B measurements
#define DEP_ARGS c->m, s.p[0], s.p[1], s.p[2], s.p[6], s.np, s.acc, scale, ahead
            if (mode == ESPIC_DEPOSIT_FP64) {
                if (v1 && ordered) k_deposit_tile<ESPIC_DEPOSIT_FP64, 0, 1, false><<<tgrid, DT_THREADS, 0, c->stream>>>(DEP_ARGS);
                else if (v1) k_deposit_tile<ESPIC_DEPOSIT_FP64><<<tgrid, DT_THREADS, 0, c->stream>>>(DEP_ARGS);
                else if (ordered) k_deposit_runs<ESPIC_DEPOSIT_FP64><<<tgrid, DG_THREADS, 0, c->stream>>>(DEP_ARGS);
                else k_deposit_group<ESPIC_DEPOSIT_FP64><<<tgrid, DG_THREADS, 0, c->stream>>>(DEP_ARGS);
            } else {
                if (v1 && ordered) k_deposit_tile<ESPIC_DEPOSIT_FIXED, 0, 1, false><<<tgrid, DT_THREADS, 0, c->stream>>>(DEP_ARGS);
                else if (v1) k_deposit_tile<ESPIC_DEPOSIT_FIXED><<<tgrid, DT_THREADS, 0, c->stream>>>(DEP_ARGS);
                else if (ordered) k_deposit_runs<ESPIC_DEPOSIT_FIXED><<<tgrid, DG_THREADS, 0, c->stream>>>(DEP_ARGS);
                else k_deposit_group<ESPIC_DEPOSIT_FIXED><<<tgrid, DG_THREADS, 0, c->stream>>>(DEP_ARGS);
            }
#undef DEP_ARGS
            LAUNCH_CHECK(c);
        }
    }
    static const bool trace_ar = getenv("ESPIC_TRACE") != nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr, t2 = nullptr;
    if (trace_ar && c->nranks > 1) {
        cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventCreate(&t2);
        cudaEventRecord(t0, c->stream);
        double dummy = 0;
        espic_comm_max_double(c, &dummy);              // a tiny collective first: absorbs the skew between the ranks
        cudaEventRecord(t1, c->stream);
    }
    int r = espic_comm_allreduce_acc(c, s);
    if (r) return r;
    if (t0) {
        cudaEventRecord(t2, c->stream);
        cudaEventSynchronize(t2);
        float skew = 0, ar = 0;
        cudaEventElapsedTime(&skew, t0, t1); cudaEventElapsedTime(&ar, t1, t2);
        fprintf(stderr, "[espic_deposit rank %d] wait for peers %.3f ms, all-reduce of %lld doubles %.3f ms\n", c->rank, skew, c->m.nn, ar);
        cudaEventDestroy(t0); cudaEventDestroy(t1); cudaEventDestroy(t2);
    }
    const double inv_scale = ldexp(1.0, -s.acc_shift);
    if (mode == ESPIC_DEPOSIT_FP64)
        k_den_finalize<ESPIC_DEPOSIT_FP64><<<nblk(c->m.nn, 256), 256, 0, c->stream>>>(c->m.nn, s.acc, c->node_vol, s.den, inv_scale);
    else
        k_den_finalize<ESPIC_DEPOSIT_FIXED><<<nblk(c->m.nn, 256), 256, 0, c->stream>>>(c->m.nn, s.acc, c->node_vol, s.den, inv_scale);
    LAUNCH_CHECK(c);
    s.acc_fresh = false;      // the accumulator now holds the cross-rank sum / has been consumed
    return 0;
}

// =====================================================================================================
// velocity moments (ch4 Species::sampleMoments / computeGasProperties / clearSamples, Species.cpp:190-241)
// =====================================================================================================

// One thread per particle; the seven sampled quantities times eight node weights are 56 sums per particle (same node and
// factor order as Field::scatter).  A warp whose 32 particles share one cell (the common case right after a cell sort)
// adds them up with shuffles and lane 0 issues the 56 REDs; any other warp scatters per particle.
__global__ void __launch_bounds__(256) k_sample_moments(MeshC m, const double *__restrict__ x, const double *__restrict__ y,
                                                        const double *__restrict__ z, const double *__restrict__ vx,
                                                        const double *__restrict__ vy, const double *__restrict__ vz,
                                                        const double *__restrict__ mpw, long long n, double *mom)
{
    const long long idx = blockIdx.x * 256ll + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (idx - lane >= n) return;                                     // whole warp past the end
    double w = 0, u = 0, v = 0, ww = 0, di = 0, dj = 0, dk = 0;
    long long u0 = -1;
    if (idx < n) {
        w = mpw[idx];
        if (w != 0) {
            int i, j, k;
            cell3(m, x[idx], y[idx], z[idx], i, j, k, di, dj, dk);
            if (i >= 0 && j >= 0 && k >= 0) { u0 = node_u(m, i, j, k); u = vx[idx]; v = vy[idx]; ww = vz[idx]; }
        }
    }
    // values in the reference's evaluation order: mpw*vel (component-wise), mpw*vx*vx = (mpw*vx)*vx, ...
    const double val[7] = {w, w * u, w * v, w * ww, w * u * u, w * v * v, w * ww * ww};
    const long long nn = m.nn, sj = m.ni, sk = (long long)m.ni * m.nj;
    const double ai = 1 - di, aj = 1 - dj, ak = 1 - dk;
    const double fi[8] = {ai, di, di, ai, ai, di, di, ai}, fj[8] = {aj, aj, dj, dj, aj, aj, dj, dj}, fk[8] = {ak, ak, ak, ak, dk, dk, dk, dk};
    const long long off[8] = {0, 1, 1 + sj, sj, sk, 1 + sk, 1 + sj + sk, sj + sk};
    const long long first = __shfl_sync(0xffffffffu, u0, 0);
    const bool uniform = __all_sync(0xffffffffu, u0 == first) && first >= 0;
    if (!uniform && u0 < 0) return;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const long long un = (uniform ? first : u0) + off[c];
        double *dst[7] = {mom + un, mom + nn + 3 * un, mom + nn + 3 * un + 1, mom + nn + 3 * un + 2,
                          mom + 4 * nn + un, mom + 5 * nn + un, mom + 6 * nn + un};
#pragma unroll
        for (int q = 0; q < 7; q++) {
            double t = val[q] * fi[c] * fj[c] * fk[c];
            if (uniform) {
                t = warp_sum(t);
                if (lane == 0) atomicAdd(dst[q], t);
            } else {
                atomicAdd(dst[q], t);
            }
        }
    }
}

// Species::computeGasProperties with Field operator/ (ch4/Field.h:192-204)
__global__ void __launch_bounds__(256) k_gas_properties(long long nn, double mass, double *mom)
{
    const long long u = blockIdx.x * 256ll + threadIdx.x;
    if (u >= nn) return;
    const double K = 1.380648e-23;          // Const::K (ch4/World.h:18)
    const double count = mom[u];
    double vel[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { vel[c] = (count != 0) ? mom[nn + 3 * u + c] / count : 0.0; mom[7 * nn + 3 * u + c] = vel[c]; }
    double T = 0;
    if (count > 0) {
        const double u2 = mom[4 * nn + u] / count, v2 = mom[5 * nn + u] / count, w2 = mom[6 * nn + u] / count;
        const double uu = u2 - vel[0] * vel[0], vv = v2 - vel[1] * vel[1], wv = w2 - vel[2] * vel[2];
        T = mass / (2 * K) * (uu + vv + wv);
    }
    mom[10 * nn + u] = T;
}

extern "C" int espic_sample_moments(espic_ctx *c, int sp)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    int r;
    if ((r = espic_ensure_moments(c, sp))) return r;
    Species &s = c->sp[sp];
    if (s.np == 0) return 0;
    static const bool simple = getenv("ESPIC_MOMENTS_SIMPLE") != nullptr;
    if (simple) {
        k_sample_moments<<<nblk(s.np, 256), 256, 0, c->stream>>>(c->m, s.p[0], s.p[1], s.p[2], s.p[3], s.p[4], s.p[5], s.p[6], s.np, s.mom);
        LAUNCH_CHECK(c);
        return 0;
    }
    // one pass of the tile-grouping scatter per sampled quantity (7 x 32..40 B/particle of traffic, but no per-particle REDs)
    const long long nn = c->m.nn;
    const unsigned grid = nblk(s.np, DT_TILE);
    const int ahead = 0;
#define TILE_ARGS(dst, v) c->m, s.p[0], s.p[1], s.p[2], s.p[6], s.np, (dst), 1.0, ahead, (v)
    k_deposit_group<ESPIC_DEPOSIT_FP64, 0, 1><<<grid, DG_THREADS, 0, c->stream>>>(TILE_ARGS(s.mom, nullptr));              // n_sum
    LAUNCH_CHECK(c);
    for (int q = 0; q < 3; q++) {
        k_deposit_group<ESPIC_DEPOSIT_FP64, 1, 3><<<grid, DG_THREADS, 0, c->stream>>>(TILE_ARGS(s.mom + nn + q, s.p[3 + q]));  // nv_sum
        LAUNCH_CHECK(c);
        k_deposit_group<ESPIC_DEPOSIT_FP64, 2, 1><<<grid, DG_THREADS, 0, c->stream>>>(TILE_ARGS(s.mom + (4 + q) * nn, s.p[3 + q]));   // nuu, nvv, nww
        LAUNCH_CHECK(c);
    }
#undef TILE_ARGS
    return 0;
}

extern "C" int espic_compute_gas_properties(espic_ctx *c, int sp)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    int r;
    if ((r = espic_ensure_moments(c, sp))) return r;
    k_gas_properties<<<nblk(c->m.nn, 256), 256, 0, c->stream>>>(c->m.nn, c->sp[sp].mass, c->sp[sp].mom);
    LAUNCH_CHECK(c);
    return 0;
}

extern "C" int espic_clear_samples(espic_ctx *c, int sp)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    int r;
    if ((r = espic_ensure_moments(c, sp))) return r;
    CK(cudaMemsetAsync(c->sp[sp].mom, 0, (size_t)c->m.nn * 7 * sizeof(double), c->stream));
    return 0;
}

// =====================================================================================================
// sort by cell (counting sort; cell key as ch4 World::XtoC: c = k*(nj-1)*(ni-1) + j*(ni-1) + i)
// =====================================================================================================

// Sort key: the cell, refined by SORT_ZBINS slabs of the cell along z.  The beam drifts along +z, so the particles of a
// cell cross into the next cell in the order of their z fraction; with the slabs contiguous in memory, the particles
// that have crossed after s steps are still a contiguous run (the warp-level run merging of the deposit and the L1
// locality of the gather survive between sorts), instead of being interleaved one by one with those that have not.
#define SORT_ZBINS 8
// order 0: the cell index of ch4 World::XtoC (k slowest); order 1 (ESPIC_SORT_DRIFT_Z): k FASTEST -- the particles of one
// (i,j) column of cells are contiguous and ordered along z.  A beam drifting along z then keeps its order from step to step
// (the whole column shifts together; only the thermal x/y motion, ~2 % of the particles per step at 300 m/s, breaks runs),
// where order 0 loses 22 % of the particles of every cell to a far-away key each step.
__device__ __forceinline__ long long cell_key(const MeshC &m, double x, double y, double z, int order)
{
    int i, j, k; double d0, d1, d2;
    cell3(m, x, y, z, i, j, k, d0, d1, d2);
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    if (k < 0) { k = 0; d2 = 0; }
    const int zbins = order >> 8;                       // (the callers pack the number of z slabs per cell above the order)
    int zb = (int)(d2 * zbins);
    zb = zb < 0 ? 0 : (zb > zbins - 1 ? zbins - 1 : zb);
    if ((order & 255) == 1) return (((long long)j * (m.ni - 1) + i) * (long long)(m.nk - 1) + k) * zbins + zb;
    return (((long long)k * (m.nj - 1) + j) * (long long)(m.ni - 1) + i) * zbins + zb;
}

// The sort runs in three passes (round 2; the one-pass scatter it replaces wrote every particle's seven values to its new
// place with 8-byte stores spread over many cache lines -- ncu: 17 sectors per store request, 5 GB of read-for-write fills,
// 35 % of the DRAM bandwidth, every warp waiting on its atomic -- and took 10.9 of the sort's 12.5 ms at 2e8 particles):
//   1. k_cell_count  keys of all particles (stored, 4 bytes each) and the population of every key;
//   2. k_cell_rank   every particle draws its place inside its key (one atomic per warp and key) and writes ITS INDEX there:
//                    the only scattered stores of the sort are 4 bytes wide;
//   3. k_cell_gather place j of the new order reads the particle src[j]: seven gathers whose neighbours in j are mostly
//                    neighbours in memory (the stream was in order a few steps ago), seven fully coalesced streaming stores.
__global__ void __launch_bounds__(256) k_cell_count(MeshC m, const double *__restrict__ x, const double *__restrict__ y,
                                                    const double *__restrict__ z, long long n, uint32_t *__restrict__ cnt,
                                                    uint32_t *__restrict__ keys, int order)
{
    long long idx = blockIdx.x * 256ll + threadIdx.x;
    if (idx >= n) return;
    const uint32_t cell = (uint32_t)cell_key(m, __ldcs(x + idx), __ldcs(y + idx), __ldcs(z + idx), order);
    keys[idx] = cell;
    // warp-aggregate equal keys (sorted input: most of a warp shares a cell)
    unsigned act = __activemask();
    unsigned peers = __match_any_sync(act, cell);
    int leader = __ffs(peers) - 1;
    if ((threadIdx.x & 31) == leader) atomicAdd(cnt + cell, (uint32_t)__popc(peers));
}

// four consecutive particles per thread: the four atomics of a thread are independent and stay in flight together (the pass is
// bound by their latency, not by its 8 bytes per particle)
__global__ void __launch_bounds__(256) k_cell_rank(long long n, uint32_t *__restrict__ cnt, const uint32_t *__restrict__ pre,
                                                   const uint32_t *__restrict__ coff, const uint32_t *__restrict__ keys,
                                                   uint32_t *__restrict__ src)
{
    const unsigned FULL = 0xffffffffu;
    const long long i0 = 4 * (blockIdx.x * 256ll + threadIdx.x);
    const int lane = threadIdx.x & 31;
    if (i0 - 4 * lane >= n) return;                               // whole warp past the end
    uint32_t k[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    if (i0 + 3 < n) {
        const uint4 k4 = *reinterpret_cast<const uint4 *>(keys + i0);
        k[0] = k4.x; k[1] = k4.y; k[2] = k4.z; k[3] = k4.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) if (i0 + j < n) k[j] = keys[i0 + j];
    }
    unsigned peers[4];
    uint32_t basec[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        peers[j] = __match_any_sync(FULL, k[j]);
        basec[j] = 0;
        if (k[j] != 0xffffffffu && lane == __ffs(peers[j]) - 1)
            basec[j] = atomicSub(cnt + k[j], (uint32_t)__popc(peers[j])) + scan_at(pre, coff, k[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t b = __shfl_sync(FULL, basec[j], __ffs(peers[j]) - 1);
        // places [b - popc, b) of the key; ranks in lane order
        if (k[j] != 0xffffffffu) src[b - __popc(peers[j]) + __popc(peers[j] & ((1u << lane) - 1u))] = (uint32_t)(i0 + j);
    }
}

__global__ void __launch_bounds__(256) k_cell_gather(long long n, const uint32_t *__restrict__ src,
                                                     const double *__restrict__ s0, const double *__restrict__ s1, const double *__restrict__ s2,
                                                     const double *__restrict__ s3, const double *__restrict__ s4, const double *__restrict__ s5,
                                                     const double *__restrict__ s6,
                                                     double *__restrict__ d0, double *__restrict__ d1, double *__restrict__ d2,
                                                     double *__restrict__ d3, double *__restrict__ d4, double *__restrict__ d5,
                                                     double *__restrict__ d6)
{
    long long j = blockIdx.x * 256ll + threadIdx.x;
    if (j >= n) return;
    const long long i = src[j];
    const double a0 = s0[i], a1 = s1[i], a2 = s2[i], a3 = s3[i], a4 = s4[i], a5 = s5[i], a6 = s6[i];
    __stcs(d0 + j, a0); __stcs(d1 + j, a1); __stcs(d2 + j, a2); __stcs(d3 + j, a3);
    __stcs(d4 + j, a4); __stcs(d5 + j, a5); __stcs(d6 + j, a6);
}

extern "C" int espic_sort_by_cell(espic_ctx *c, int sp) { return espic_sort_particles(c, sp, ESPIC_SORT_XTOC); }

extern "C" int espic_sort_particles(espic_ctx *c, int sp, int order)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    if (order != ESPIC_SORT_XTOC && order != ESPIC_SORT_DRIFT_Z) { espic_set_error("espic_sort_particles: bad order %d", order); return -1; }
    Species &s = c->sp[sp];
    MIG_GUARD(c, s, "espic_sort_by_cell");
    const long long n = s.np;
    if (n < 2) return 0;
    if (s.substep && s.n_settled < n) {
        espic_set_error("espic_sort_by_cell: species %d holds particles added since its last espic_push_surface; their "
                        "Particle::dt is tracked by index -- sort after the advance", sp);
        return -1;
    }
    static const int zbins_env = getenv("ESPIC_SORT_ZBINS") ? atoi(getenv("ESPIC_SORT_ZBINS")) : SORT_ZBINS;
    const int zbins = zbins_env < 1 ? 1 : (zbins_env > 64 ? 64 : zbins_env);
    const long long nc = (long long)(c->m.ni - 1) * (c->m.nj - 1) * (c->m.nk - 1) * zbins;
    const int korder = order | (zbins << 8);
    int r;
    if (s.alt_cap < s.cap) {
        for (int q = 0; q < 7; q++) {
            if (s.alt[q]) { CK(cudaStreamSynchronize(c->stream)); CK(cudaFree(s.alt[q])); s.alt[q] = nullptr; }
            CK(cudaMalloc(&s.alt[q], (size_t)s.cap * sizeof(double)));
        }
        s.alt_cap = s.cap;
    }
    if (nc >= (1ll << 32) || n >= (1ll << 32)) { espic_set_error("espic_sort_particles: more than 2^32 keys or particles"); return -1; }
    if ((r = ensure_buf(&c->cell_cnt, &c->cell_cap, nc, c->stream))) return r;
    if ((r = ensure_buf(&c->sort_key, &c->sort_key_cap, n, c->stream))) return r;
    if ((r = ensure_buf(&c->sort_src, &c->sort_src_cap, n, c->stream))) return r;
    CK(cudaMemsetAsync(c->cell_cnt, 0, (size_t)nc * sizeof(uint32_t), c->stream));
    k_cell_count<<<nblk(n, 256), 256, 0, c->stream>>>(c->m, s.p[0], s.p[1], s.p[2], n, c->cell_cnt, c->sort_key, korder);
    LAUNCH_CHECK(c);
    if ((r = espic_scan_u32(c, c->cell_cnt, nc, c->dscal + 1))) return r;
    k_cell_rank<<<nblk((n + 3) / 4, 256), 256, 0, c->stream>>>(n, c->cell_cnt, c->scan_pre, c->scan_coff, c->sort_key, c->sort_src);
    LAUNCH_CHECK(c);
    k_cell_gather<<<nblk(n, 256), 256, 0, c->stream>>>(n, c->sort_src, s.p[0], s.p[1], s.p[2], s.p[3], s.p[4], s.p[5], s.p[6],
                                                       s.alt[0], s.alt[1], s.alt[2], s.alt[3], s.alt[4], s.alt[5], s.alt[6]);
    LAUNCH_CHECK(c);
    for (int q = 0; q < 7; q++) std::swap(s.p[q], s.alt[q]);
    std::swap(s.cap, s.alt_cap);
    s.pushes_since_sort = 0;
    s.sort_order = order;
    return 0;
}

// =====================================================================================================
// addParticle (Species.cpp:65-81) and ColdBeamSource::sample (Source.cpp:4-27)
// =====================================================================================================

// Philox4x32-10 (Salmon et al. 2011); counter = (idx_lo, idx_hi, step, stream), key = (seed_lo, seed_hi).
__device__ __forceinline__ void philox_uniform2(uint64_t seed, uint32_t stream, uint32_t step, uint64_t idx, double &u0, double &u1)
{
    uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = step, c3 = stream;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    uint64_t a = ((uint64_t)c1 << 32) | c0, b = ((uint64_t)c3 << 32) | c2;
    u0 = (double)(a >> 11) * (1.0 / 9007199254740992.0);
    u1 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
}

struct AddSrc {
    int philox;                 // 0: staged host particles, 1: cold beam from Philox, 2: warm (Maxwellian) beam from Philox
    const double *in[7];
    uint64_t seed; uint32_t stream, step;
    double Lx, Ly, v_drift, mpw0;
    double v_th;                // sqrt(2*K*T/mass), warm beam only
};

__device__ __forceinline__ void add_candidate(const MeshC &m, const AddSrc &a, long long i, double q[7])
{
    if (a.philox == 1) {
        double u0, u1;
        philox_uniform2(a.seed, a.stream, a.step, (uint64_t)i, u0, u1);
        q[0] = m.x0[0] + u0 * a.Lx;       // Source.cpp:22
        q[1] = m.x0[1] + u1 * a.Ly;
        q[2] = m.x0[2];
        q[3] = 0; q[4] = 0; q[5] = a.v_drift;
        q[6] = a.mpw0;
    } else if (a.philox == 2) {
        // WarmBeamSource::sample (ch4/Source.cpp:48-55) with Species::sampleIsotropicVel / sampleVth (ch4/Species.cpp:149-173).
        // Seven Philox blocks per particle, idx = 8*i + j: (x, y), (theta, cosine), then the nine Birdsall uniforms.
        double w[14];
#pragma unroll
        for (int j = 0; j < 7; j++) philox_uniform2(a.seed, a.stream, a.step, 8 * (uint64_t)i + j, w[2 * j], w[2 * j + 1]);
        q[0] = m.x0[0] + w[0] * a.Lx;
        q[1] = m.x0[1] + w[1] * a.Ly;
        q[2] = m.x0[2];
        const double theta = 2 * 3.141592653 * w[2];          // Const::PI as the reference spells it
        const double r = -1.0 + 2 * w[3];
        const double sc = sqrt(1 - r * r);
        const double v1 = a.v_th * (w[4] + w[5] + w[6] - 1.5);
        const double v2 = a.v_th * (w[7] + w[8] + w[9] - 1.5);
        const double v3 = a.v_th * (w[10] + w[11] + w[12] - 1.5);
        const double mag = 3 / sqrt(2.0 + 2 + 2) * sqrt(v1 * v1 + v2 * v2 + v3 * v3);
        q[3] = r * mag;
        q[4] = cos(theta) * sc * mag;
        q[5] = sin(theta) * sc * mag + a.v_drift;
        q[6] = a.mpw0;
    } else {
#pragma unroll
        for (int t = 0; t < 7; t++) q[t] = a.in[t][i];
    }
}

// wmax (caller-supplied candidates only): largest weight among the candidates, as the bit pattern of a non-negative double
// (the fixed-point scatter scales by it; the host used to walk the pinned array for this while the GPU sat idle)
__global__ void __launch_bounds__(256) k_add_flags(MeshC m, AddSrc a, long long n, uint32_t *__restrict__ flags,
                                                   unsigned long long *__restrict__ wmax)
{
    long long i = blockIdx.x * 256ll + threadIdx.x;
    double q[7];
    q[6] = 0;
    if (i < n) {
        add_candidate(m, a, i, q);
        flags[i] = in_bounds(m, q[0], q[1], q[2]) ? 1u : 0u;
    }
    if (wmax) {
        unsigned long long b = q[6] > 0 ? (unsigned long long)__double_as_longlong(q[6]) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
            b = t > b ? t : b;
        }
        if ((threadIdx.x & 31) == 0 && b) atomicMax(wmax, b);
    }
}

__global__ void __launch_bounds__(256) k_add_write(MeshC m, AddSrc a, long long n, const double *__restrict__ ef4,
                                                   const uint32_t *__restrict__ flags, const uint32_t *__restrict__ pre,
                                                   const uint32_t *__restrict__ coff, long long base, double qm, double hdt,
                                                   double *p0, double *p1, double *p2, double *p3, double *p4, double *p5, double *p6)
{
    long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= n || !flags[i]) return;
    double q[7];
    add_candidate(m, a, i, q);
    int ci, cj, ck; double di, dj, dk;
    cell3(m, q[0], q[1], q[2], ci, cj, ck, di, dj, dk);
    double e[3];
    gather_ef(m, ef4, ci, cj, ck, di, dj, dk, e);
    // vel -= charge/mass*ef_part*(0.5*dt)   (Species.cpp:77)
    long long dst = base + (long long)scan_at(pre, coff, i);
    p0[dst] = q[0]; p1[dst] = q[1]; p2[dst] = q[2];
    p3[dst] = q[3] - e[0] * qm * hdt;
    p4[dst] = q[4] - e[1] * qm * hdt;
    p5[dst] = q[5] - e[2] * qm * hdt;
    p6[dst] = q[6];
}

static int add_common(espic_ctx *c, int sp, AddSrc &a, long long n, double dt, long long *n_added)
{
    Species &s = c->sp[sp];
    if (n_added) *n_added = 0;
    s.diag_valid = false;
    MIG_GUARD(c, s, "espic_species_add / espic_inject_*");
    if (n <= 0) return 0;
    int r;
    if ((r = espic_species_reserve(c, sp, s.np + n))) return r;
    if ((r = ensure_buf(&c->cell_cnt, &c->cell_cap, n, c->stream))) return r;
    unsigned long long *wmax = a.philox == 0 ? c->dscal + 110 : nullptr;
    if (wmax) CK(cudaMemsetAsync(wmax, 0, sizeof(unsigned long long), c->stream));
    k_add_flags<<<nblk(n, 256), 256, 0, c->stream>>>(c->m, a, n, c->cell_cnt, wmax);
    LAUNCH_CHECK(c);
    if ((r = espic_scan_u32(c, c->cell_cnt, n, c->dscal + 2))) return r;
    const double qm = s.charge / s.mass, hdt = 0.5 * dt;
    k_add_write<<<nblk(n, 256), 256, 0, c->stream>>>(c->m, a, n, c->ef4, c->cell_cnt, c->scan_pre, c->scan_coff, s.np, qm, hdt,
                                                     s.p[0], s.p[1], s.p[2], s.p[3], s.p[4], s.p[5], s.p[6]);
    LAUNCH_CHECK(c);
    unsigned long long *h = (unsigned long long *)c->hpin;
    CK(cudaMemcpyAsync(h + 2, c->dscal + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    if (wmax) CK(cudaMemcpyAsync(h + 110, wmax, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (wmax) {
        double w;
        memcpy(&w, h + 110, sizeof(double));
        if (w > s.mpw_max) s.mpw_max = w;
    }
    s.np += (long long)h[2];
    s.acc_fresh = false;
    if (n_added) *n_added = (long long)h[2];
    return 0;
}

// Start the host -> device copy of the candidates a later espic_species_add(ctx, sp, comp, n, ...) will admit, on the copy
// stream: called one step ahead (comp in pinned memory), the transfer overlaps the current step's kernels.  The matching add
// recognises the staged data by (species, comp[0], n); anything else simply copies as before.
extern "C" int espic_species_prefetch(espic_ctx *c, int sp, const double *const comp[7], long long n)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    if (n <= 0) return 0;
    if (7 * n > c->stage_cap) {
        CK(cudaStreamSynchronize(c->copy_stream));
        CK(cudaStreamSynchronize(c->stream));
        if (c->stage) CK(cudaFree(c->stage));
        CK(cudaMalloc(&c->stage, (size_t)7 * n * sizeof(double)));
        c->stage_cap = 7 * n;
    }
    // the staging buffer may still be read by the kernels of the previous add: order behind the compute stream
    CK(cudaEventRecord(c->ev_snap, c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->ev_snap, 0));
    for (int q = 0; q < 7; q++)
        CK(cudaMemcpyAsync(c->stage + q * n, comp[q], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
    CK(cudaEventRecord(c->ev_stage, c->copy_stream));
    c->stage_host = comp[0]; c->stage_n = n; c->stage_sp = sp;
    return 0;
}

extern "C" int espic_species_add(espic_ctx *c, int sp, const double *const comp[7], long long n, double dt, long long *n_added)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    if (n <= 0) { if (n_added) *n_added = 0; return 0; }
    int r;
    AddSrc a;
    memset(&a, 0, sizeof(a));
    a.philox = 0;
    if (c->stage_host == comp[0] && c->stage_n == n && c->stage_sp == sp) {
        // prefetched by espic_species_prefetch: wait for the copy engine, read the staging buffer in place
        CK(cudaStreamWaitEvent(c->stream, c->ev_stage, 0));
        for (int q = 0; q < 7; q++) a.in[q] = c->stage + q * n;
        c->stage_host = nullptr;
    } else {
        // stage the candidates in the reduction scratch
        if ((r = ensure_buf(&c->red, &c->red_cap, 7 * n, c->stream))) return r;
        for (int q = 0; q < 7; q++) {
            CK(cudaMemcpyAsync(c->red + q * n, comp[q], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            a.in[q] = c->red + q * n;
        }
    }
    return add_common(c, sp, a, n, dt, n_added);      // (the largest weight comes back with the admission count)
}

extern "C" int espic_inject_cold_beam(espic_ctx *c, int sp, double v_drift, double den, double dt,
                                      uint64_t seed, uint32_t stream, uint32_t step, long long *n_added)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    Species &s = c->sp[sp];
    const MeshC &m = c->m;
    // Source.cpp:6-18
    double Lx = m.dh[0] * (m.ni - 1);
    double Ly = m.dh[1] * (m.nj - 1);
    double A = Lx * Ly;
    double num_real = den * v_drift * A * dt;
    // the Bernoulli fraction comes from counter idx = 2^64-1 (host evaluation of the same Philox block)
    uint32_t c0 = 0xffffffffu, c1 = 0xffffffffu, c2 = step, c3 = stream, k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    uint64_t a64 = ((uint64_t)c1 << 32) | c0;
    double u = (double)(a64 >> 11) * (1.0 / 9007199254740992.0);
    long long num_sim = (int)(num_real / s.mpw0 + u);
    AddSrc a;
    memset(&a, 0, sizeof(a));
    a.philox = 1; a.seed = seed; a.stream = stream; a.step = step;
    a.Lx = Lx; a.Ly = Ly; a.v_drift = v_drift; a.mpw0 = s.mpw0;
    if (s.mpw0 > s.mpw_max) s.mpw_max = s.mpw0;
    return add_common(c, sp, a, num_sim, dt, n_added);
}

// WarmBeamSource::sample (ch4/Source.cpp:31-56): same draw count and admission as the cold beam, Maxwellian velocities
extern "C" int espic_inject_warm_beam(espic_ctx *c, int sp, double v_drift, double den, double T, double dt,
                                      uint64_t seed, uint32_t stream, uint32_t step, long long *n_added)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    Species &s = c->sp[sp];
    const MeshC &m = c->m;
    double Lx = m.dh[0] * (m.ni - 1);
    double Ly = m.dh[1] * (m.nj - 1);
    double A = Lx * Ly;
    double num_real = den * v_drift * A * dt;
    uint32_t c0 = 0xffffffffu, c1 = 0xffffffffu, c2 = step, c3 = stream, k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    uint64_t a64 = ((uint64_t)c1 << 32) | c0;
    double u = (double)(a64 >> 11) * (1.0 / 9007199254740992.0);
    long long num_sim = (int)(num_real / s.mpw0 + u);
    AddSrc a;
    memset(&a, 0, sizeof(a));
    a.philox = 2; a.seed = seed; a.stream = stream; a.step = step;
    a.Lx = Lx; a.Ly = Ly; a.v_drift = v_drift; a.mpw0 = s.mpw0;
    a.v_th = sqrt(2 * 1.380648e-23 * T / s.mass);          // Const::K (ch4/World.h:18), Species.cpp:152
    if (s.mpw0 > s.mpw_max) s.mpw_max = s.mpw0;
    return add_common(c, sp, a, num_sim, dt, n_added);
}

// =====================================================================================================
// diagnostics (Species.cpp:84-108)
// =====================================================================================================

__global__ void __launch_bounds__(256) k_diag(const double *__restrict__ vx, const double *__restrict__ vy, const double *__restrict__ vz,
                                              const double *__restrict__ mpw, long long n, double *__restrict__ part)
{
    __shared__ double sh[32];
    double a[5] = {0, 0, 0, 0, 0};
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        double w = mpw[i], x = vx[i], y = vy[i], z = vz[i];
        a[0] += w;
        a[1] += x * w; a[2] += y * w; a[3] += z * w;
        double v2 = x * x + y * y + z * z;
        a[4] += w * v2;
    }
    for (int q = 0; q < 5; q++) {
        double t = block_sum(a[q], sh);
        if (threadIdx.x == 0) part[(size_t)q * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) k_reduce_final(const double *__restrict__ part, int nparts, int nq, double *__restrict__ out)
{
    __shared__ double sh[32];
    for (int q = 0; q < nq; q++) {
        double a = 0;
        for (int i = threadIdx.x; i < nparts; i += 256) a += part[(size_t)q * nparts + i];
        double t = block_sum(a, sh);
        if (threadIdx.x == 0) out[q] = t;
    }
}

extern "C" int espic_species_diag(espic_ctx *c, int sp, double out[5])
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    Species &s = c->sp[sp];
    for (int q = 0; q < 5; q++) out[q] = 0;
    if (s.np == 0) return 0;
    if (s.diag_valid) {          // summed by the last espic_push(ESPIC_PUSH_DIAG); nothing touched the particles since
        out[0] = s.diag_sums[0];
        out[1] = s.diag_sums[1] * s.mass; out[2] = s.diag_sums[2] * s.mass; out[3] = s.diag_sums[3] * s.mass;
        out[4] = 0.5 * s.mass * s.diag_sums[4];
        return 0;
    }
    int nb = (int)std::min<long long>(nblk(s.np, 256), 4 * c->sm_count);
    int r;
    if ((r = ensure_buf(&c->red, &c->red_cap, 5ll * nb, c->stream))) return r;
    k_diag<<<nb, 256, 0, c->stream>>>(s.p[3], s.p[4], s.p[5], s.p[6], s.np, c->red);
    LAUNCH_CHECK(c);
    double *dout = reinterpret_cast<double *>(c->dscal + 8);
    k_reduce_final<<<1, 256, 0, c->stream>>>(c->red, nb, 5, dout);
    LAUNCH_CHECK(c);
    double *h = reinterpret_cast<double *>(c->hpin) + 8;
    CK(cudaMemcpyAsync(h, dout, 5 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    out[0] = h[0];
    out[1] = h[1] * s.mass; out[2] = h[2] * s.mass; out[3] = h[3] * s.mass;    // mass*mom (Species.cpp:97)
    out[4] = 0.5 * s.mass * h[4];                                             // Species.cpp:107
    return 0;
}

#include "espic_surface.cuh"
#include "espic_collide.cuh"
#include "espic_migrate.cuh"
