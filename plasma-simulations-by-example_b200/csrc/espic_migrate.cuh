// espic_migrate.cuh -- spatial domain decomposition with particle migration (replaces ch9/MPI
// World::initMPIDomain, ch9/MPI/include/World.h:73-128, and Species::transferParticles, ch9/MPI/src/Species.cpp:204-313).
// Included at the end of espic_particles.cu (shares its scan and removal code).
//
//   * the mesh is cut along k (the slowest index of Field::U, the drift direction of the beam and the direction the slab
//     solver cuts) into `parts` slabs of CELLS: part r owns the cells kb[r] <= k < kb[r+1].  Every part keeps GLOBAL
//     coordinates and the full-size node arrays, so XtoL, gather and scatter are the single-domain arithmetic bit for bit
//     (ch9/MPI shifts x0 per process instead, World.h:118-127); a part's scatter touches only its own node planes
//     kb[r]..kb[r+1], the shared plane gets both neighbours' contributions through the density all-reduce exactly as
//     ch9/MPI Field::updateBoundaries adds them.
//   * owner(particle) = the part of the cell its z lies in (the same cell_frac as gather/scatter -- ch9/MPI compares the
//     logical coordinate against the local node count, Species.cpp:169-181).  A particle can go to ANY part, not only to a
//     face neighbour, and nothing is discarded (the reference drops particles that land outside the neighbour, :285-288).
//   * deterministic order: leavers are listed in particle order (ballot + scan), split per destination by a stable
//     partition, packed into one SoA segment [7][count] per destination; the holes are closed in the swap-with-last order
//     of espic_push; arrivals are appended by ascending source part, each source in its own particle order.
//   * espic_migrate: counts travel as one all-gathered parts x parts matrix, then ONE NCCL group of send/recv pairs moves
//     every segment straight into the free slots behind the receiver's live particles (no unpack pass).
//   * espic_migrate_pack / espic_migrate_segment expose the two halves so that several parts can live on one GPU
//     (tests/test_migration.py drives R contexts on one device and hands the segments over with espic_species_upload_device).

#define ESPIC_MAX_PARTS 64

struct DomC {
    int parts, part;
    int kb[ESPIC_MAX_PARTS + 1];
};

struct MigState {
    DomC dom;
    long long *idx = nullptr;       long long idx_cap = 0;      // leavers in particle order
    uint8_t *dest = nullptr;        long long dest_cap = 0;     // their destination parts
    long long *pidx = nullptr;      long long pidx_cap = 0;     // per destination: leaver indices (parts x L)
    double *send = nullptr;         long long send_cap = 0;     // packed segments, destination after destination
    unsigned long long *dcnt = nullptr;                         // device: counts[parts] | matrix[parts][parts+1] (last column: mpw_max bits)
    unsigned long long *hcnt = nullptr;                         // pinned mirror
    long long counts[ESPIC_MAX_PARTS];                          // of the most recent pack
    long long offs[ESPIC_MAX_PARTS + 1];
};

static void mig_free(espic_ctx *c)
{
    MigState *g = (MigState *)c->mig;
    if (!g) return;
    cudaFree(g->idx); cudaFree(g->dest); cudaFree(g->pidx); cudaFree(g->send); cudaFree(g->dcnt);
    if (g->hcnt) cudaFreeHost(g->hcnt);
    delete g;
    c->mig = nullptr;
}

void espic_migrate_destroy(espic_ctx *c) { mig_free(c); }

__device__ __forceinline__ int dom_owner(const DomC &d, int k)
{
    int r = 0;
    while (r + 1 < d.parts && k >= d.kb[r + 1]) r++;
    return r;
}

__device__ __forceinline__ int mig_dest_of(const MeshC &m, const DomC &d, double z)
{
    int k; double dk;
    cell_frac(z, m.x0[2], m.dh[2], m.rdh[2], m.nk, k, dk);
    if (k < 0) k = 0;
    return dom_owner(d, k);
}

// one lane per particle: bit (i & 31) of word i >> 5 says "leaves this part" (the layout of k_push's kill words, so the
// removal code of espic_push closes the holes); cnt[w] = popcount for the scan
__global__ void __launch_bounds__(256) k_mig_flags(MeshC m, DomC d, const double *__restrict__ pz, long long n,
                                                   uint32_t *__restrict__ words, uint32_t *__restrict__ cnt)
{
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    const bool leave = (i < n) && (mig_dest_of(m, d, pz[i]) != d.part);
    const uint32_t b = __ballot_sync(0xffffffffu, leave);
    if ((threadIdx.x & 31) == 0 && (i >> 5) < ((n + 31) >> 5)) { words[i >> 5] = b; cnt[i >> 5] = __popc(b); }
}

__global__ void __launch_bounds__(256) k_mig_list(MeshC m, DomC d, const double *__restrict__ pz, long long n,
                                                  const uint32_t *__restrict__ words, const uint32_t *__restrict__ pre,
                                                  const uint32_t *__restrict__ coff, long long *__restrict__ idx,
                                                  uint8_t *__restrict__ dest)
{
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= n) return;
    const uint32_t bits = words[i >> 5];
    const int b = (int)(i & 31);
    if (!((bits >> b) & 1u)) return;
    const long long e = (long long)scan_at(pre, coff, i >> 5) + __popc(bits & ((1u << b) - 1u));
    idx[e] = i;
    dest[e] = (uint8_t)mig_dest_of(m, d, pz[i]);
}

// stable partition of the leaver list by destination: block d walks the list in order and keeps the entries bound for d
__global__ void __launch_bounds__(1024) k_mig_partition(long long L, const uint8_t *__restrict__ dest, const long long *__restrict__ idx,
                                                        long long *__restrict__ pidx, unsigned long long *__restrict__ counts)
{
    __shared__ uint32_t wsum[32];
    const int d = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long running = 0;
    for (long long base = 0; base < L; base += 1024) {
        const long long e = base + threadIdx.x;
        const bool f = e < L && dest[e] == (uint8_t)d;
        const uint32_t b = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wsum[w] = __popc(b);
        __syncthreads();
        uint32_t woff = 0, tot = 0;
        for (int q = 0; q < 32; q++) { const uint32_t v = wsum[q]; if (q < w) woff += v; tot += v; }
        if (f) pidx[(long long)d * L + running + woff + __popc(b & ((1u << lane) - 1u))] = idx[e];
        running += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[d] = (unsigned long long)running;
}

// segment layout: [7][cnt] (SoA), so that the receiver appends each component with one contiguous transfer
__global__ void __launch_bounds__(256) k_mig_pack(long long cnt, const long long *__restrict__ pidx,
                                                  const double *__restrict__ p0, const double *__restrict__ p1, const double *__restrict__ p2,
                                                  const double *__restrict__ p3, const double *__restrict__ p4, const double *__restrict__ p5,
                                                  const double *__restrict__ p6, double *__restrict__ seg)
{
    const long long r = blockIdx.x * 256ll + threadIdx.x;
    if (r >= cnt) return;
    const long long i = pidx[r];
    seg[r] = p0[i]; seg[cnt + r] = p1[i]; seg[2 * cnt + r] = p2[i]; seg[3 * cnt + r] = p3[i];
    seg[4 * cnt + r] = p4[i]; seg[5 * cnt + r] = p5[i]; seg[6 * cnt + r] = p6[i];
}

extern "C" int espic_domain_set(espic_ctx *c, int parts, int part, const int *k_bounds)
{
    CK(cudaSetDevice(c->device));
    if (parts < 1 || parts > ESPIC_MAX_PARTS || part < 0 || part >= parts) { espic_set_error("espic_domain_set: part %d of %d (max %d parts)", part, parts, ESPIC_MAX_PARTS); return -1; }
    if (k_bounds[0] != 0 || k_bounds[parts] != c->m.nk - 1) { espic_set_error("espic_domain_set: k_bounds must run from 0 to nk-1 = %d cells", c->m.nk - 1); return -1; }
    for (int r = 0; r < parts; r++)
        if (k_bounds[r + 1] <= k_bounds[r]) { espic_set_error("espic_domain_set: k_bounds must increase (part %d is empty)", r); return -1; }
    mig_free(c);
    MigState *g = new MigState();
    g->dom.parts = parts; g->dom.part = part;
    for (int r = 0; r <= parts; r++) g->dom.kb[r] = k_bounds[r];
    for (int r = 0; r < parts; r++) { g->counts[r] = 0; g->offs[r] = 0; }
    g->offs[parts] = 0;
    const size_t nb = (size_t)(parts + parts * (parts + 1)) * sizeof(unsigned long long);
    c->mig = g;
    CK(cudaMalloc(&g->dcnt, nb));
    CK(cudaMallocHost(&g->hcnt, nb));
    CK(cudaMemsetAsync(g->dcnt, 0, nb, c->stream));
    return 0;
}

extern "C" int espic_domain_get(espic_ctx *c, int *parts, int *part, int *k_bounds)
{
    MigState *g = (MigState *)c->mig;
    if (!g) { *parts = 1; *part = 0; if (k_bounds) { k_bounds[0] = 0; k_bounds[1] = c->m.nk - 1; } return 0; }
    *parts = g->dom.parts; *part = g->dom.part;
    if (k_bounds) for (int r = 0; r <= g->dom.parts; r++) k_bounds[r] = g->dom.kb[r];
    return 0;
}

extern "C" int espic_migrate_pack(espic_ctx *c, int sp, long long *counts)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    MigState *g = (MigState *)c->mig;
    if (!g) { espic_set_error("espic_migrate_pack: no domain (espic_domain_set)"); return -1; }
    Species &s = c->sp[sp];
    if (s.substep && s.n_settled != s.np) { espic_set_error("espic_migrate_pack: species %d holds particles added since its last surface advance", sp); return -1; }
    const int P = g->dom.parts;
    for (int r = 0; r < P; r++) { g->counts[r] = 0; g->offs[r] = 0; if (counts) counts[r] = 0; }
    g->offs[P] = 0;
    const long long n = s.np;
    if (n == 0 || P == 1) return 0;
    const long long nw = (n + 31) / 32;
    int r;
    if ((r = ensure_buf(&c->dead_words, &c->dead_words_cap, nw, c->stream))) return r;
    if ((r = ensure_buf(&c->cell_cnt, &c->cell_cap, nw, c->stream))) return r;
    k_mig_flags<<<nblk(nw * 32, 256), 256, 0, c->stream>>>(c->m, g->dom, s.p[2], n, c->dead_words, c->cell_cnt);
    LAUNCH_CHECK(c);
    if ((r = espic_scan_u32(c, c->cell_cnt, nw, c->dscal))) return r;
    unsigned long long *h = (unsigned long long *)c->hpin;
    CK(cudaMemcpyAsync(h, c->dscal, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const long long L = (long long)h[0];
    if (L == 0) return 0;
    if ((r = ensure_buf(&g->idx, &g->idx_cap, L, c->stream))) return r;
    if ((r = ensure_buf(&g->dest, &g->dest_cap, L, c->stream))) return r;
    if ((r = ensure_buf(&g->pidx, &g->pidx_cap, L * P, c->stream))) return r;
    if ((r = ensure_buf(&g->send, &g->send_cap, 7 * L, c->stream))) return r;
    k_mig_list<<<nblk(n, 256), 256, 0, c->stream>>>(c->m, g->dom, s.p[2], n, c->dead_words, c->scan_pre, c->scan_coff, g->idx, g->dest);
    LAUNCH_CHECK(c);
    k_mig_partition<<<P, 1024, 0, c->stream>>>(L, g->dest, g->idx, g->pidx, g->dcnt);
    LAUNCH_CHECK(c);
    CK(cudaMemcpyAsync(g->hcnt, g->dcnt, (size_t)P * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    long long off = 0;
    for (int d = 0; d < P; d++) {
        g->counts[d] = (long long)g->hcnt[d];
        g->offs[d] = off;
        off += g->counts[d];
        if (counts) counts[d] = g->counts[d];
    }
    g->offs[P] = off;
    if (off != L || g->counts[g->dom.part] != 0) { espic_set_error("espic_migrate_pack: internal count mismatch (%lld of %lld)", off, L); return -1; }
    for (int d = 0; d < P; d++) {
        if (g->counts[d] == 0) continue;
        k_mig_pack<<<nblk(g->counts[d], 256), 256, 0, c->stream>>>(g->counts[d], g->pidx + (long long)d * L, s.p[0], s.p[1], s.p[2], s.p[3],
                                                                  s.p[4], s.p[5], s.p[6], g->send + 7 * g->offs[d]);
        LAUNCH_CHECK(c);
    }
    // close the holes in the reference's swap-with-last order (the leave bits sit in dead_words)
    if ((r = compact_dead(c, s, n))) return r;
    s.n_settled = s.np;
    s.acc_fresh = false;
    return 0;
}

extern "C" int espic_migrate_segment(espic_ctx *c, int dest, void **dptr, long long *count)
{
    MigState *g = (MigState *)c->mig;
    if (!g || dest < 0 || dest >= g->dom.parts) { espic_set_error("espic_migrate_segment: bad destination %d", dest); return -1; }
    *count = g->counts[dest];
    *dptr = g->counts[dest] > 0 ? (void *)(g->send + 7 * g->offs[dest]) : nullptr;
    return 0;
}

extern "C" int espic_migrate(espic_ctx *c, int sp, long long *n_sent, long long *n_received)
{
    SP_CHECK(c, sp);
    if (n_sent) *n_sent = 0;
    if (n_received) *n_received = 0;
    MigState *g = (MigState *)c->mig;
    if (!g) { espic_set_error("espic_migrate: no domain (espic_domain_set)"); return -1; }
    const int P = g->dom.parts, me = g->dom.part;
    if (P == 1) return 0;
    if (c->nranks != P || c->rank != me || !c->nccl) { espic_set_error("espic_migrate: domain part %d/%d needs a communicator of the same shape (rank %d/%d)", me, P, c->rank, c->nranks); return -1; }
    int r;
    if ((r = espic_migrate_pack(c, sp, nullptr))) return r;
    Species &s = c->sp[sp];
    // counts matrix M[src][dst] plus one column with the sender's largest weight (the fixed-point scale needs a global bound and
    // weights travel with the particles): my row goes up, the all-gather returns every row
    const int W = P + 1;
    unsigned long long *hmat = g->hcnt + P, *dmat = g->dcnt + P;
    unsigned long long *hrow = hmat + (size_t)me * W;
    for (int d = 0; d < P; d++) hrow[d] = (unsigned long long)g->counts[d];
    memcpy(&hrow[P], &s.mpw_max, sizeof(double));
    CK(cudaMemcpyAsync(dmat + (size_t)me * W, hrow, (size_t)W * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
    if ((r = espic_comm_allgather_bytes(c, dmat, (size_t)W * sizeof(unsigned long long)))) return r;
    CK(cudaMemcpyAsync(hmat, dmat, (size_t)P * W * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    long long recv_total = 0, recv_off[ESPIC_MAX_PARTS];
    for (int src = 0; src < P; src++) {
        recv_off[src] = recv_total;
        recv_total += src == me ? 0 : (long long)hmat[(size_t)src * W + me];
        double w; memcpy(&w, &hmat[(size_t)src * W + P], sizeof(double));
        if (w > s.mpw_max) s.mpw_max = w;
    }
    if ((r = espic_species_reserve(c, sp, s.np + recv_total))) return r;
    const double *sendp[ESPIC_MAX_PARTS]; double *recvp[ESPIC_MAX_PARTS][7]; long long sendn[ESPIC_MAX_PARTS], recvn[ESPIC_MAX_PARTS];
    for (int peer = 0; peer < P; peer++) {
        sendn[peer] = g->counts[peer];
        sendp[peer] = g->send + 7 * g->offs[peer];
        recvn[peer] = peer == me ? 0 : (long long)hmat[(size_t)peer * W + me];
        for (int q = 0; q < 7; q++) recvp[peer][q] = s.p[q] + s.np + recv_off[peer];
    }
    if ((r = espic_comm_exchange_segments(c, P, me, sendp, sendn, recvp, recvn))) return r;
    s.np += recv_total;
    s.n_settled = s.np;
    s.acc_fresh = false;
    if (recv_total > 0) s.pushes_since_sort = std::max(s.pushes_since_sort, 1);
    if (n_sent) *n_sent = g->offs[P];
    if (n_received) *n_received = recv_total;
    return 0;
}
