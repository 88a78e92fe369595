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
//   * the reference's sequence (Species::move, Species.cpp:189-213): the push clears `alive` for sphere hits, box exits AND
//     part crossings; transferParticles appends the arrivals; ONE removal sweep then fills every hole from the end of the
//     array -- so arrivals are the first to fill holes.  Same here: espic_push(ESPIC_PUSH_MIGRATE) writes the leave bits in
//     the push kernel (the new z is in registers) and leaves everything in place, espic_migrate packs the leavers, receives
//     the arrivals behind the last slot and runs the swap-with-last removal of espic_push once over old + new particles.
//   * deterministic order: leavers are listed in particle order (popcount scan of the leave words), split per destination
//     by a stable partition (per 1024-entry chunk: __match_any ranks + per-warp counts; a scan over the chunks) that writes
//     the SoA segment [7][count] of each destination directly; arrivals are appended by ascending source part, each
//     source in its own particle order.  tests/migration_model.py states the resulting order; the GPU reproduces it bit for bit.
//   * espic_migrate: the per-destination counts never visit the host on their own -- they are copied device-to-device into
//     this rank's row of the parts x (parts+1) matrix that is all-gathered (last column: the sender's largest weight, which
//     the fixed-point scale needs globally); then ONE NCCL group of send/recv pairs moves every segment straight into the
//     free slots behind the receiver's particles (no unpack pass).  Host synchronisations per call: leaver total, matrix,
//     removal count.
//   * espic_migrate_pack / espic_migrate_segment / espic_migrate_finish expose the halves so that several parts can live
//     on one GPU (tests/test_migration.py drives R contexts on one device and hands the segments over with
//     espic_species_upload_device).

#define ESPIC_MAX_PARTS 64
#define MIG_CHUNK 1024

struct DomC {
    int parts, part;
    int kb[ESPIC_MAX_PARTS + 1];
};

struct MigState {
    DomC dom;
    long long *idx = nullptr;       long long idx_cap = 0;      // leavers in particle order
    uint8_t *dest = nullptr;        long long dest_cap = 0;     // their destination parts
    uint32_t *hist = nullptr;       long long hist_cap = 0;     // per chunk and destination: count, then exclusive prefix
    double *send = nullptr;         long long send_cap = 0;     // packed segments, destination after destination
    unsigned long long *dcnt = nullptr;     // device: counts[P] | offsets[P] | matrix[P][P+1] (last column: mpw_max bits)
    unsigned long long *hcnt = nullptr;     // pinned mirror
    long long counts[ESPIC_MAX_PARTS];      // of the most recent pack (host copy, once read)
    long long offs[ESPIC_MAX_PARTS + 1];
    long long L = 0;                        // leavers of the most recent pack
    bool counts_on_host = false;
    bool warmed = false;                    // the NCCL point-to-point connections to every peer exist
};

// NCCL opens a point-to-point connection (and further channels for large messages) the first time a pair of ranks uses it in a
// given direction: 280-400 ms each on this box (measured), and in a beam case the backward direction is first used many steps
// into the run.  One dummy exchange of full-size messages with every peer, at the first espic_migrate, pays for all of them.
#define MIG_WARM_DOUBLES (1ll << 19)
static int mig_warm_connections(espic_ctx *c, MigState *g)
{
    const int P = g->dom.parts, me = g->dom.part;
    double *tmp = nullptr;
    const long long per = 7 * MIG_WARM_DOUBLES;
    CK(cudaMalloc(&tmp, (size_t)(2 * P) * per * sizeof(double)));
    CK(cudaMemsetAsync(tmp, 0, (size_t)(2 * P) * per * sizeof(double), c->stream));
    const double *sendp[ESPIC_MAX_PARTS]; double *recvp[ESPIC_MAX_PARTS][7]; long long sendn[ESPIC_MAX_PARTS], recvn[ESPIC_MAX_PARTS];
    for (int peer = 0; peer < P; peer++) {
        sendn[peer] = recvn[peer] = peer == me ? 0 : MIG_WARM_DOUBLES;
        sendp[peer] = tmp + (long long)peer * per;
        for (int q = 0; q < 7; q++) recvp[peer][q] = tmp + (long long)(P + peer) * per + q * MIG_WARM_DOUBLES;
    }
    int r = espic_comm_exchange_segments(c, P, me, sendp, sendn, recvp, recvn);
    cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    g->warmed = true;
    return r;
}

static void mig_free(espic_ctx *c)
{
    MigState *g = (MigState *)c->mig;
    if (!g) return;
    cudaFree(g->idx); cudaFree(g->dest); cudaFree(g->hist); cudaFree(g->send); cudaFree(g->dcnt);
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

// leave bits for particles that did not come out of a MIGRATE push (initial placement, injected particles): one lane per
// particle, bit (i & 31) of word i >> 5, the layout of k_push's kill words; every particle is alive, so kill words = leave words
__global__ void __launch_bounds__(256) k_mig_flags(MeshC m, DomC d, const double *__restrict__ pz, long long n,
                                                   uint32_t *__restrict__ leave_words, uint32_t *__restrict__ dead_words)
{
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    const bool leave = (i < n) && (mig_dest_of(m, d, pz[i]) != d.part);
    const uint32_t b = __ballot_sync(0xffffffffu, leave);
    if ((threadIdx.x & 31) == 0 && (i >> 5) < ((n + 31) >> 5)) { leave_words[i >> 5] = b; dead_words[i >> 5] = b; }
}

// one thread per word: the leavers of the word, in particle order, go to list slots prefix(word) ...
__global__ void __launch_bounds__(256) k_mig_list(MeshC m, DomC d, const double *__restrict__ pz, long long nw,
                                                  const uint32_t *__restrict__ words, const uint32_t *__restrict__ pre,
                                                  const uint32_t *__restrict__ coff, long long *__restrict__ idx,
                                                  uint8_t *__restrict__ dest)
{
    const long long w = blockIdx.x * 256ll + threadIdx.x;
    if (w >= nw) return;
    uint32_t bits = words[w];
    if (!bits) return;
    long long e = (long long)scan_at(pre, coff, w);
    while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const long long i = w * 32 + b;
        idx[e] = i;
        dest[e] = (uint8_t)mig_dest_of(m, d, pz[i]);
        e++;
    }
}

// Stable partition of the leaver list by destination, chunk by chunk.  Inside a chunk the rank of an entry among the
// entries with the same destination = (entries of earlier warps) + (earlier lanes of its __match_any group).
//   PACK = false: hist[chunk][d] = number of entries of the chunk bound for d
//   PACK = true : hist holds the exclusive prefix over the chunks; the entry's particle is copied to slot
//                 prefix + rank of destination d's SoA segment, which starts 7 * offs[d] doubles into `send`
template <bool PACK>
__global__ void __launch_bounds__(MIG_CHUNK) k_mig_partition(long long L, int P, const uint8_t *__restrict__ dest,
                                                             const long long *__restrict__ idx, uint32_t *__restrict__ hist,
                                                             const unsigned long long *__restrict__ counts,
                                                             const unsigned long long *__restrict__ offs,
                                                             const double *__restrict__ p0, const double *__restrict__ p1,
                                                             const double *__restrict__ p2, const double *__restrict__ p3,
                                                             const double *__restrict__ p4, const double *__restrict__ p5,
                                                             const double *__restrict__ p6, double *__restrict__ send)
{
    __shared__ uint16_t wcnt[MIG_CHUNK / 32][ESPIC_MAX_PARTS];
    for (int q = threadIdx.x; q < (MIG_CHUNK / 32) * ESPIC_MAX_PARTS; q += MIG_CHUNK) (&wcnt[0][0])[q] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long e = blockIdx.x * (long long)MIG_CHUNK + threadIdx.x;
    const unsigned d = e < L ? dest[e] : 0xffu;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (d != 0xffu && rank == 0) wcnt[w][d] = (uint16_t)__popc(peers);
    __syncthreads();
    if (!PACK) {
        if ((int)threadIdx.x < P) {
            uint32_t t = 0;
            for (int q = 0; q < MIG_CHUNK / 32; q++) t += wcnt[q][threadIdx.x];
            hist[(long long)blockIdx.x * P + threadIdx.x] = t;
        }
    } else if (d != 0xffu) {
        uint32_t before = 0;
        for (int q = 0; q < w; q++) before += wcnt[q][d];
        const long long cnt = (long long)counts[d];
        const long long r = (long long)hist[(long long)blockIdx.x * P + d] + before + rank;
        double *seg = send + 7 * (long long)offs[d];
        const long long i = idx[e];
        seg[r] = p0[i]; seg[cnt + r] = p1[i]; seg[2 * cnt + r] = p2[i]; seg[3 * cnt + r] = p3[i];
        seg[4 * cnt + r] = p4[i]; seg[5 * cnt + r] = p5[i]; seg[6 * cnt + r] = p6[i];
    }
}

// per destination: exclusive prefix of the chunk counts (in place), the total, and the segment offsets
__global__ void __launch_bounds__(ESPIC_MAX_PARTS) k_mig_partition_scan(long long nchunks, int P, uint32_t *__restrict__ hist,
                                                                        unsigned long long *__restrict__ counts,
                                                                        unsigned long long *__restrict__ offs)
{
    __shared__ unsigned long long tot[ESPIC_MAX_PARTS];
    const int d = threadIdx.x;
    if (d < P) {
        unsigned long long run = 0;
        for (long long b = 0; b < nchunks; b++) {
            const uint32_t v = hist[b * P + d];
            hist[b * P + d] = (uint32_t)run;
            run += v;
        }
        counts[d] = run;
        tot[d] = run;
    }
    __syncthreads();
    if (d == 0) {
        unsigned long long o = 0;
        for (int q = 0; q < P; q++) { offs[q] = o; o += tot[q]; }
    }
}

extern "C" int espic_domain_set(espic_ctx *c, int parts, int part, const int *k_bounds)
{
    CK(cudaSetDevice(c->device));
    if (parts < 1 || parts > ESPIC_MAX_PARTS || part < 0 || part >= parts) { espic_set_error("espic_domain_set: part %d of %d (max %d parts)", part, parts, ESPIC_MAX_PARTS); return -1; }
    if (k_bounds[0] != 0 || k_bounds[parts] != c->m.nk - 1) { espic_set_error("espic_domain_set: k_bounds must run from 0 to nk-1 = %d cells", c->m.nk - 1); return -1; }
    for (int r = 0; r < parts; r++)
        if (k_bounds[r + 1] <= k_bounds[r]) { espic_set_error("espic_domain_set: k_bounds must increase (part %d is empty)", r); return -1; }
    for (int q = 0; q < c->nsp; q++)
        if (c->sp[q].mig_stage != 0) { espic_set_error("espic_domain_set: a migration of species %d is pending", q); return -1; }
    mig_free(c);
    MigState *g = new MigState();
    g->dom.parts = parts; g->dom.part = part;
    for (int r = 0; r <= parts; r++) g->dom.kb[r] = k_bounds[r];
    for (int r = 0; r < parts; r++) { g->counts[r] = 0; g->offs[r] = 0; }
    g->offs[parts] = 0;
    const size_t nb = (size_t)(2 * parts + parts * (parts + 1)) * sizeof(unsigned long long);
    c->mig = g;
    c->dom_klo = k_bounds[part];
    c->dom_khi = part == parts - 1 ? (1 << 30) : k_bounds[part + 1];       // the last part also owns the clamped top cell
    CK(cudaMalloc(&g->dcnt, nb));
    CK(cudaMallocHost(&g->hcnt, nb));
    CK(cudaMemsetAsync(g->dcnt, 0, nb, c->stream));
    memset(g->hcnt, 0, nb);
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

// flags (unless a MIGRATE push left them) -> leaver list -> per-destination SoA segments.  Nothing is removed.
// After the call: g->L on the host; counts[P] and offsets[P] on the device (g->dcnt), NOT yet on the host.
static int mig_pack_enqueue(espic_ctx *c, int sp)
{
    MigState *g = (MigState *)c->mig;
    Species &s = c->sp[sp];
    const int P = g->dom.parts;
    // the packed segments live in ONE staging buffer per context: a second species may only be packed once the first is finished
    for (int q = 0; q < c->nsp; q++)
        if (q != sp && c->sp[q].mig_stage == 2) { espic_set_error("espic_migrate_pack: species %d is packed and not finished (espic_migrate_finish) -- its segments would be overwritten", q); return -1; }
    g->L = 0;
    g->counts_on_host = false;
    for (int r = 0; r < P; r++) { g->counts[r] = 0; g->offs[r] = 0; }
    g->offs[P] = 0;
    CK(cudaMemsetAsync(g->dcnt, 0, (size_t)2 * P * sizeof(unsigned long long), c->stream));
    int r;
    if (s.mig_stage == 0) {            // no MIGRATE push before: every particle is alive, flag the foreign ones now
        s.mig_n = s.np;
        const long long nw = (s.mig_n + 31) / 32;
        if (nw > 0) {
            if ((r = ensure_buf(&s.kill_words, &s.kill_cap, nw + nw / 16 + 1024, c->stream))) return r;
            if ((r = ensure_buf(&s.leave_words, &s.leave_cap, nw, c->stream))) return r;
            k_mig_flags<<<nblk(nw * 32, 256), 256, 0, c->stream>>>(c->m, g->dom, s.p[2], s.mig_n, s.leave_words, s.kill_words);
            LAUNCH_CHECK(c);
        }
    }
    s.mig_stage = 2;
    const long long n = s.mig_n, nw = (n + 31) / 32;
    if (n == 0 || P == 1) return 0;
    if ((r = ensure_buf(&c->cell_cnt, &c->cell_cap, nw, c->stream))) return r;
    k_dead_popc<<<nblk(nw, 256), 256, 0, c->stream>>>(s.leave_words, nw, c->cell_cnt);
    LAUNCH_CHECK(c);
    if ((r = espic_scan_u32(c, c->cell_cnt, nw, c->dscal))) return r;
    unsigned long long *h = (unsigned long long *)c->hpin;
    CK(cudaMemcpyAsync(h, c->dscal, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const long long L = (long long)h[0];
    g->L = L;
    if (L == 0) return 0;
    const long long nchunks = (L + MIG_CHUNK - 1) / MIG_CHUNK;
    // the number of leavers drifts from step to step and a reallocation stalls the device (cudaFree synchronises: 180 ms measured
    // with 11 GB resident), so size for at least 1/128 of the particles (ensure_buf doubles that)
    const long long Lcap = std::max(L, n / 128 + 1024), ccap = (Lcap + MIG_CHUNK - 1) / MIG_CHUNK;
    if ((r = ensure_buf(&g->idx, &g->idx_cap, Lcap, c->stream))) return r;
    if ((r = ensure_buf(&g->dest, &g->dest_cap, Lcap, c->stream))) return r;
    if ((r = ensure_buf(&g->hist, &g->hist_cap, ccap * P, c->stream))) return r;
    if ((r = ensure_buf(&g->send, &g->send_cap, 7 * Lcap, c->stream))) return r;
    k_mig_list<<<nblk(nw, 256), 256, 0, c->stream>>>(c->m, g->dom, s.p[2], nw, s.leave_words, c->scan_pre, c->scan_coff, g->idx, g->dest);
    LAUNCH_CHECK(c);
#define MIG_PART_ARGS L, P, g->dest, g->idx, g->hist, g->dcnt, g->dcnt + P, s.p[0], s.p[1], s.p[2], s.p[3], s.p[4], s.p[5], s.p[6], g->send
    k_mig_partition<false><<<(unsigned)nchunks, MIG_CHUNK, 0, c->stream>>>(MIG_PART_ARGS);
    LAUNCH_CHECK(c);
    k_mig_partition_scan<<<1, ESPIC_MAX_PARTS, 0, c->stream>>>(nchunks, P, g->hist, g->dcnt, g->dcnt + P);
    LAUNCH_CHECK(c);
    k_mig_partition<true><<<(unsigned)nchunks, MIG_CHUNK, 0, c->stream>>>(MIG_PART_ARGS);
    LAUNCH_CHECK(c);
#undef MIG_PART_ARGS
    return 0;
}

static int mig_set_host_counts(espic_ctx *c, const unsigned long long *row)
{
    MigState *g = (MigState *)c->mig;
    const int P = g->dom.parts;
    long long off = 0;
    for (int d = 0; d < P; d++) { g->counts[d] = (long long)row[d]; g->offs[d] = off; off += g->counts[d]; }
    g->offs[P] = off;
    g->counts_on_host = true;
    if (off != g->L || g->counts[g->dom.part] != 0) { espic_set_error("espic_migrate: internal count mismatch (%lld of %lld leavers)", off, g->L); return -1; }
    return 0;
}

extern "C" int espic_migrate_pack(espic_ctx *c, int sp, long long *counts)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    MigState *g = (MigState *)c->mig;
    if (!g) { espic_set_error("espic_migrate_pack: no domain (espic_domain_set)"); return -1; }
    Species &s = c->sp[sp];
    if (s.mig_stage == 2) { espic_set_error("espic_migrate_pack: species %d is already packed (espic_migrate_finish)", sp); return -1; }
    if (s.substep && s.n_settled != s.np) { espic_set_error("espic_migrate_pack: species %d holds particles added since its last surface advance", sp); return -1; }
    int r;
    if ((r = mig_pack_enqueue(c, sp))) return r;
    const int P = g->dom.parts;
    if (g->L > 0) {
        CK(cudaMemcpyAsync(g->hcnt, g->dcnt, (size_t)P * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    } else {
        for (int d = 0; d < P; d++) g->hcnt[d] = 0;
    }
    if ((r = mig_set_host_counts(c, g->hcnt))) return r;
    if (counts) for (int d = 0; d < P; d++) counts[d] = g->counts[d];
    return 0;
}

extern "C" int espic_migrate_segment(espic_ctx *c, int dest, void **dptr, long long *count)
{
    MigState *g = (MigState *)c->mig;
    if (!g || dest < 0 || dest >= g->dom.parts || !g->counts_on_host) { espic_set_error("espic_migrate_segment: bad destination %d or nothing packed", dest); return -1; }
    *count = g->counts[dest];
    *dptr = g->counts[dest] > 0 ? (void *)(g->send + 7 * g->offs[dest]) : nullptr;
    return 0;
}

// the removal sweep of Species::move (ch9/MPI/src/Species.cpp:205-210) over the old particles and the arrivals behind them
extern "C" int espic_migrate_finish(espic_ctx *c, int sp)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    Species &s = c->sp[sp];
    if (s.mig_stage != 2) { espic_set_error("espic_migrate_finish: species %d has no packed migration", sp); return -1; }
    const long long nw_old = (s.mig_n + 31) / 32, nw_new = (s.np + 31) / 32;
    if (nw_new > s.kill_cap) {            // grow, keeping the kill bits
        uint32_t *nb = nullptr;
        const long long ncap = 2 * nw_new;
        CK(cudaMalloc(&nb, (size_t)ncap * sizeof(uint32_t)));
        if (nw_old > 0) CK(cudaMemcpyAsync(nb, s.kill_words, (size_t)nw_old * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaFree(s.kill_words));
        s.kill_words = nb;
        s.kill_cap = ncap;
    }
    if (nw_new > nw_old) CK(cudaMemsetAsync(s.kill_words + nw_old, 0, (size_t)(nw_new - nw_old) * sizeof(uint32_t), c->stream));
    s.mig_stage = 0;
    int r = 0;
    if (s.np > 0) r = compact_dead(c, s, s.np);
    s.n_settled = s.np;
    s.acc_fresh = false;
    return r;
}

extern "C" int espic_migrate(espic_ctx *c, int sp, long long *n_sent, long long *n_received)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    if (n_sent) *n_sent = 0;
    if (n_received) *n_received = 0;
    MigState *g = (MigState *)c->mig;
    if (!g) { espic_set_error("espic_migrate: no domain (espic_domain_set)"); return -1; }
    const int P = g->dom.parts, me = g->dom.part;
    Species &s = c->sp[sp];
    if (s.mig_stage == 2) { espic_set_error("espic_migrate: species %d is already packed (espic_migrate_finish)", sp); return -1; }
    if (P > 1 && (c->nranks != P || c->rank != me || !c->nccl)) { espic_set_error("espic_migrate: domain part %d/%d needs a communicator of the same shape (rank %d/%d)", me, P, c->rank, c->nranks); return -1; }
    int r;
    if (P > 1 && !g->warmed && (r = mig_warm_connections(c, g))) return r;
    static const bool trace = getenv("ESPIC_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = trace ? now() : 0;
    if ((r = mig_pack_enqueue(c, sp))) return r;
    const double t1 = trace ? now() : 0;
    long long recv_total = 0;
    double t2 = t1;
    if (P > 1) {
        // counts matrix M[src][dst] plus one column with the sender's largest weight: my row is filled on the device (the counts
        // never visit the host on their own), the all-gather returns every row
        const int W = P + 1;
        unsigned long long *hmat = g->hcnt + 2 * P, *dmat = g->dcnt + 2 * P;
        CK(cudaMemcpyAsync(dmat + (size_t)me * W, g->dcnt, (size_t)P * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
        unsigned long long *hw = g->hcnt + P;         // pinned staging word for the weight bound
        memcpy(hw, &s.mpw_max, sizeof(double));
        CK(cudaMemcpyAsync(dmat + (size_t)me * W + P, hw, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
        if ((r = espic_comm_allgather_bytes(c, dmat, (size_t)W * sizeof(unsigned long long)))) return r;
        CK(cudaMemcpyAsync(hmat, dmat, (size_t)P * W * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        t2 = trace ? now() : 0;
        if ((r = mig_set_host_counts(c, hmat + (size_t)me * W))) return r;
        long long recv_off[ESPIC_MAX_PARTS];
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
    } else {
        g->counts_on_host = true;
    }
    const double t3 = trace ? now() : 0;
    const long long sent = g->offs[P];
    if ((r = espic_migrate_finish(c, sp))) return r;
    if (trace) {
        CK(cudaStreamSynchronize(c->stream));
        fprintf(stderr, "[espic_migrate rank %d] %lld out / %lld in, host ms: flags+list+pack (1 sync) %.3f  count matrix (waits for the slowest rank) %.3f  "
                        "exchange enqueue %.3f  removal %.3f\n", me, sent, recv_total, t1 - t0, t2 - t1, t3 - t2, now() - t3);
    }
    if (n_sent) *n_sent = sent;
    if (n_received) *n_received = recv_total;
    return 0;
}
