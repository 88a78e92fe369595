// espic_collide.cuh -- ch4 DSMC_MEX::apply (ch4/Collisions.cpp:84-182) and Species::computeMPC (ch4/Species.cpp:228-235).
// Included at the end of espic_particles.cu.
//
// The reference bins particle POINTERS per cell (World::XtoC, ch4/World.h:88-98), then walks the cells one after another: per cell
// ng = (int)(0.5*np*np*Fn*sigma_cr_max*dt/dV + 0.5) candidate pairs are drawn, each accepted with probability
// sigma*cr/sigma_cr_max (VHS cross-section, ch4/Collisions.h:69-72) and scattered isotropically in the centre-of-mass frame
// (collide(), Bird's VHS).  Pairs of one cell depend on each other (a particle can collide twice), cells do not, so:
//   1. count particles per cell (atomics), exclusive scan, fill a per-cell index list (atomics: arbitrary order inside a cell);
//   2. one warp per cell rank-sorts its list by particle index -- that IS the reference's push_back order;
//   3. one thread per cell runs the reference's sequential pair loop on its list, reading and writing the velocities in place.
// Random numbers: Philox counters, draw q of cell c is element q&1 of block (c << 24) + (q >> 1) (oracle: orc_dsmc_mex, mode 1).
// sigma_cr_max of the next call is the maximum sigma*cr seen (atomicMax on the bit pattern: positive doubles order like integers).

__device__ __forceinline__ long long xtoc(const MeshC &m, double x, double y, double z)
{
    int i, j, k; double di, dj, dk;
    cell3(m, x, y, z, i, j, k, di, dj, dk);
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    if (k < 0) k = 0;
    return ((long long)k * (m.nj - 1) + j) * (long long)(m.ni - 1) + i;
}

__global__ void __launch_bounds__(256) k_xtoc_count(MeshC m, const double *__restrict__ x, const double *__restrict__ y,
                                                    const double *__restrict__ z, long long n, uint32_t *__restrict__ cnt)
{
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i < n) atomicAdd(&cnt[xtoc(m, x[i], y[i], z[i])], 1u);
}

__global__ void __launch_bounds__(256) k_mpc_store(long long nc, const uint32_t *__restrict__ cnt, double *__restrict__ mpc)
{
    const long long c = blockIdx.x * 256ll + threadIdx.x;
    if (c < nc) mpc[c] = (double)cnt[c];
}

__global__ void __launch_bounds__(256) k_xtoc_fill(MeshC m, const double *__restrict__ x, const double *__restrict__ y,
                                                   const double *__restrict__ z, long long n, const uint32_t *__restrict__ pre,
                                                   const uint32_t *__restrict__ coff, uint32_t *__restrict__ cursor,
                                                   uint32_t *__restrict__ list)
{
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= n) return;
    const long long c = xtoc(m, x[i], y[i], z[i]);
    list[scan_at(pre, coff, c) + atomicAdd(&cursor[c], 1u)] = (uint32_t)i;
}

// one warp per cell: order the cell's particle ids ascending (rank = number of smaller ids)
__global__ void __launch_bounds__(256) k_cell_rank_sort(long long nc, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ pre,
                                                        const uint32_t *__restrict__ coff, const uint32_t *__restrict__ list,
                                                        uint32_t *__restrict__ sorted)
{
    const long long c = (blockIdx.x * 256ll + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= nc) return;
    const uint32_t np = cnt[c];
    if (np < 2) return;                       // such cells never collide
    const unsigned long long base = scan_at(pre, coff, c);
    for (uint32_t e = lane; e < np; e += 32) {
        const uint32_t v = list[base + e];
        uint32_t rank = 0;
        for (uint32_t f = 0; f < np; f++) rank += list[base + f] < v;
        sorted[base + rank] = v;
    }
}

struct DsmcPar {
    double mass, Fn, dt, dV, sigma_cr_max;
    double k0, c2, expo, c3;          // evalSigma: k0 = PI*c0*c0, pow(c2/(g*g), expo)/c3
    uint64_t seed; uint32_t stream, step;
};

__global__ void __launch_bounds__(128) k_dsmc(DsmcPar par, long long nc, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ pre,
                                              const uint32_t *__restrict__ coff, const uint32_t *__restrict__ sorted,
                                              double *vx, double *vy, double *vz, unsigned long long *out /* [0] max bits, [1] collisions */)
{
    const long long c = blockIdx.x * 128ll + threadIdx.x;
    if (c >= nc) return;
    const int np = (int)cnt[c];
    if (np < 2) return;
    const uint32_t *parts = sorted + scan_at(pre, coff, c);
    unsigned q = 0, have = 0xffffffffu;
    double u0 = 0, u1 = 0;
    auto draw = [&]() -> double {
        const unsigned blk = q >> 1;
        if (blk != have) { philox_uniform2(par.seed, par.stream, par.step, ((uint64_t)c << 24) + blk, u0, u1); have = blk; }
        return (q++ & 1u) ? u1 : u0;
    };
    const double ng_f = 0.5 * np * np * par.Fn * par.sigma_cr_max * par.dt / par.dV;
    const int ng = (int)(ng_f + 0.5);
    double smax = 0;
    unsigned long long cols = 0;
    for (int g = 0; g < ng; g++) {
        const int p1 = (int)(draw() * np);
        int p2;
        do { p2 = (int)(draw() * np); } while (p2 == p1);
        const uint32_t a = parts[p1], b = parts[p2];
        const double v1[3] = {vx[a], vy[a], vz[a]}, v2[3] = {vx[b], vy[b], vz[b]};
        const double r0 = v1[0] - v2[0], r1 = v1[1] - v2[1], r2 = v1[2] - v2[2];
        const double cr = sqrt((r0 * r0 + r1 * r1) + r2 * r2);
        const double sigma = par.k0 * pow(par.c2 / (cr * cr), par.expo) / par.c3;
        const double sigma_cr = sigma * cr;
        if (sigma_cr > smax) smax = sigma_cr;
        const double P = sigma_cr / par.sigma_cr_max;
        if (P > draw()) {
            cols++;
            // DSMC_MEX::collide (ch4/Collisions.cpp:84-106), equal masses
            const double msum = par.mass + par.mass;
            const double cm0 = (v1[0] * par.mass + v2[0] * par.mass) / msum;
            const double cm1 = (v1[1] * par.mass + v2[1] * par.mass) / msum;
            const double cm2 = (v1[2] * par.mass + v2[2] * par.mass) / msum;
            const double cos_chi = 2 * draw() - 1;
            const double sin_chi = sqrt(1 - cos_chi * cos_chi);
            const double eps = 2 * 3.141592653 * draw();
            const double c0 = cr * cos_chi, c1 = cr * sin_chi * cos(eps), c2 = cr * sin_chi * sin(eps);
            const double f2 = par.mass / msum;
            vx[a] = cm0 + c0 * f2; vy[a] = cm1 + c1 * f2; vz[a] = cm2 + c2 * f2;
            vx[b] = cm0 - c0 * f2; vy[b] = cm1 - c1 * f2; vz[b] = cm2 - c2 * f2;
        }
    }
    if (smax > 0) atomicMax(out, (unsigned long long)__double_as_longlong(smax));
    if (cols) atomicAdd(out + 1, cols);
}

// bins of World::XtoC: counts in c->cell_cnt, scan in c->scan_pre / c->scan_coff
static int bin_by_cell(espic_ctx *c, Species &s, long long nc)
{
    int r;
    if ((r = ensure_buf(&c->cell_cnt, &c->cell_cap, nc, c->stream))) return r;
    CK(cudaMemsetAsync(c->cell_cnt, 0, (size_t)nc * sizeof(uint32_t), c->stream));
    if (s.np > 0) {
        k_xtoc_count<<<nblk(s.np, 256), 256, 0, c->stream>>>(c->m, s.p[0], s.p[1], s.p[2], s.np, c->cell_cnt);
        LAUNCH_CHECK(c);
    }
    return 0;
}

extern "C" int espic_compute_mpc(espic_ctx *c, int sp)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    Species &s = c->sp[sp];
    const long long nc = (long long)(c->m.ni - 1) * (c->m.nj - 1) * (c->m.nk - 1);
    if (!s.mpc) CK(cudaMalloc(&s.mpc, (size_t)nc * sizeof(double)));
    int r;
    if ((r = bin_by_cell(c, s, nc))) return r;
    k_mpc_store<<<nblk(nc, 256), 256, 0, c->stream>>>(nc, c->cell_cnt, s.mpc);
    LAUNCH_CHECK(c);
    return 0;
}

extern "C" int espic_dsmc_mex(espic_ctx *c, int sp, double dt, double *sigma_cr_max, uint64_t seed, uint32_t stream, uint32_t step,
                              long long *num_cols)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    Species &s = c->sp[sp];
    if (num_cols) *num_cols = 0;
    MIG_GUARD(c, s, "espic_dsmc_mex");
    s.diag_valid = false;
    const long long n = s.np;
    if (n < 2) return 0;
    if (n >= (1ll << 32)) { espic_set_error("espic_dsmc_mex: more than 2^32 particles in one species"); return -1; }
    if (!sigma_cr_max || !(*sigma_cr_max > 0)) { espic_set_error("espic_dsmc_mex: sigma_cr_max must be positive"); return -1; }
    const long long nc = (long long)(c->m.ni - 1) * (c->m.nj - 1) * (c->m.nk - 1);
    int r;
    if ((r = bin_by_cell(c, s, nc))) return r;
    if ((r = espic_scan_u32(c, c->cell_cnt, nc, c->dscal + 6))) return r;
    if ((r = ensure_buf(&c->dead_words, &c->dead_words_cap, nc, c->stream))) return r;          // per-cell fill cursors
    CK(cudaMemsetAsync(c->dead_words, 0, (size_t)nc * sizeof(uint32_t), c->stream));
    if ((r = ensure_buf(&c->lists, &c->lists_cap, n, c->stream))) return r;                     // n x 8 B = two uint32 lists
    uint32_t *list = reinterpret_cast<uint32_t *>(c->lists), *sorted = list + n;
    k_xtoc_fill<<<nblk(n, 256), 256, 0, c->stream>>>(c->m, s.p[0], s.p[1], s.p[2], n, c->scan_pre, c->scan_coff, c->dead_words, list);
    LAUNCH_CHECK(c);
    k_cell_rank_sort<<<nblk(nc * 32, 256), 256, 0, c->stream>>>(nc, c->cell_cnt, c->scan_pre, c->scan_coff, list, sorted);
    LAUNCH_CHECK(c);
    DsmcPar p;
    p.mass = s.mass; p.Fn = s.mpw0; p.dt = dt;
    p.dV = c->m.dh[0] * c->m.dh[1] * c->m.dh[2];          // World::getCellVolume
    p.sigma_cr_max = *sigma_cr_max;
    // DSMC_MEX constructor (ch4/Collisions.h:61-67): Bird's VHS reference parameters at 273.15 K
    const double mr = s.mass * s.mass / (s.mass + s.mass);
    const double c0 = 4.07e-10, c1 = 0.77;
    p.k0 = 3.141592653 * c0 * c0;
    p.c2 = 2 * 1.380648e-23 * 273.15 / mr;
    p.expo = c1 - 0.5;
    p.c3 = tgamma(2.5 - c1);
    p.seed = seed; p.stream = stream; p.step = step;
    CK(cudaMemsetAsync(c->dscal + 4, 0, 2 * sizeof(unsigned long long), c->stream));
    k_dsmc<<<nblk(nc, 128), 128, 0, c->stream>>>(p, nc, c->cell_cnt, c->scan_pre, c->scan_coff, sorted, s.p[3], s.p[4], s.p[5], c->dscal + 4);
    LAUNCH_CHECK(c);
    unsigned long long *h = (unsigned long long *)c->hpin;
    CK(cudaMemcpyAsync(h + 4, c->dscal + 4, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (h[5]) {                                           // if (num_cols) sigma_cr_max = sigma_cr_max_temp
        double v;
        memcpy(&v, &h[4], sizeof(double));
        *sigma_cr_max = v;
    }
    if (num_cols) *num_cols = (long long)h[5];
    return 0;
}

// =====================================================================================================
// MCC_CEX::apply (ch4/Collisions.cpp:43-82): every source particle collides with the mesh-averaged target gas with
// probability P = 1 - exp(-n*sigma*|v - u|*dt), n and u gathered from the target's density and stream-velocity fields
// (Field::gather, ch4/Field.h:234-256), sigma = 1e-16 m^2; a colliding particle's velocity becomes zero (the reference has the
// charge-exchange assignment itself commented out, :78-79).  Philox: particle i compares against element 0 of block i.
// =====================================================================================================

// trilinear gather of component c of a node array with `stride` doubles per node, the reference's term order
__device__ __forceinline__ double gather_node_field(const MeshC &m, const double *__restrict__ f, int stride, int c,
                                                    int i, int j, int k, double di, double dj, double dk)
{
    const long long u = node_u(m, i, j, k), sj = m.ni, sk = (long long)m.ni * m.nj;
    const double ai = 1 - di, aj = 1 - dj, ak = 1 - dk;
    double v = f[stride * u + c] * ai * aj * ak;
    v = v + f[stride * (u + 1) + c] * di * aj * ak;
    v = v + f[stride * (u + 1 + sj) + c] * di * dj * ak;
    v = v + f[stride * (u + sj) + c] * ai * dj * ak;
    v = v + f[stride * (u + sk) + c] * ai * aj * dk;
    v = v + f[stride * (u + 1 + sk) + c] * di * aj * dk;
    v = v + f[stride * (u + 1 + sj + sk) + c] * di * dj * dk;
    v = v + f[stride * (u + sj + sk) + c] * ai * dj * dk;
    return v;
}

__global__ void __launch_bounds__(256) k_mcc_cex(MeshC m, const double *__restrict__ x, const double *__restrict__ y,
                                                 const double *__restrict__ z, double *vx, double *vy, double *vz, long long n,
                                                 const double *__restrict__ tden, const double *__restrict__ tvel, double dt,
                                                 uint64_t seed, uint32_t stream, uint32_t step, unsigned long long *ncols)
{
    const long long q = blockIdx.x * 256ll + threadIdx.x;
    bool hit = false;
    if (q < n) {
        int i, j, k; double di, dj, dk;
        cell3(m, x[q], y[q], z[q], i, j, k, di, dj, dk);
        if (i < 0) i = 0;
        if (j < 0) j = 0;
        if (k < 0) k = 0;
        const double ut0 = gather_node_field(m, tvel, 3, 0, i, j, k, di, dj, dk);
        const double ut1 = gather_node_field(m, tvel, 3, 1, i, j, k, di, dj, dk);
        const double ut2 = gather_node_field(m, tvel, 3, 2, i, j, k, di, dj, dk);
        const double nn = gather_node_field(m, tden, 1, 0, i, j, k, di, dj, dk);
        const double r0 = vx[q] - ut0, r1 = vy[q] - ut1, r2 = vz[q] - ut2;
        const double v_rel_mag = sqrt((r0 * r0 + r1 * r1) + r2 * r2);
        const double sigma = 1e-16;
        const double P = 1 - exp(-nn * sigma * v_rel_mag * dt);
        double u0, u1;
        philox_uniform2(seed, stream, step, (uint64_t)q, u0, u1);
        if (P >= u0) { vx[q] = 0; vy[q] = 0; vz[q] = 0; hit = true; }
    }
    const unsigned b = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(ncols, (unsigned long long)__popc(b));
}

extern "C" int espic_mcc_cex(espic_ctx *c, int source_sp, int target_sp, double dt, uint64_t seed, uint32_t stream, uint32_t step,
                             long long *num_cols)
{
    SP_CHECK(c, source_sp);
    SP_CHECK(c, target_sp);
    CK(cudaSetDevice(c->device));
    Species &s = c->sp[source_sp];
    if (num_cols) *num_cols = 0;
    s.diag_valid = false;
    if (s.np == 0) return 0;
    int r;
    if ((r = espic_ensure_moments(c, target_sp))) return r;       // stream velocity: zero until computeGasProperties ran
    Species &t = c->sp[target_sp];
    CK(cudaMemsetAsync(c->dscal + 7, 0, sizeof(unsigned long long), c->stream));
    k_mcc_cex<<<nblk(s.np, 256), 256, 0, c->stream>>>(c->m, s.p[0], s.p[1], s.p[2], s.p[3], s.p[4], s.p[5], s.np, t.den,
                                                      t.mom + 7 * c->m.nn, dt, seed, stream, step, c->dscal + 7);
    LAUNCH_CHECK(c);
    unsigned long long *h = (unsigned long long *)c->hpin;
    CK(cudaMemcpyAsync(h + 7, c->dscal + 7, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (num_cols) *num_cols = (long long)h[7];
    return 0;
}
