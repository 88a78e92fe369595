/*
 * espic_oracle.c -- CPU ORACLE: test infrastructure only (see espic_oracle.h).
 *
 * A plain-C restatement, on flat arrays, of the algorithms in the reference
 * (particleincell/plasma-simulations-by-example).  Operation order inside every floating
 * point expression follows the reference so results are bit-identical to the reference
 * built with g++ -O2 on x86-64 (no FMA contraction: build with -ffp-contract=off).
 *
 * NOT on the product path.  The product is the CUDA library behind include/espic.h.
 */
#include "espic_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define U(m, i, j, k) ((size_t)(k) * (size_t)(m)->ni * (size_t)(m)->nj + (size_t)(j) * (size_t)(m)->ni + (size_t)(i))

/* ------------------------------------------------------------------ RNG */

/* mt19937 (Matsumoto & Nishimura 1998) as std::mt19937; seeded like std::mt19937(seed). */
void orc_mt_seed(orc_mt19937 *g, uint32_t seed)
{
    g->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

uint32_t orc_mt_next(orc_mt19937 *g)
{
    if (g->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            uint32_t v = g->mt[(i + 397) % 624] ^ (y >> 1);
            if (y & 1u) v ^= 0x9908b0dfu;
            g->mt[i] = v;
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* std::uniform_real_distribution<double>(0,1.0) over mt19937 in libstdc++ (the reference's
 * Rnd::operator(), ch3/ver2/World.h:24-33): generate_canonical<double,53> draws two 32-bit
 * words, sum = w0 + w1*2^32 (rounded to double), divides by 2^64, clamps below 1. */
double orc_mt_uniform(orc_mt19937 *g)
{
    double sum = 0.0, tmp = 1.0;
    const double r = 4294967296.0;
    for (int k = 0; k < 2; k++) {
        sum += (double)orc_mt_next(g) * tmp;
        tmp *= r;
    }
    double ret = sum / tmp;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret;
}

/* Philox4x32-10 (Salmon et al., SC'11), the counter-based generator the CUDA injector uses. */
void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox_uniform2(uint64_t seed, uint32_t stream, uint32_t step, uint64_t idx, double out[2])
{
    uint32_t ctr[4] = { (uint32_t)idx, (uint32_t)(idx >> 32), step, stream };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t r[4];
    orc_philox4x32_10(ctr, key, r);
    uint64_t a = ((uint64_t)r[1] << 32) | r[0];
    uint64_t b = ((uint64_t)r[3] << 32) | r[2];
    out[0] = (double)(a >> 11) * (1.0 / 9007199254740992.0);
    out[1] = (double)(b >> 11) * (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------------ mesh */

/* World::World + World::setExtents, ch3/ver2/World.cpp:14-36 */
void orc_mesh_init(orc_mesh *m, int ni, int nj, int nk, const double x0[3], const double xm[3])
{
    m->ni = ni; m->nj = nj; m->nk = nk;
    int nn[3] = { ni, nj, nk };
    for (int c = 0; c < 3; c++) {
        m->x0[c] = x0[c];
        m->xm[c] = xm[c];
        m->dh[c] = (xm[c] - x0[c]) / (nn[c] - 1);
        m->xc[c] = (x0[c] + xm[c]) * 0.5;
        m->sphere_c[c] = 0;
    }
    m->sphere_r2 = 0;
}

/* World::computeNodeVolumes, ch3/ver2/World.cpp:58-69 */
void orc_node_volumes(const orc_mesh *m, double *node_vol)
{
    for (int i = 0; i < m->ni; i++)
        for (int j = 0; j < m->nj; j++)
            for (int k = 0; k < m->nk; k++) {
                double V = m->dh[0] * m->dh[1] * m->dh[2];
                if (i == 0 || i == m->ni - 1) V *= 0.5;
                if (j == 0 || j == m->nj - 1) V *= 0.5;
                if (k == 0 || k == m->nk - 1) V *= 0.5;
                node_vol[U(m, i, j, k)] = V;
            }
}

/* World::inSphere, ch3/ver2/World.cpp:118-125 */
int orc_in_sphere(const orc_mesh *m, const double x[3])
{
    double r0 = x[0] - m->sphere_c[0];
    double r1 = x[1] - m->sphere_c[1];
    double r2 = x[2] - m->sphere_c[2];
    double r_mag2 = (r0 * r0 + r1 * r1 + r2 * r2);
    return r_mag2 <= m->sphere_r2;
}

/* World::inBounds, ch3/ver2/World.h:59-63 */
int orc_in_bounds(const orc_mesh *m, const double pos[3])
{
    for (int c = 0; c < 3; c++)
        if (pos[c] < m->x0[c] || pos[c] >= m->xm[c]) return 0;
    return 1;
}

/* World::addSphere, ch3/ver2/World.cpp:87-105 (node position from World::pos, World.h:84-94) */
void orc_add_sphere(orc_mesh *m, const double c[3], double radius, double phi_sphere, int32_t *object_id, double *phi)
{
    for (int a = 0; a < 3; a++) m->sphere_c[a] = c[a];
    m->sphere_r2 = radius * radius;
    for (int i = 0; i < m->ni; i++)
        for (int j = 0; j < m->nj; j++)
            for (int k = 0; k < m->nk; k++) {
                double x[3];
                x[0] = m->x0[0] + m->dh[0] * (double)i;
                x[1] = m->x0[1] + m->dh[1] * (double)j;
                x[2] = m->x0[2] + m->dh[2] * (double)k;
                if (orc_in_sphere(m, x)) {
                    object_id[U(m, i, j, k)] = 1;
                    phi[U(m, i, j, k)] = phi_sphere;
                }
            }
}

/* World::addInlet, ch3/ver2/World.cpp:108-115 */
void orc_add_inlet(const orc_mesh *m, int32_t *object_id, double *phi)
{
    for (int i = 0; i < m->ni; i++)
        for (int j = 0; j < m->nj; j++) {
            object_id[U(m, i, j, 0)] = 2;
            phi[U(m, i, j, 0)] = 0;
        }
}

/* World::XtoL, ch3/ver2/World.h:75-81 (true division) */
void orc_xtol(const orc_mesh *m, const double x[3], double lc[3])
{
    lc[0] = (x[0] - m->x0[0]) / m->dh[0];
    lc[1] = (x[1] - m->x0[1]) / m->dh[1];
    lc[2] = (x[2] - m->x0[2]) / m->dh[2];
}

/* cell index + fraction.  The reference truncates (int)lc and reads node i+1 even when
 * lc rounds to exactly n-1 (out-of-bounds read times a zero weight, Field.h:189-208).
 * Oracle and CUDA both clamp the cell to n-2 there (fraction becomes exactly 1), which
 * yields the same value whenever the reference's phantom read is finite. */
static inline void cell_of(double lc, int n, int *i, double *d)
{
    int ii = (int)lc;
    if (ii > n - 2) ii = n - 2;
    *i = ii;
    *d = lc - ii;
}

/* Field3::gather, ch3/ver2/Field.h:189-211: eight terms, each data*w_i*w_j*w_k left to right */
void orc_gather3(const orc_mesh *m, const double *ef, const double lc[3], double out[3])
{
    int i, j, k; double di, dj, dk;
    cell_of(lc[0], m->ni, &i, &di);
    cell_of(lc[1], m->nj, &j, &dj);
    cell_of(lc[2], m->nk, &k, &dk);
    const size_t n[8] = { U(m, i, j, k), U(m, i + 1, j, k), U(m, i + 1, j + 1, k), U(m, i, j + 1, k),
                          U(m, i, j, k + 1), U(m, i + 1, j, k + 1), U(m, i + 1, j + 1, k + 1), U(m, i, j + 1, k + 1) };
    const double wi[8] = { 1 - di, di, di, 1 - di, 1 - di, di, di, 1 - di };
    const double wj[8] = { 1 - dj, 1 - dj, dj, dj, 1 - dj, 1 - dj, dj, dj };
    const double wk[8] = { 1 - dk, 1 - dk, 1 - dk, 1 - dk, dk, dk, dk, dk };
    for (int c = 0; c < 3; c++) {
        double val = ef[3 * n[0] + c] * wi[0] * wj[0] * wk[0];
        for (int t = 1; t < 8; t++)
            val = val + ef[3 * n[t] + c] * wi[t] * wj[t] * wk[t];
        out[c] = val;
    }
}

/* Field::scatter, ch3/ver2/Field.h:167-186 */
void orc_scatter(const orc_mesh *m, double *f, const double lc[3], double value)
{
    int i, j, k; double di, dj, dk;
    cell_of(lc[0], m->ni, &i, &di);
    cell_of(lc[1], m->nj, &j, &dj);
    cell_of(lc[2], m->nk, &k, &dk);
    f[U(m, i, j, k)]             += value * (1 - di) * (1 - dj) * (1 - dk);
    f[U(m, i + 1, j, k)]         += value * (di) * (1 - dj) * (1 - dk);
    f[U(m, i + 1, j + 1, k)]     += value * (di) * (dj) * (1 - dk);
    f[U(m, i, j + 1, k)]         += value * (1 - di) * (dj) * (1 - dk);
    f[U(m, i, j, k + 1)]         += value * (1 - di) * (1 - dj) * (dk);
    f[U(m, i + 1, j, k + 1)]     += value * (di) * (1 - dj) * (dk);
    f[U(m, i + 1, j + 1, k + 1)] += value * (di) * (dj) * (dk);
    f[U(m, i, j + 1, k + 1)]     += value * (1 - di) * (dj) * (dk);
}

/* ------------------------------------------------------------------ particles */

static inline void leapfrog(const orc_mesh *m, const double *ef, orc_particles *p, int64_t q, double s, double dt)
{
    double pos[3] = { p->x[q], p->y[q], p->z[q] };
    double lc[3], e[3];
    orc_xtol(m, pos, lc);
    orc_gather3(m, ef, lc, e);
    /* part.vel += ef_part*(dt*charge/mass);  Species.cpp:22 */
    p->vx[q] += e[0] * s;
    p->vy[q] += e[1] * s;
    p->vz[q] += e[2] * s;
    /* part.pos += part.vel*dt;  Species.cpp:25 */
    p->x[q] += p->vx[q] * dt;
    p->y[q] += p->vy[q] * dt;
    p->z[q] += p->vz[q] * dt;
}

/* first half of Species::advance, ch3/ver2/Species.cpp:7-33 */
void orc_push_sphere_nocompact(const orc_mesh *m, const double *ef, orc_particles *p, double charge, double mass, double dt)
{
    const double s = dt * charge / mass;
    for (int64_t q = 0; q < p->np; q++) {
        leapfrog(m, ef, p, q, s, dt);
        double pos[3] = { p->x[q], p->y[q], p->z[q] };
        if (orc_in_sphere(m, pos) || !orc_in_bounds(m, pos))
            p->mpw[q] = 0;
    }
}

static inline void copy_particle(orc_particles *p, int64_t dst, int64_t src)
{
    p->x[dst] = p->x[src]; p->y[dst] = p->y[src]; p->z[dst] = p->z[src];
    p->vx[dst] = p->vx[src]; p->vy[dst] = p->vy[src]; p->vz[dst] = p->vz[src];
    p->mpw[dst] = p->mpw[src];
}

/* Species::advance, ch3/ver2/Species.cpp:7-48 (push, kill, swap-with-last removal) */
int64_t orc_advance_sphere(const orc_mesh *m, const double *ef, orc_particles *p, double charge, double mass, double dt)
{
    orc_push_sphere_nocompact(m, ef, p, charge, mass, dt);
    int64_t np = p->np;
    for (int64_t q = 0; q < np; q++) {
        if (p->mpw[q] > 0) continue;
        copy_particle(p, q, np - 1);
        np--;
        q--;
    }
    p->np = np;
    return np;
}

/* Species::advance of the grounded box, ch2/Species.cpp:7-38 (specular reflection) */
void orc_advance_box(const orc_mesh *m, const double *ef, orc_particles *p, double charge, double mass, double dt)
{
    const double s = dt * charge / mass;
    for (int64_t q = 0; q < p->np; q++) {
        leapfrog(m, ef, p, q, s, dt);
        double *pos[3] = { &p->x[q], &p->y[q], &p->z[q] };
        double *vel[3] = { &p->vx[q], &p->vy[q], &p->vz[q] };
        for (int c = 0; c < 3; c++) {
            if (*pos[c] < m->x0[c]) { *pos[c] = 2 * m->x0[c] - *pos[c]; *vel[c] *= -1.0; }
            else if (*pos[c] >= m->xm[c]) { *pos[c] = 2 * m->xm[c] - *pos[c]; *vel[c] *= -1.0; }
        }
    }
}

/* Species::addParticle, ch3/ver2/Species.cpp:65-81.  Returns 1 if added. */
int orc_add_particle(const orc_mesh *m, const double *ef, orc_particles *p, const double pos[3], const double vel_in[3],
                     double mpw, double charge, double mass, double dt)
{
    if (!orc_in_bounds(m, pos)) return 0;
    if (p->np >= p->cap) return 0;
    double lc[3], e[3];
    orc_xtol(m, pos, lc);
    orc_gather3(m, ef, lc, e);
    /* vel -= charge/mass*ef_part*(0.5*world.getDt()); */
    const double qm = charge / mass;
    const double hdt = 0.5 * dt;
    int64_t q = p->np++;
    p->x[q] = pos[0]; p->y[q] = pos[1]; p->z[q] = pos[2];
    p->vx[q] = vel_in[0] - e[0] * qm * hdt;
    p->vy[q] = vel_in[1] - e[1] * qm * hdt;
    p->vz[q] = vel_in[2] - e[2] * qm * hdt;
    p->mpw[q] = mpw;
    return 1;
}

/* Species::computeNumberDensity, ch3/ver2/Species.cpp:51-62 ; Field::operator/=, Field.h:125-134 */
void orc_number_density(const orc_mesh *m, const double *node_vol, const orc_particles *p, double *den)
{
    size_t nn = (size_t)m->ni * m->nj * m->nk;
    memset(den, 0, nn * sizeof(double));
    for (int64_t q = 0; q < p->np; q++) {
        double pos[3] = { p->x[q], p->y[q], p->z[q] }, lc[3];
        orc_xtol(m, pos, lc);
        orc_scatter(m, den, lc, p->mpw[q]);
    }
    for (size_t u = 0; u < nn; u++) {
        if (node_vol[u] != 0) den[u] /= node_vol[u];
        else den[u] = 0;
    }
}

/* Species::sampleMoments, ch4/Species.cpp:190-200: n_sum, nv_sum (3 interleaved components), nuu_sum, nvv_sum, nww_sum
 * accumulate the trilinear scatter of mpw, mpw*vel, mpw*vx*vx, mpw*vy*vy, mpw*vz*vz (products left to right). */
void orc_sample_moments(const orc_mesh *m, const orc_particles *p, double *n_sum, double *nv_sum, double *nuu_sum,
                        double *nvv_sum, double *nww_sum)
{
    size_t nn = (size_t)m->ni * m->nj * m->nk;
    /* Field3::scatter adds value*(w_i)*(w_j)*(w_k) component-wise: scatter each component into a strided view */
    double *tmp = (double *)calloc(nn, sizeof(double));
    for (int c = 0; c < 3; c++) {
        memset(tmp, 0, nn * sizeof(double));
        for (size_t u = 0; u < nn; u++) tmp[u] = nv_sum[3 * u + c];
        for (int64_t q = 0; q < p->np; q++) {
            double pos[3] = { p->x[q], p->y[q], p->z[q] }, lc[3];
            const double v = c == 0 ? p->vx[q] : (c == 1 ? p->vy[q] : p->vz[q]);
            orc_xtol(m, pos, lc);
            orc_scatter(m, tmp, lc, p->mpw[q] * v);
        }
        for (size_t u = 0; u < nn; u++) nv_sum[3 * u + c] = tmp[u];
    }
    free(tmp);
    for (int64_t q = 0; q < p->np; q++) {
        double pos[3] = { p->x[q], p->y[q], p->z[q] }, lc[3];
        orc_xtol(m, pos, lc);
        orc_scatter(m, n_sum, lc, p->mpw[q]);
        orc_scatter(m, nuu_sum, lc, p->mpw[q] * p->vx[q] * p->vx[q]);
        orc_scatter(m, nvv_sum, lc, p->mpw[q] * p->vy[q] * p->vy[q]);
        orc_scatter(m, nww_sum, lc, p->mpw[q] * p->vz[q] * p->vz[q]);
    }
}

/* Species::computeGasProperties, ch4/Species.cpp:203-226 with Field operator/ (ch4/Field.h:192-204):
 * vel = nv_sum/n_sum (0 where n_sum == 0); T = mass/(2K) * ((<u2>-<u>^2) + (<v2>-<v>^2) + (<w2>-<w>^2)), 0 where count <= 0 */
void orc_gas_properties(const orc_mesh *m, double mass, const double *n_sum, const double *nv_sum, const double *nuu_sum,
                        const double *nvv_sum, const double *nww_sum, double *vel, double *T)
{
    const double K = 1.380648e-23;      /* Const::K, ch4/World.h:18 */
    size_t nn = (size_t)m->ni * m->nj * m->nk;
    for (size_t u = 0; u < nn; u++) {
        for (int c = 0; c < 3; c++) vel[3 * u + c] = (n_sum[u] != 0) ? nv_sum[3 * u + c] / n_sum[u] : 0.0;
        double count = n_sum[u];
        if (count <= 0) { T[u] = 0; continue; }
        double u_ave = vel[3 * u], v_ave = vel[3 * u + 1], w_ave = vel[3 * u + 2];
        double u2_ave = nuu_sum[u] / count, v2_ave = nvv_sum[u] / count, w2_ave = nww_sum[u] / count;
        double uu = u2_ave - u_ave * u_ave, vv = v2_ave - v_ave * v_ave, ww = w2_ave - w_ave * w_ave;
        T[u] = mass / (2 * K) * (uu + vv + ww);
    }
}

/* World::computeChargeDensity, ch3/ver2/World.cpp:46-54: rho=0; rho += charge*den (charge!=0) */
void orc_rho_clear(const orc_mesh *m, double *rho)
{
    memset(rho, 0, (size_t)m->ni * m->nj * m->nk * sizeof(double));
}

void orc_rho_add(const orc_mesh *m, double *rho, const double *den, double charge)
{
    if (charge == 0) return;
    size_t nn = (size_t)m->ni * m->nj * m->nk;
    for (size_t u = 0; u < nn; u++) rho[u] += den[u] * charge;
}

/* Species::getRealCount / getMomentum / getKE, ch3/ver2/Species.cpp:84-108 */
double orc_real_count(const orc_particles *p)
{
    double s = 0;
    for (int64_t q = 0; q < p->np; q++) s += p->mpw[q];
    return s;
}

void orc_momentum(const orc_particles *p, double mass, double out[3])
{
    double mx = 0, my = 0, mz = 0;
    for (int64_t q = 0; q < p->np; q++) {
        mx += p->vx[q] * p->mpw[q];
        my += p->vy[q] * p->mpw[q];
        mz += p->vz[q] * p->mpw[q];
    }
    out[0] = mx * mass; out[1] = my * mass; out[2] = mz * mass;
}

double orc_ke(const orc_particles *p, double mass)
{
    double ke = 0;
    for (int64_t q = 0; q < p->np; q++) {
        double v2 = p->vx[q] * p->vx[q] + p->vy[q] * p->vy[q] + p->vz[q] * p->vz[q];
        ke += p->mpw[q] * v2;
    }
    return 0.5 * mass * ke;
}

/* World::getPE, ch3/ver2/World.cpp:72-84 */
double orc_pe(const orc_mesh *m, const double *ef, const double *node_vol)
{
    double pe = 0;
    for (int i = 0; i < m->ni; i++)
        for (int j = 0; j < m->nj; j++)
            for (int k = 0; k < m->nk; k++) {
                size_t u = U(m, i, j, k);
                double ef2 = ef[3 * u] * ef[3 * u] + ef[3 * u + 1] * ef[3 * u + 1] + ef[3 * u + 2] * ef[3 * u + 2];
                pe += ef2 * node_vol[u];
            }
    return 0.5 * ORC_EPS_0 * pe;
}

/* Field::updateAverage, ch3/ver2/Field.h:214-221 */
void orc_update_average(const orc_mesh *m, double *ave, const double *inst, int *ave_samples)
{
    size_t nn = (size_t)m->ni * m->nj * m->nk;
    int s = *ave_samples;
    for (size_t u = 0; u < nn; u++) ave[u] = (inst[u] + s * ave[u]) / (s + 1);
    *ave_samples = s + 1;
}

/* ------------------------------------------------------------------ sources / loaders */

/* ColdBeamSource::sample, ch3/ver2/Source.cpp:4-18: number of macroparticles for uniform u */
int64_t orc_cold_beam_num_sim(const orc_mesh *m, double den, double v_drift, double dt, double mpw0, double u)
{
    double Lx = m->dh[0] * (m->ni - 1);
    double Ly = m->dh[1] * (m->nj - 1);
    double A = Lx * Ly;
    double num_real = den * v_drift * A * dt;
    return (int)(num_real / mpw0 + u);
}

/* ColdBeamSource::sample with the reference's RNG, ch3/ver2/Source.cpp:4-27 */
int64_t orc_cold_beam_sample_mt(const orc_mesh *m, const double *ef, orc_particles *p, double charge, double mass,
                                double mpw0, double v_drift, double den, double dt, orc_mt19937 *g)
{
    double Lx = m->dh[0] * (m->ni - 1);
    double Ly = m->dh[1] * (m->nj - 1);
    int64_t num_sim = orc_cold_beam_num_sim(m, den, v_drift, dt, mpw0, orc_mt_uniform(g));
    int64_t added = 0;
    for (int64_t i = 0; i < num_sim; i++) {
        double pos[3];
        pos[0] = m->x0[0] + orc_mt_uniform(g) * Lx;
        pos[1] = m->x0[1] + orc_mt_uniform(g) * Ly;
        pos[2] = m->x0[2];
        double vel[3] = { 0, 0, v_drift };
        added += orc_add_particle(m, ef, p, pos, vel, mpw0, charge, mass, dt);
    }
    return added;
}

/* Same sampler, uniforms from Philox4x32-10.  Draw layout (shared with the CUDA injector):
 *   idx = 2^64-1            -> out[0] is the Bernoulli fraction for num_sim
 *   idx = particle number i -> (out[0], out[1]) are the x and y uniforms               */
int64_t orc_cold_beam_sample_philox(const orc_mesh *m, const double *ef, orc_particles *p, double charge, double mass,
                                    double mpw0, double v_drift, double den, double dt,
                                    uint64_t seed, uint32_t stream, uint32_t step)
{
    double Lx = m->dh[0] * (m->ni - 1);
    double Ly = m->dh[1] * (m->nj - 1);
    double u2[2];
    orc_philox_uniform2(seed, stream, step, ~(uint64_t)0, u2);
    int64_t num_sim = orc_cold_beam_num_sim(m, den, v_drift, dt, mpw0, u2[0]);
    int64_t added = 0;
    for (int64_t i = 0; i < num_sim; i++) {
        orc_philox_uniform2(seed, stream, step, (uint64_t)i, u2);
        double pos[3];
        pos[0] = m->x0[0] + u2[0] * Lx;
        pos[1] = m->x0[1] + u2[1] * Ly;
        pos[2] = m->x0[2];
        double vel[3] = { 0, 0, v_drift };
        added += orc_add_particle(m, ef, p, pos, vel, mpw0, charge, mass, dt);
    }
    return added;
}

/* Species::sampleIsotropicVel + sampleVth, ch4/Species.cpp:149-173, from eleven uniforms in the reference's draw order:
 * u[0] theta, u[1] direction cosine, u[2..10] three sums of three (Birdsall) */
void orc_isotropic_vel(double T, double mass, const double u[11], double vel[3])
{
    const double K = 1.380648e-23, PI = 3.141592653;       /* Const::K, Const::PI (ch4/World.h:18-19) */
    double theta = 2 * PI * u[0];
    double r = -1.0 + 2 * u[1];
    double a = sqrt(1 - r * r);
    double d[3] = { r, cos(theta) * a, sin(theta) * a };
    double v_th = sqrt(2 * K * T / mass);
    double v1 = v_th * (u[2] + u[3] + u[4] - 1.5);
    double v2 = v_th * (u[5] + u[6] + u[7] - 1.5);
    double v3 = v_th * (u[8] + u[9] + u[10] - 1.5);
    double mag = 3 / sqrt(2 + 2 + 2) * sqrt(v1 * v1 + v2 * v2 + v3 * v3);
    for (int c = 0; c < 3; c++) vel[c] = d[c] * mag;         /* v_th*d: vec3 scalar multiplication a(c)*s */
}

/* WarmBeamSource::sample, ch4/Source.cpp:31-56, uniforms from mt19937 in the reference's call order */
int64_t orc_warm_beam_sample_mt(const orc_mesh *m, const double *ef, orc_particles *p, double charge, double mass,
                                double mpw0, double v_drift, double den, double T, double dt, orc_mt19937 *g)
{
    double Lx = m->dh[0] * (m->ni - 1);
    double Ly = m->dh[1] * (m->nj - 1);
    int64_t num_sim = orc_cold_beam_num_sim(m, den, v_drift, dt, mpw0, orc_mt_uniform(g));
    int64_t added = 0;
    for (int64_t i = 0; i < num_sim; i++) {
        double pos[3], vel[3], u[11];
        pos[0] = m->x0[0] + orc_mt_uniform(g) * Lx;
        pos[1] = m->x0[1] + orc_mt_uniform(g) * Ly;
        pos[2] = m->x0[2];
        for (int q = 0; q < 11; q++) u[q] = orc_mt_uniform(g);
        orc_isotropic_vel(T, mass, u, vel);
        vel[2] += v_drift;
        added += orc_add_particle(m, ef, p, pos, vel, mpw0, charge, mass, dt);
    }
    return added;
}

/* Same sampler, uniforms from Philox4x32-10.  Draw layout (shared with the CUDA injector): idx = 2^64-1 -> Bernoulli
 * fraction; particle i uses the seven blocks idx = 8*i + j: j=0 (x, y), j=1 (theta, cosine), j=2..6 the nine Birdsall
 * uniforms in order (the second output of block 6 is unused). */
int64_t orc_warm_beam_sample_philox(const orc_mesh *m, const double *ef, orc_particles *p, double charge, double mass,
                                    double mpw0, double v_drift, double den, double T, double dt,
                                    uint64_t seed, uint32_t stream, uint32_t step)
{
    double Lx = m->dh[0] * (m->ni - 1);
    double Ly = m->dh[1] * (m->nj - 1);
    double u2[2];
    orc_philox_uniform2(seed, stream, step, ~(uint64_t)0, u2);
    int64_t num_sim = orc_cold_beam_num_sim(m, den, v_drift, dt, mpw0, u2[0]);
    int64_t added = 0;
    for (int64_t i = 0; i < num_sim; i++) {
        double w[14];
        for (int j = 0; j < 7; j++) orc_philox_uniform2(seed, stream, step, 8 * (uint64_t)i + j, w + 2 * j);
        double pos[3] = { m->x0[0] + w[0] * Lx, m->x0[1] + w[1] * Ly, m->x0[2] }, vel[3];
        orc_isotropic_vel(T, mass, w + 2, vel);
        vel[2] += v_drift;
        added += orc_add_particle(m, ef, p, pos, vel, mpw0, charge, mass, dt);
    }
    return added;
}

/* Species::loadParticlesBoxQS, ch2/Species.cpp:101-141 */
int64_t orc_load_box_qs(const orc_mesh *m, const double *ef, orc_particles *p, const double x1[3], const double x2[3],
                        double num_den, const int num_mp[3], double charge, double mass, double dt)
{
    double box_vol = (x2[0] - x1[0]) * (x2[1] - x1[1]) * (x2[2] - x1[2]);
    int num_mp_tot = (num_mp[0] - 1) * (num_mp[1] - 1) * (num_mp[2] - 1);
    double num_real = num_den * box_vol;
    double mpw = num_real / num_mp_tot;
    double di = (x2[0] - x1[0]) / (num_mp[0] - 1);
    double dj = (x2[1] - x1[1]) / (num_mp[1] - 1);
    double dk = (x2[2] - x1[2]) / (num_mp[2] - 1);
    int64_t added = 0;
    for (int i = 0; i < num_mp[0]; i++)
        for (int j = 0; j < num_mp[1]; j++)
            for (int k = 0; k < num_mp[2]; k++) {
                double pos[3];
                pos[0] = x1[0] + i * di;
                pos[1] = x1[1] + j * dj;
                pos[2] = x1[2] + k * dk;
                if (pos[0] == x2[0]) pos[0] -= 1e-4 * di;
                if (pos[1] == x2[1]) pos[1] -= 1e-4 * dj;
                if (pos[2] == x2[2]) pos[2] -= 1e-4 * dk;
                double w = 1;
                if (i == 0 || i == num_mp[0] - 1) w *= 0.5;
                if (j == 0 || j == num_mp[1] - 1) w *= 0.5;
                if (k == 0 || k == num_mp[2] - 1) w *= 0.5;
                double vel[3] = { 0, 0, 0 };
                added += orc_add_particle(m, ef, p, pos, vel, mpw * w, charge, mass, dt);
            }
    return added;
}

/* ------------------------------------------------------------------ field solvers */

/* PotentialSolver::solveQN, ch3/ver2/PotentialSolver.cpp:204-222 */
void orc_solve_qn(const orc_mesh *m, const int32_t *object_id, const double *rho, double *phi,
                  double phi0, double Te0, double n0)
{
    double rho0 = n0 * ORC_QE;
    double rho_ratio_min = 1e-6;
    size_t nn = (size_t)m->ni * m->nj * m->nk;
    for (size_t u = 0; u < nn; u++) {
        if (object_id[u] > 0) continue;
        double rho_ratio = rho[u] / rho0;
        if (rho_ratio < rho_ratio_min) rho_ratio = rho_ratio_min;
        phi[u] = phi0 + Te0 * log(rho_ratio);
    }
}

/* PotentialSolver::solveGS (nonlinear, Boltzmann electrons), ch3/ver2/PotentialSolver.cpp:334-430 */
int orc_solve_gs(const orc_mesh *m, const int32_t *object_id, const double *rho, double *phi,
                 double phi0, double Te0, double n0, unsigned max_it, double tol, orc_solve_info *info)
{
    const int ni = m->ni, nj = m->nj, nk = m->nk;
    double idx2 = 1.0 / (m->dh[0] * m->dh[0]);
    double idy2 = 1.0 / (m->dh[1] * m->dh[1]);
    double idz2 = 1.0 / (m->dh[2] * m->dh[2]);
    double L2 = 0;
    int converged = 0;
    unsigned it;
    for (it = 0; it < max_it; it++) {
        for (int i = 0; i < ni; i++)
            for (int j = 0; j < nj; j++)
                for (int k = 0; k < nk; k++) {
                    size_t u = U(m, i, j, k);
                    if (object_id[u] > 0) continue;
                    if (i == 0) phi[u] = phi[U(m, i + 1, j, k)];
                    else if (i == ni - 1) phi[u] = phi[U(m, i - 1, j, k)];
                    else if (j == 0) phi[u] = phi[U(m, i, j + 1, k)];
                    else if (j == nj - 1) phi[u] = phi[U(m, i, j - 1, k)];
                    else if (k == 0) phi[u] = phi[U(m, i, j, k + 1)];
                    else if (k == nk - 1) phi[u] = phi[U(m, i, j, k - 1)];
                    else {
                        double ne = n0 * exp((phi[u] - phi0) / Te0);
                        double phi_new = ((rho[u] - ORC_QE * ne) / ORC_EPS_0 +
                                          idx2 * (phi[U(m, i - 1, j, k)] + phi[U(m, i + 1, j, k)]) +
                                          idy2 * (phi[U(m, i, j - 1, k)] + phi[U(m, i, j + 1, k)]) +
                                          idz2 * (phi[U(m, i, j, k - 1)] + phi[U(m, i, j, k + 1)])) /
                                         (2 * idx2 + 2 * idy2 + 2 * idz2);
                        phi[u] = phi[u] + 1.4 * (phi_new - phi[u]);
                    }
                }
        if (it % 25 == 0) {
            double sum = 0;
            for (int i = 0; i < ni; i++)
                for (int j = 0; j < nj; j++)
                    for (int k = 0; k < nk; k++) {
                        size_t u = U(m, i, j, k);
                        if (object_id[u] > 0) continue;
                        double R = 0;
                        if (i == 0) R = phi[u] - phi[U(m, i + 1, j, k)];
                        else if (i == ni - 1) R = phi[u] - phi[U(m, i - 1, j, k)];
                        else if (j == 0) R = phi[u] - phi[U(m, i, j + 1, k)];
                        else if (j == nj - 1) R = phi[u] - phi[U(m, i, j - 1, k)];
                        else if (k == 0) R = phi[u] - phi[U(m, i, j, k + 1)];
                        else if (k == nk - 1) R = phi[u] - phi[U(m, i, j, k - 1)];
                        else {
                            double ne = n0 * exp((phi[u] - phi0) / Te0);
                            R = -phi[u] * (2 * idx2 + 2 * idy2 + 2 * idz2) +
                                (rho[u] - ORC_QE * ne) / ORC_EPS_0 +
                                idx2 * (phi[U(m, i - 1, j, k)] + phi[U(m, i + 1, j, k)]) +
                                idy2 * (phi[U(m, i, j - 1, k)] + phi[U(m, i, j + 1, k)]) +
                                idz2 * (phi[U(m, i, j, k - 1)] + phi[U(m, i, j, k + 1)]);
                        }
                        sum += R * R;
                    }
            L2 = sqrt(sum / (ni * nj * nk));
            if (L2 < tol) { converged = 1; break; }
        }
    }
    if (info) {
        memset(info, 0, sizeof(*info));
        info->converged = converged;
        info->gs_iters = converged ? (int64_t)it + 1 : (int64_t)it;
        info->residual = L2;
    }
    return converged;
}

/* PotentialSolver::solve of the grounded box (linear, interior nodes only), ch2/PotentialSolver.cpp:11-67 */
int orc_solve_gs_box(const orc_mesh *m, const double *rho, double *phi, unsigned max_it, double tol, orc_solve_info *info)
{
    const int ni = m->ni, nj = m->nj, nk = m->nk;
    double idx2 = 1.0 / (m->dh[0] * m->dh[0]);
    double idy2 = 1.0 / (m->dh[1] * m->dh[1]);
    double idz2 = 1.0 / (m->dh[2] * m->dh[2]);
    double L2 = 0;
    int converged = 0;
    unsigned it;
    for (it = 0; it < max_it; it++) {
        for (int i = 1; i < ni - 1; i++)
            for (int j = 1; j < nj - 1; j++)
                for (int k = 1; k < nk - 1; k++) {
                    size_t u = U(m, i, j, k);
                    double phi_new = (rho[u] / ORC_EPS_0 +
                                      idx2 * (phi[U(m, i - 1, j, k)] + phi[U(m, i + 1, j, k)]) +
                                      idy2 * (phi[U(m, i, j - 1, k)] + phi[U(m, i, j + 1, k)]) +
                                      idz2 * (phi[U(m, i, j, k - 1)] + phi[U(m, i, j, k + 1)])) /
                                     (2 * idx2 + 2 * idy2 + 2 * idz2);
                    phi[u] = phi[u] + 1.4 * (phi_new - phi[u]);
                }
        if (it % 25 == 0) {
            double sum = 0;
            for (int i = 1; i < ni - 1; i++)
                for (int j = 1; j < nj - 1; j++)
                    for (int k = 1; k < nk - 1; k++) {
                        size_t u = U(m, i, j, k);
                        double R = -phi[u] * (2 * idx2 + 2 * idy2 + 2 * idz2) +
                                   rho[u] / ORC_EPS_0 +
                                   idx2 * (phi[U(m, i - 1, j, k)] + phi[U(m, i + 1, j, k)]) +
                                   idy2 * (phi[U(m, i, j - 1, k)] + phi[U(m, i, j + 1, k)]) +
                                   idz2 * (phi[U(m, i, j, k - 1)] + phi[U(m, i, j, k + 1)]);
                        sum += R * R;
                    }
            L2 = sqrt(sum / (ni * nj * nk));
            if (L2 < tol) { converged = 1; break; }
        }
    }
    if (info) {
        memset(info, 0, sizeof(*info));
        info->converged = converged;
        info->gs_iters = converged ? (int64_t)it + 1 : (int64_t)it;
        info->residual = L2;
    }
    return converged;
}

/* --- 7-slot rows, ch3/ver2/PotentialSolver.h:8-38 : a[7*u+s], col[7*u+s] (col<0 ends the row) */

static double *mat_at(double *a, int32_t *col, int r, int c)
{
    /* Matrix::operator(), ch3/ver2/PotentialSolver.cpp:38-49 */
    int v;
    for (v = 0; v < 7; v++) {
        if (col[7 * (size_t)r + v] == c) break;
        if (col[7 * (size_t)r + v] < 0) { col[7 * (size_t)r + v] = c; break; }
    }
    if (v == 7) abort();
    return &a[7 * (size_t)r + v];
}

/* PotentialSolver::buildMatrix, ch3/ver2/PotentialSolver.cpp:146-199 (the trailing solveQN() is the caller's job) */
void orc_build_matrix(const orc_mesh *m, const int32_t *object_id, double *a, int32_t *col, int32_t *node_type)
{
    double idx = 1.0 / m->dh[0];
    double idy = 1.0 / m->dh[1];
    double idz = 1.0 / m->dh[2];
    double idx2 = idx * idx;
    double idy2 = idy * idy;
    double idz2 = idz * idz;
    const int ni = m->ni, nj = m->nj, nk = m->nk;
    for (int k = 0; k < nk; k++)
        for (int j = 0; j < nj; j++)
            for (int i = 0; i < ni; i++) {
                int u = (int)U(m, i, j, k);
                for (int s = 0; s < 7; s++) { a[7 * (size_t)u + s] = 0; col[7 * (size_t)u + s] = -1; }
                if (object_id[u] > 0) {
                    *mat_at(a, col, u, u) = 1;
                    node_type[u] = ORC_DIRICHLET;
                    continue;
                }
                node_type[u] = ORC_NEUMANN;
                if (i == 0) { *mat_at(a, col, u, u) = idx; *mat_at(a, col, u, u + 1) = -idx; }
                else if (i == ni - 1) { *mat_at(a, col, u, u) = idx; *mat_at(a, col, u, u - 1) = -idx; }
                else if (j == 0) { *mat_at(a, col, u, u) = idy; *mat_at(a, col, u, u + ni) = -idy; }
                else if (j == nj - 1) { *mat_at(a, col, u, u) = idy; *mat_at(a, col, u, u - ni) = -idy; }
                else if (k == 0) { *mat_at(a, col, u, u) = idz; *mat_at(a, col, u, u + ni * nj) = -idz; }
                else if (k == nk - 1) { *mat_at(a, col, u, u) = idz; *mat_at(a, col, u, u - ni * nj) = -idz; }
                else {
                    *mat_at(a, col, u, u - ni * nj) = idz2;
                    *mat_at(a, col, u, u - ni) = idy2;
                    *mat_at(a, col, u, u - 1) = idx2;
                    *mat_at(a, col, u, u) = -2.0 * (idx2 + idy2 + idz2);
                    *mat_at(a, col, u, u + 1) = idx2;
                    *mat_at(a, col, u, u + ni) = idy2;
                    *mat_at(a, col, u, u + ni * nj) = idz2;
                    node_type[u] = ORC_REG;
                }
            }
}

/* Matrix::multRow / one row of Matrix::operator*, ch3/ver2/PotentialSolver.cpp:24-35,67-76 */
static inline double row_dot(const double *a, const int32_t *col, int u, const double *v)
{
    double r = 0;
    for (int s = 0; s < 7; s++) {
        int c = col[7 * (size_t)u + s];
        if (c >= 0) r += a[7 * (size_t)u + s] * v[c];
        else break;
    }
    return r;
}

static void mat_vec(const double *a, const int32_t *col, int nu, const double *v, double *r)
{
    for (int u = 0; u < nu; u++) r[u] = row_dot(a, col, u, v);
}

static double vdot(const double *v1, const double *v2, int nu)
{
    double d = 0;
    for (int j = 0; j < nu; j++) d += v1[j] * v2[j];
    return d;
}

static double vnorm(const double *v, int nu)
{
    double s = 0;
    for (int j = 0; j < nu; j++) s += v[j] * v[j];
    return sqrt(s / nu);
}

/* PotentialSolver::solvePCGLinear, ch3/ver2/PotentialSolver.cpp:299-331.
 * J given by rows (a,col) with the diagonal slot index dslot[u]; M = 1/diag. */
static int pcg_linear(const double *a, const int32_t *col, const int *dslot, int nu, double *x, const double *b,
                      unsigned max_it, double tol, double *w, int64_t *iters, double *l2_out)
{
    double *g = w, *s = w + nu, *d = w + 2 * (size_t)nu, *z = w + 3 * (size_t)nu, *minv = w + 4 * (size_t)nu;
    int converged = 0;
    double l2 = 0;
    for (int u = 0; u < nu; u++) minv[u] = 1.0 / a[7 * (size_t)u + dslot[u]];
    mat_vec(a, col, nu, x, g);
    for (int u = 0; u < nu; u++) g[u] = g[u] - b[u];
    for (int u = 0; u < nu; u++) s[u] = 0 + minv[u] * g[u];
    for (int u = 0; u < nu; u++) d[u] = -1 * s[u];
    unsigned it;
    for (it = 0; it < max_it; it++) {
        mat_vec(a, col, nu, d, z);
        double alpha = vdot(g, s, nu);
        double beta = vdot(d, z, nu);
        double ab = alpha / beta;
        for (int u = 0; u < nu; u++) x[u] = x[u] + ab * d[u];
        for (int u = 0; u < nu; u++) g[u] = g[u] + ab * z[u];
        for (int u = 0; u < nu; u++) s[u] = 0 + minv[u] * g[u];
        beta = alpha;
        alpha = vdot(g, s, nu);
        ab = alpha / beta;
        for (int u = 0; u < nu; u++) d[u] = ab * d[u] - s[u];
        l2 = vnorm(g, nu);
        if (l2 < tol) { converged = 1; it++; break; }
    }
    *iters += it;
    *l2_out = l2;
    return converged;
}

/* PotentialSolver::solveGSLinear, ch3/ver2/PotentialSolver.cpp:433-461 */
static int gs_linear(const double *a, const int32_t *col, const int *dslot, int nu, double *x, const double *b,
                     unsigned max_it, double tol, double *w, int64_t *iters, double *l2_out)
{
    double *R = w;
    double L2 = 0;
    int converged = 0;
    unsigned it;
    for (it = 0; it < max_it; it++) {
        for (int u = 0; u < nu; u++) {
            double diag = a[7 * (size_t)u + dslot[u]];
            double S = row_dot(a, col, u, x) - diag * x[u];
            double phi_new = (b[u] - S) / diag;
            x[u] = x[u] + 1. * (phi_new - x[u]);
        }
        if (it % 25 == 0) {
            mat_vec(a, col, nu, x, R);
            for (int u = 0; u < nu; u++) R[u] = R[u] - b[u];
            L2 = vnorm(R, nu);
            if (L2 < tol) { converged = 1; it++; break; }
        }
    }
    *iters += it;
    *l2_out = L2;
    return converged;
}

/* PotentialSolver::solveNRPCG, ch3/ver2/PotentialSolver.cpp:225-296.  NR_MAX_IT / NR_TOL
 * (compile-time 20 / 1e-3 in the reference, :228-229) are parameters here. */
int orc_solve_nrpcg(const orc_mesh *m, const int32_t *object_id, const double *rho, double *phi,
                    double phi0, double Te0, double n0, unsigned max_it, double tol,
                    int nr_max_it, double nr_tol, orc_solve_info *info)
{
    const int nu = m->ni * m->nj * m->nk;
    double *a = (double *)malloc(sizeof(double) * 7 * (size_t)nu);
    double *ja = (double *)malloc(sizeof(double) * 7 * (size_t)nu);
    int32_t *col = (int32_t *)malloc(sizeof(int32_t) * 7 * (size_t)nu);
    int32_t *node_type = (int32_t *)malloc(sizeof(int32_t) * (size_t)nu);
    int *dslot = (int *)malloc(sizeof(int) * (size_t)nu);
    double *x = (double *)malloc(sizeof(double) * (size_t)nu);
    double *b = (double *)malloc(sizeof(double) * (size_t)nu);
    double *F = (double *)malloc(sizeof(double) * (size_t)nu);
    double *P = (double *)calloc((size_t)nu, sizeof(double));
    double *y = (double *)calloc((size_t)nu, sizeof(double));
    double *w = (double *)malloc(sizeof(double) * 5 * (size_t)nu);
    orc_solve_info li;
    memset(&li, 0, sizeof(li));

    orc_build_matrix(m, object_id, a, col, node_type);
    for (int u = 0; u < nu; u++) {
        int s;
        for (s = 0; s < 7; s++) if (col[7 * (size_t)u + s] == u) break;
        dslot[u] = s;
    }
    /* deflate: flat index already is U, ch3/ver2/PotentialSolver.cpp:124-132 */
    for (int u = 0; u < nu; u++) { x[u] = phi[u]; b[u] = rho[u]; }
    for (int u = 0; u < nu; u++) {
        if (node_type[u] == ORC_NEUMANN) b[u] = 0;
        else if (node_type[u] == ORC_DIRICHLET) b[u] = x[u];
        else b[u] = -b[u] / ORC_EPS_0;
    }
    double norm = 0;
    int converged = 0;
    for (int it = 0; it < nr_max_it; it++) {
        li.nr_iters++;
        mat_vec(a, col, nu, x, F);
        for (int u = 0; u < nu; u++) F[u] = F[u] - b[u];
        for (int n = 0; n < nu; n++)
            if (node_type[n] == ORC_REG)
                F[n] -= ORC_QE * n0 * exp((x[n] - phi0) / Te0) / ORC_EPS_0;
        for (int n = 0; n < nu; n++)
            if (node_type[n] == ORC_REG)
                P[n] = n0 * ORC_QE / (ORC_EPS_0 * Te0) * exp((x[n] - phi0) / Te0);
        /* J = A.diagSubtract(P) */
        memcpy(ja, a, sizeof(double) * 7 * (size_t)nu);
        for (int u = 0; u < nu; u++) ja[7 * (size_t)u + dslot[u]] = a[7 * (size_t)u + dslot[u]] - P[u];
        li.lin_calls++;
        if (!pcg_linear(ja, col, dslot, nu, y, F, max_it, tol, w, &li.lin_iters, &li.residual)) {
            li.gs_fallbacks++;
            gs_linear(ja, col, dslot, nu, y, F, max_it, tol, w, &li.gs_iters, &li.residual);
        }
        for (int u = 0; u < nu; u++) if (node_type[u] == ORC_DIRICHLET) y[u] = 0;
        for (int u = 0; u < nu; u++) x[u] = x[u] - y[u];
        norm = vnorm(y, nu);
        if (norm < nr_tol) { converged = 1; break; }
    }
    for (int u = 0; u < nu; u++) phi[u] = x[u];
    li.converged = converged;
    li.residual = norm;
    if (info) *info = li;
    free(a); free(ja); free(col); free(node_type); free(dslot);
    free(x); free(b); free(F); free(P); free(y); free(w);
    return converged;
}

/* PotentialSolver::computeEF, ch3/ver2/PotentialSolver.cpp:465-504 */
void orc_compute_ef(const orc_mesh *m, const double *phi, double *ef)
{
    const int ni = m->ni, nj = m->nj, nk = m->nk;
    double dx = m->dh[0], dy = m->dh[1], dz = m->dh[2];
    for (int i = 0; i < ni; i++)
        for (int j = 0; j < nj; j++)
            for (int k = 0; k < nk; k++) {
                size_t u = U(m, i, j, k);
                double p = phi[u];
                if (i == 0) ef[3 * u] = -(-3 * p + 4 * phi[U(m, i + 1, j, k)] - phi[U(m, i + 2, j, k)]) / (2 * dx);
                else if (i == ni - 1) ef[3 * u] = -(phi[U(m, i - 2, j, k)] - 4 * phi[U(m, i - 1, j, k)] + 3 * p) / (2 * dx);
                else ef[3 * u] = -(phi[U(m, i + 1, j, k)] - phi[U(m, i - 1, j, k)]) / (2 * dx);

                if (j == 0) ef[3 * u + 1] = -(-3 * p + 4 * phi[U(m, i, j + 1, k)] - phi[U(m, i, j + 2, k)]) / (2 * dy);
                else if (j == nj - 1) ef[3 * u + 1] = -(phi[U(m, i, j - 2, k)] - 4 * phi[U(m, i, j - 1, k)] + 3 * p) / (2 * dy);
                else ef[3 * u + 1] = -(phi[U(m, i, j + 1, k)] - phi[U(m, i, j - 1, k)]) / (2 * dy);

                if (k == 0) ef[3 * u + 2] = -(-3 * p + 4 * phi[U(m, i, j, k + 1)] - phi[U(m, i, j, k + 2)]) / (2 * dz);
                else if (k == nk - 1) ef[3 * u + 2] = -(phi[U(m, i, j, k - 2)] - 4 * phi[U(m, i, j, k - 1)] + 3 * p) / (2 * dz);
                else ef[3 * u + 2] = -(phi[U(m, i, j, k + 1)] - phi[U(m, i, j, k - 1)]) / (2 * dz);
            }
}

/* ======================================================================================================
 * ch4: Species::advance(neutrals, spherium) with surface interactions (ch4/Species.cpp:8-91),
 * World::lineSphereIntersect (ch4/World.cpp:160-183), World::sphereDiffuseVector (:185-199),
 * Species::sampleReflectedVelocity (ch4/Species.cpp:93-100), Species::sampleVth (:149-159).
 *
 * Particle::dt (ch4/Species.h:15) is carried in a separate array pdt[] parallel to the SoA particle arrays.
 * Random numbers: mode 0 = the reference's sequential mt19937 stream (pinned against oracle/_ref/ref_ch4_surface),
 * mode 1 = Philox counters shared with the CUDA engine: particle i, bounce/emission e use the six blocks
 * idx = (i << 20) + 8*e + j (j = 0..5, e >= 1 for emissions); the two emission-count uniforms of ion i are block (i << 20).
 * ====================================================================================================== */
typedef struct {
    const orc_surface_rng *cfg;
    uint64_t base;      /* Philox: first block of the current group */
    int k;              /* Philox: uniforms already taken from the group */
} rng_cursor;

static double rng_draw(rng_cursor *r)
{
    if (r->cfg->mode == 0) return orc_mt_uniform(r->cfg->mt);
    double u2[2];
    orc_philox_uniform2(r->cfg->seed, r->cfg->stream, r->cfg->step, r->base + (uint64_t)(r->k >> 1), u2);
    return u2[(r->k++) & 1];
}
static void rng_seek(rng_cursor *r, uint64_t base, int k) { r->base = base; r->k = k; }

double orc_line_sphere_intersect(const orc_mesh *m, const double x1[3], const double x2[3])
{
    double B[3], A[3];
    for (int c = 0; c < 3; c++) { B[c] = x2[c] - x1[c]; A[c] = x1[c] - m->sphere_c[c]; }
    double a = 0, b = 0, cc = 0;
    for (int c = 0; c < 3; c++) a += B[c] * B[c];
    for (int c = 0; c < 3; c++) b += A[c] * B[c];
    b = 2 * b;
    for (int c = 0; c < 3; c++) cc += A[c] * A[c];
    cc = cc - m->sphere_r2;
    double det = b * b - 4 * a * cc;
    if (det < 0) return 0.5;
    double tp = (-b + sqrt(det)) / (2 * a);
    if (tp < 0 || tp > 1.0) {
        tp = (-b - sqrt(det)) / (2 * a);
        if (tp < 0 || tp > 1.0) tp = 0.5;
    }
    return tp;
}

static void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/* u[0..8]: the nine sampleVth uniforms, u[9]: sin_theta, u[10]: psi/(2 pi) */
void orc_reflected_velocity(const orc_mesh *m, const double pos[3], double v_mag1, double mass, const double u[11], double vel[3])
{
    const double K = 1.380648e-23, PI = 3.141592653;
    /* sampleVth(1000) */
    double v_th = sqrt(2 * K * 1000 / mass);
    double v1 = v_th * (u[0] + u[1] + u[2] - 1.5);
    double v2 = v_th * (u[3] + u[4] + u[5] - 1.5);
    double v3 = v_th * (u[6] + u[7] + u[8] - 1.5);
    double vth = 3 / sqrt(2 + 2 + 2) * sqrt(v1 * v1 + v2 * v2 + v3 * v3);
    const double a_th = 1;
    double v_mag2 = v_mag1 + a_th * (vth - v_mag1);
    /* sphereDiffuseVector(pos) */
    double sin_theta = u[9];
    double cos_theta = sqrt(1 - sin_theta * sin_theta);
    double psi = 2 * PI * u[10];
    double n[3], t1[3], t2[3], d[3];
    for (int c = 0; c < 3; c++) d[c] = pos[c] - m->sphere_c[c];
    double s = 0;
    for (int c = 0; c < 3; c++) s += d[c] * d[c];
    double mg = sqrt(s);
    for (int c = 0; c < 3; c++) n[c] = d[c] / mg;
    const double ex[3] = { 1, 0, 0 }, ey[3] = { 0, 1, 0 };
    double dn = 0;
    for (int c = 0; c < 3; c++) dn += n[c] * ex[c];
    if (dn != 0) cross3(n, ex, t1); else cross3(n, ey, t1);
    cross3(n, t1, t2);
    double cp = sin_theta * cos(psi), sp = sin_theta * sin(psi);
    for (int c = 0; c < 3; c++) {
        double r = t1[c] * cp + t2[c] * sp + n[c] * cos_theta;       /* s*vec evaluates a(i)*s (ch4/Field.h:59-60) */
        vel[c] = r * v_mag2;
    }
}

static void draw_reflected(const orc_mesh *m, rng_cursor *r, const double pos[3], double v_mag1, double mass, double vel[3])
{
    double u[11];
    for (int q = 0; q < 11; q++) u[q] = rng_draw(r);
    orc_reflected_velocity(m, pos, v_mag1, mass, u, vel);
}

static int add_with_dt(const orc_mesh *m, const double *ef, orc_surface_target *t, const double pos[3], const double vel[3], double dt)
{
    int64_t before = t->p->np;
    int ok = orc_add_particle(m, ef, t->p, pos, vel, t->mpw0, t->charge, t->mass, dt);
    if (ok) t->pdt[before] = dt;          /* addParticle(pos,vel) -> dt = world.getDt() (ch4/Species.h:65) */
    return ok;
}

int64_t orc_advance_surface(const orc_mesh *m, const double *ef, orc_particles *p, double *pdt, double charge, double mass,
                            double mpw0, double dt, orc_surface_target *neutrals, orc_surface_target *sput,
                            const orc_surface_rng *rng, int64_t emitted[2])
{
    rng_cursor rc = { rng, 0, 0 };
    emitted[0] = emitted[1] = 0;
    for (int64_t q = 0; q < p->np; q++) {
        pdt[q] += dt;
        double pos[3] = { p->x[q], p->y[q], p->z[q] }, vel[3] = { p->vx[q], p->vy[q], p->vz[q] };
        double lc[3], e[3];
        orc_xtol(m, pos, lc);
        orc_gather3(m, ef, lc, e);
        const double s = pdt[q] * charge / mass;
        for (int c = 0; c < 3; c++) vel[c] += e[c] * s;
        int64_t bounce = 0;
        while (pdt[q] > 0 && p->mpw[q] > 0) {
            double pos_old[3] = { pos[0], pos[1], pos[2] };
            for (int c = 0; c < 3; c++) pos[c] += vel[c] * pdt[q];
            if (!orc_in_bounds(m, pos)) {
                p->mpw[q] = 0;
            } else if (orc_in_sphere(m, pos)) {
                double tp = orc_line_sphere_intersect(m, pos_old, pos);
                double dt_rem = (1 - tp) * pdt[q];
                pdt[q] -= dt_rem;
                const double f = 0.999 * tp;
                for (int c = 0; c < 3; c++) pos[c] = pos_old[c] + (pos[c] - pos_old[c]) * f;
                double v2 = 0;
                for (int c = 0; c < 3; c++) v2 += vel[c] * vel[c];
                double v_mag1 = sqrt(v2);
                if (charge == 0) {
                    rng_seek(&rc, ((uint64_t)q << 20) + 8 * (uint64_t)bounce, 0);
                    draw_reflected(m, &rc, pos, v_mag1, mass, vel);
                    bounce++;
                } else {
                    double mpw_ratio = mpw0 / neutrals->mpw0;
                    p->mpw[q] = 0;
                    rng_seek(&rc, (uint64_t)q << 20, 0);
                    int mp_create = (int)(mpw_ratio + rng_draw(&rc));
                    for (int i = 0; i < mp_create; i++) {
                        double v[3];
                        rng_seek(&rc, ((uint64_t)q << 20) + 8 * (uint64_t)(1 + i), 0);
                        draw_reflected(m, &rc, pos, v_mag1, mass, v);
                        emitted[0] += add_with_dt(m, ef, neutrals, pos, v, dt);
                    }
                    double sput_yield = (v_mag1 > 5000) ? 0.1 : 0;
                    double sput_mpw_ratio = sput_yield * mpw0 / sput->mpw0;
                    rng_seek(&rc, (uint64_t)q << 20, 1);
                    int sput_mp_create = (int)(sput_mpw_ratio + rng_draw(&rc));
                    for (int i = 0; i < sput_mp_create; i++) {
                        double v[3];
                        rng_seek(&rc, ((uint64_t)q << 20) + 8 * (uint64_t)(1 + mp_create + i), 0);
                        draw_reflected(m, &rc, pos, v_mag1, mass, v);
                        emitted[1] += add_with_dt(m, ef, sput, pos, v, dt);
                    }
                }
                continue;
            }
            pdt[q] = 0;
        }
        p->x[q] = pos[0]; p->y[q] = pos[1]; p->z[q] = pos[2];
        p->vx[q] = vel[0]; p->vy[q] = vel[1]; p->vz[q] = vel[2];
    }
    /* removal, ch4/Species.cpp:77-88 (the whole Particle incl. dt moves) */
    int64_t np = p->np;
    for (int64_t q = 0; q < np; q++) {
        if (p->mpw[q] > 0) continue;
        copy_particle(p, q, np - 1);
        pdt[q] = pdt[np - 1];
        np--;
        q--;
    }
    p->np = np;
    return np;
}

/* ======================================================================================================
 * ch4: DSMC_MEX::apply / collide / evalSigma (ch4/Collisions.cpp:84-182, ch4/Collisions.h:59-83) -- Bird's NTC scheme
 * with VHS cross-sections on the particles of one species, cell by cell -- and Species::computeMPC (ch4/Species.cpp:228-235).
 * Cell of a particle: World::XtoC (ch4/World.h:88-98), c = k*(nj-1)*(ni-1) + j*(ni-1) + i.  The per-cell particle lists keep
 * the particle order (push_back while looping over the particles).
 * Random numbers: mode 0 = the reference's sequential mt19937 stream; mode 1 = Philox, draw number q of cell c is element
 * q&1 of block (c << 24) + (q >> 1)  (shared with the CUDA engine).
 * ====================================================================================================== */
static int64_t cell_index(const orc_mesh *m, double x, double y, double z)
{
    double pos[3] = { x, y, z }, lc[3], d;
    int i, j, k;
    orc_xtol(m, pos, lc);
    cell_of(lc[0], m->ni, &i, &d);
    cell_of(lc[1], m->nj, &j, &d);
    cell_of(lc[2], m->nk, &k, &d);
    return ((int64_t)k * (m->nj - 1) + j) * (m->ni - 1) + i;
}

void orc_compute_mpc(const orc_mesh *m, const orc_particles *p, double *mpc)
{
    int64_t nc = (int64_t)(m->ni - 1) * (m->nj - 1) * (m->nk - 1);
    for (int64_t c = 0; c < nc; c++) mpc[c] = 0;
    for (int64_t q = 0; q < p->np; q++) mpc[cell_index(m, p->x[q], p->y[q], p->z[q])] += 1;
}

double orc_vhs_sigma(double mass, double g_rel)
{
    const double K = 1.380648e-23, PI = 3.141592653;
    double mr = mass * mass / (mass + mass);
    double c0 = 4.07e-10, c1 = 0.77, c2 = 2 * K * 273.15 / mr, c3 = tgamma(2.5 - c1);
    return PI * c0 * c0 * pow(c2 / (g_rel * g_rel), c1 - 0.5) / c3;
}

int64_t orc_dsmc_mex(const orc_mesh *m, orc_particles *p, double mass, double mpw0, double dt, double *sigma_cr_max_io,
                     const orc_surface_rng *rng)
{
    const double PI = 3.141592653;
    int64_t nc = (int64_t)(m->ni - 1) * (m->nj - 1) * (m->nk - 1);
    int64_t *start = (int64_t *)calloc((size_t)nc + 1, sizeof(int64_t));
    int64_t *list = (int64_t *)malloc((size_t)(p->np > 0 ? p->np : 1) * sizeof(int64_t));
    int64_t *cell = (int64_t *)malloc((size_t)(p->np > 0 ? p->np : 1) * sizeof(int64_t));
    for (int64_t q = 0; q < p->np; q++) { cell[q] = cell_index(m, p->x[q], p->y[q], p->z[q]); start[cell[q] + 1]++; }
    for (int64_t c = 0; c < nc; c++) start[c + 1] += start[c];
    {
        int64_t *cur = (int64_t *)malloc((size_t)nc * sizeof(int64_t));
        memcpy(cur, start, (size_t)nc * sizeof(int64_t));
        for (int64_t q = 0; q < p->np; q++) list[cur[cell[q]]++] = q;
        free(cur);
    }
    double sigma_cr_max = *sigma_cr_max_io, sigma_cr_max_temp = 0;
    double dV = m->dh[0] * m->dh[1] * m->dh[2];
    double Fn = mpw0;
    int64_t num_cols = 0;
    rng_cursor rc = { rng, 0, 0 };
    for (int64_t c = 0; c < nc; c++) {
        const int64_t *parts = list + start[c];
        int np = (int)(start[c + 1] - start[c]);
        if (np < 2) continue;
        rng_seek(&rc, (uint64_t)c << 24, 0);
        double ng_f = 0.5 * np * np * Fn * sigma_cr_max * dt / dV;
        int ng = (int)(ng_f + 0.5);
        for (int g = 0; g < ng; g++) {
            int p1 = (int)(rng_draw(&rc) * np), p2;
            do { p2 = (int)(rng_draw(&rc) * np); } while (p2 == p1);
            int64_t a = parts[p1], b = parts[p2];
            double v1[3] = { p->vx[a], p->vy[a], p->vz[a] }, v2[3] = { p->vx[b], p->vy[b], p->vz[b] };
            double cr_vec[3], s = 0;
            for (int d = 0; d < 3; d++) { cr_vec[d] = v1[d] - v2[d]; s += cr_vec[d] * cr_vec[d]; }
            double cr = sqrt(s);
            double sigma = orc_vhs_sigma(mass, cr);
            double sigma_cr = sigma * cr;
            if (sigma_cr > sigma_cr_max_temp) sigma_cr_max_temp = sigma_cr;
            double P = sigma_cr / sigma_cr_max;
            if (P > rng_draw(&rc)) {
                num_cols++;
                /* DSMC_MEX::collide(vel1, vel2, mass, mass), ch4/Collisions.cpp:84-106 */
                double cm[3], crr[3];
                for (int d = 0; d < 3; d++) cm[d] = (v1[d] * mass + v2[d] * mass) / (mass + mass);
                double cr_mag = cr;                       /* mag(vel1 - vel2): the same expression again */
                double cos_chi = 2 * rng_draw(&rc) - 1;
                double sin_chi = sqrt(1 - cos_chi * cos_chi);
                double eps = 2 * PI * rng_draw(&rc);
                crr[0] = cr_mag * cos_chi;
                crr[1] = cr_mag * sin_chi * cos(eps);
                crr[2] = cr_mag * sin_chi * sin(eps);
                double f2 = mass / (mass + mass);
                p->vx[a] = cm[0] + crr[0] * f2; p->vy[a] = cm[1] + crr[1] * f2; p->vz[a] = cm[2] + crr[2] * f2;
                p->vx[b] = cm[0] - crr[0] * f2; p->vy[b] = cm[1] - crr[1] * f2; p->vz[b] = cm[2] - crr[2] * f2;
            }
        }
    }
    free(start); free(list); free(cell);
    if (num_cols) *sigma_cr_max_io = sigma_cr_max_temp;
    return num_cols;
}

/* ======================================================================================================
 * ch4: MCC_CEX::apply (ch4/Collisions.cpp:43-82): Monte Carlo collisions of `source` particles with the mesh-averaged
 * target gas (density den, stream velocity vel[3*u+c], temperature T on the nodes).  A colliding particle's velocity is set to
 * zero (the reference has the charge-exchange assignment commented out, :78-79) after Species::sampleIsotropicVel consumed
 * eleven random numbers.  mode 0: sequential mt19937; mode 1: Philox, particle q compares against element 0 of block q.
 * ====================================================================================================== */
double orc_gather1(const orc_mesh *m, const double *f, const double lc[3])
{
    int i, j, k; double di, dj, dk;
    cell_of(lc[0], m->ni, &i, &di);
    cell_of(lc[1], m->nj, &j, &dj);
    cell_of(lc[2], m->nk, &k, &dk);
    const size_t n[8] = { U(m, i, j, k), U(m, i + 1, j, k), U(m, i + 1, j + 1, k), U(m, i, j + 1, k),
                          U(m, i, j, k + 1), U(m, i + 1, j, k + 1), U(m, i + 1, j + 1, k + 1), U(m, i, j + 1, k + 1) };
    const double wi[8] = { 1 - di, di, di, 1 - di, 1 - di, di, di, 1 - di };
    const double wj[8] = { 1 - dj, 1 - dj, dj, dj, 1 - dj, 1 - dj, dj, dj };
    const double wk[8] = { 1 - dk, 1 - dk, 1 - dk, 1 - dk, dk, dk, dk, dk };
    double val = f[n[0]] * wi[0] * wj[0] * wk[0];
    for (int t = 1; t < 8; t++) val = val + f[n[t]] * wi[t] * wj[t] * wk[t];
    return val;
}

int64_t orc_mcc_cex(const orc_mesh *m, orc_particles *p, const double *target_den, const double *target_vel,
                    const double *target_T, double target_mass, double dt, const orc_surface_rng *rng)
{
    int64_t cols = 0;
    rng_cursor rc = { rng, 0, 0 };
    for (int64_t q = 0; q < p->np; q++) {
        double pos[3] = { p->x[q], p->y[q], p->z[q] }, lc[3], vt[3];
        orc_xtol(m, pos, lc);
        orc_gather3(m, target_vel, lc, vt);
        double nn = orc_gather1(m, target_den, lc);
        double r0 = p->vx[q] - vt[0], r1 = p->vy[q] - vt[1], r2 = p->vz[q] - vt[2];
        double s = 0;
        s += r0 * r0; s += r1 * r1; s += r2 * r2;
        double v_rel_mag = sqrt(s);
        double sigma = 1e-16;
        double P = 1 - exp(-nn * sigma * v_rel_mag * dt);
        rng_seek(&rc, (uint64_t)q, 0);
        if (P >= rng_draw(&rc)) {
            if (rng->mode == 0) {
                double T_target = orc_gather1(m, target_T, lc), u[11], v[3];
                for (int t = 0; t < 11; t++) u[t] = rng_draw(&rc);
                orc_isotropic_vel(T_target, target_mass, u, v);       /* sampled and dropped, as in the reference */
            }
            p->vx[q] = 0; p->vy[q] = 0; p->vz[q] = 0;
            cols++;
        }
    }
    return cols;
}
