// espic_surface.cuh -- ch4 surface interactions: Species::advance(neutrals, spherium) (ch4/Species.cpp:8-100).
// Included at the end of espic_particles.cu (shares its device helpers, scan and removal code).
//
//   * every particle carries its own remaining time Particle::dt (ch4/Species.h:15).  After an advance it is 0 for every
//     survivor, and addParticle(pos,vel) gives new particles dt = world dt (ch4/Species.h:65) -- so instead of an eighth
//     particle array the engine keeps Species::n_settled: particles [0, n_settled) went through the last advance (dt = 0),
//     particles [n_settled, np) were added since (dt = world dt).  "part.dt += world.getDt()" (Species.cpp:14) is then
//     0 + dt or dt + dt.
//   * sub-step loop (Species.cpp:27-74): move; outside the box -> dead; inside the sphere -> World::lineSphereIntersect
//     (ch4/World.cpp:160-183), step back to 0.999 of the way to the surface, part.dt -= (1-tp)*part.dt; a NEUTRAL leaves
//     again with sampleReflectedVelocity (Species.cpp:93-100: Birdsall speed of a 1000 K wall, full accommodation, along
//     World::sphereDiffuseVector, ch4/World.cpp:185-199) and keeps moving; an ION dies and emits (int)(mpw0/neutrals.mpw0 + R)
//     neutrals plus (int)(yield*mpw0/spherium.mpw0 + R) sputtered particles (yield 0.1 above 5 km/s) through addParticle.
//   * random numbers are Philox4x32-10 counters, so the result does not depend on thread order: bounce b of particle i uses
//     blocks (i << 20) + 8*b + j (j = 0..5: nine Birdsall uniforms, sin_theta, psi); ion i draws its two emission counts from
//     block (i << 20) and emission e from blocks (i << 20) + 8*(1+e) + j.  The oracle (orc_advance_surface, Philox mode) uses
//     the same counters.
//   * emitted particles are appended in the reference's order (ascending ion index, neutrals before sputtered material of the
//     same ion) with an exclusive scan of the per-ion counts; removal is the reference's swap-with-last order as in espic_push.

struct SurfPar {
    double charge, mass, dt;
    long long n_settled;
    double v_th;                  // sqrt(2*K*1000/mass) of the ADVANCING species (sampleVth is its member function)
    uint64_t seed; uint32_t stream, step;
};

#define SURF_MAX_SUBSTEPS 4096    // the reference loop has no bound; a particle still bouncing after this many sub-steps is killed

// World::lineSphereIntersect (ch4/World.cpp:160-183)
__device__ __forceinline__ double line_sphere_intersect(const MeshC &m, const double x1[3], const double x2[3])
{
    double B[3], A[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { B[c] = x2[c] - x1[c]; A[c] = x1[c] - m.sc[c]; }
    const double a = (B[0] * B[0] + B[1] * B[1]) + B[2] * B[2];
    const double b = 2 * ((A[0] * B[0] + A[1] * B[1]) + A[2] * B[2]);
    const double cc = ((A[0] * A[0] + A[1] * A[1]) + A[2] * A[2]) - m.sr2;
    const double det = b * b - 4 * a * cc;
    if (det < 0) return 0.5;
    double tp = (-b + sqrt(det)) / (2 * a);
    if (tp < 0 || tp > 1.0) {
        tp = (-b - sqrt(det)) / (2 * a);
        if (tp < 0 || tp > 1.0) tp = 0.5;
    }
    return tp;
}

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// Species::sampleReflectedVelocity (ch4/Species.cpp:93-100) with sampleVth(1000) (:149-159) and sphereDiffuseVector
// (ch4/World.cpp:185-199); the eleven uniforms come from six Philox blocks starting at `block`
__device__ __forceinline__ void reflected_velocity(const MeshC &m, const SurfPar &sp, uint64_t block, const double pos[3],
                                                   double v_mag1, double vel[3])
{
    double u[12];
#pragma unroll
    for (int j = 0; j < 6; j++) philox_uniform2(sp.seed, sp.stream, sp.step, block + j, u[2 * j], u[2 * j + 1]);
    const double v1 = sp.v_th * (u[0] + u[1] + u[2] - 1.5);
    const double v2 = sp.v_th * (u[3] + u[4] + u[5] - 1.5);
    const double v3 = sp.v_th * (u[6] + u[7] + u[8] - 1.5);
    const double vth = 3 / sqrt(2.0 + 2 + 2) * sqrt(v1 * v1 + v2 * v2 + v3 * v3);
    const double a_th = 1;
    const double v_mag2 = v_mag1 + a_th * (vth - v_mag1);
    const double sin_theta = u[9];
    const double cos_theta = sqrt(1 - sin_theta * sin_theta);
    const double psi = 2 * 3.141592653 * u[10];          // Const::PI as the reference spells it
    double d[3], n[3], t1[3], t2[3];
#pragma unroll
    for (int c = 0; c < 3; c++) d[c] = pos[c] - m.sc[c];
    const double mg = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
#pragma unroll
    for (int c = 0; c < 3; c++) n[c] = d[c] / mg;
    const double ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0};
    const double dn = (n[0] * ex[0] + n[1] * ex[1]) + n[2] * ex[2];
    if (dn != 0) cross3(n, ex, t1); else cross3(n, ey, t1);
    cross3(n, t1, t2);
    const double cp = sin_theta * cos(psi), sn = sin_theta * sin(psi);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double r = (t1[c] * cp + t2[c] * sn) + n[c] * cos_theta;
        vel[c] = r * v_mag2;
    }
}

template <bool CHARGED>
__global__ void __launch_bounds__(256) k_push_surface(MeshC m, const double *__restrict__ ef4,
                                                      double *__restrict__ px, double *__restrict__ py, double *__restrict__ pz,
                                                      double *__restrict__ pvx, double *__restrict__ pvy, double *__restrict__ pvz,
                                                      double *__restrict__ pmpw, long long n, SurfPar sp,
                                                      uint32_t *__restrict__ dead_words, uint32_t *__restrict__ hit_words)
{
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    const bool own = i < n;
    bool dead = false, hit = false;
    if (own) {
        double pos[3] = {px[i], py[i], pz[i]}, vel[3] = {pvx[i], pvy[i], pvz[i]};
        double mpw = pmpw[i];
        const double mpw_in = mpw;
        double pdt = (i >= sp.n_settled) ? sp.dt : 0.0;
        pdt += sp.dt;                                        // part.dt += world.getDt()
        int ci, cj, ck; double di, dj, dk;
        cell3(m, pos[0], pos[1], pos[2], ci, cj, ck, di, dj, dk);
        if (ci < 0) ci = 0;
        if (cj < 0) cj = 0;
        if (ck < 0) ck = 0;
        double e[3];
        gather_ef(m, ef4, ci, cj, ck, di, dj, dk, e);
        const double s = pdt * sp.charge / sp.mass;          // part.vel += ef_part*(part.dt*charge/mass)
#pragma unroll
        for (int c = 0; c < 3; c++) vel[c] = vel[c] + e[c] * s;
        unsigned bounce = 0;
        int guard = 0;
        while (pdt > 0 && mpw > 0) {
            if (++guard > SURF_MAX_SUBSTEPS) { mpw = 0; break; }
            double pos_old[3] = {pos[0], pos[1], pos[2]};
#pragma unroll
            for (int c = 0; c < 3; c++) pos[c] = pos[c] + vel[c] * pdt;
            if (!in_bounds(m, pos[0], pos[1], pos[2])) {
                mpw = 0;
            } else if (in_sphere(m, pos[0], pos[1], pos[2])) {
                const double tp = line_sphere_intersect(m, pos_old, pos);
                const double dt_rem = (1 - tp) * pdt;
                pdt -= dt_rem;
                const double f = 0.999 * tp;
#pragma unroll
                for (int c = 0; c < 3; c++) pos[c] = pos_old[c] + (pos[c] - pos_old[c]) * f;
                if (!CHARGED) {
                    const double v_mag1 = sqrt((vel[0] * vel[0] + vel[1] * vel[1]) + vel[2] * vel[2]);
                    reflected_velocity(m, sp, ((uint64_t)i << 20) + 8ull * bounce, pos, v_mag1, vel);
                    bounce++;
                } else {
                    mpw = 0;            // the ion dies here; k_emit_* read the impact point and the impact velocity from its slot
                    hit = true;
                }
                continue;
            }
            pdt = 0;
        }
        px[i] = pos[0]; py[i] = pos[1]; pz[i] = pos[2];
        pvx[i] = vel[0]; pvy[i] = vel[1]; pvz[i] = vel[2];
        if (mpw != mpw_in) pmpw[i] = mpw;
        dead = !(mpw > 0);
    }
    const uint32_t bd = __ballot_sync(0xffffffffu, dead);
    const uint32_t bh = CHARGED ? __ballot_sync(0xffffffffu, hit) : 0u;
    if ((threadIdx.x & 31) == 0 && own) {
        dead_words[i >> 5] = bd;
        if (CHARGED) hit_words[i >> 5] = bh;
    }
}

struct EmitPar {
    double mpw_ratio;             // this->mpw0/neutrals.mpw0
    double mpw0, sput_mpw0;       // sput_yield*this->mpw0/spherium.mpw0 is evaluated per impact
    int mask;                     // bit 0: neutrals, bit 1: sputtered material go to the target of this pass
};

__device__ __forceinline__ void emit_counts(const MeshC &m, const SurfPar &sp, const EmitPar &ep, long long i, double vx, double vy,
                                            double vz, int &mp_create, int &sput_create, double &v_mag1)
{
    double u0, u1;
    philox_uniform2(sp.seed, sp.stream, sp.step, (uint64_t)i << 20, u0, u1);
    v_mag1 = sqrt((vx * vx + vy * vy) + vz * vz);
    mp_create = (int)(ep.mpw_ratio + u0);
    const double sput_yield = (v_mag1 > 5000) ? 0.1 : 0;
    const double sput_mpw_ratio = sput_yield * ep.mpw0 / ep.sput_mpw0;
    sput_create = (int)(sput_mpw_ratio + u1);
}

__global__ void __launch_bounds__(256) k_emit_count(MeshC m, SurfPar sp, EmitPar ep, long long n, const uint32_t *__restrict__ hit_words,
                                                    const double *__restrict__ px, const double *__restrict__ py,
                                                    const double *__restrict__ pz, const double *__restrict__ pvx,
                                                    const double *__restrict__ pvy, const double *__restrict__ pvz,
                                                    uint32_t *__restrict__ cnt)
{
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= n) return;
    uint32_t k = 0;
    if ((hit_words[i >> 5] >> (i & 31)) & 1u) {
        int a, b; double vm;
        emit_counts(m, sp, ep, i, pvx[i], pvy[i], pvz[i], a, b, vm);
        if (in_bounds(m, px[i], py[i], pz[i])) k = (uint32_t)(((ep.mask & 1) ? a : 0) + ((ep.mask & 2) ? b : 0));   // addParticle's own test
    }
    cnt[i] = k;
}

__global__ void __launch_bounds__(256) k_emit_write(MeshC m, SurfPar sp, EmitPar ep, long long n, const double *__restrict__ ef4,
                                                    const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ pre,
                                                    const uint32_t *__restrict__ coff,
                                                    const double *__restrict__ px, const double *__restrict__ py,
                                                    const double *__restrict__ pz, const double *__restrict__ pvx,
                                                    const double *__restrict__ pvy, const double *__restrict__ pvz,
                                                    long long base, double qm, double hdt, double mpw_new,
                                                    double *t0, double *t1, double *t2, double *t3, double *t4, double *t5, double *t6)
{
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= n || cnt[i] == 0) return;
    int a, b; double v_mag1;
    emit_counts(m, sp, ep, i, pvx[i], pvy[i], pvz[i], a, b, v_mag1);
    const double pos[3] = {px[i], py[i], pz[i]};
    int ci, cj, ck; double di, dj, dk;
    cell3(m, pos[0], pos[1], pos[2], ci, cj, ck, di, dj, dk);
    double e[3];
    gather_ef(m, ef4, ci, cj, ck, di, dj, dk, e);
    long long dst = base + (long long)scan_at(pre, coff, i);
    const int e_lo = (ep.mask & 1) ? 0 : a, e_hi = (ep.mask & 2) ? a + b : a;
    for (int em = e_lo; em < e_hi; em++) {
        double vel[3];
        reflected_velocity(m, sp, ((uint64_t)i << 20) + 8ull * (uint64_t)(1 + em), pos, v_mag1, vel);
        t0[dst] = pos[0]; t1[dst] = pos[1]; t2[dst] = pos[2];
        t3[dst] = vel[0] - e[0] * qm * hdt;          // addParticle: vel -= charge/mass*ef_part*(0.5*dt)
        t4[dst] = vel[1] - e[1] * qm * hdt;
        t5[dst] = vel[2] - e[2] * qm * hdt;
        t6[dst] = mpw_new;
        dst++;
    }
}

static int emit_pass(espic_ctx *c, Species &s, long long n, const SurfPar &sp, EmitPar ep, int target, double dt, long long *emitted)
{
    Species &t = c->sp[target];
    int r;
    if ((r = ensure_buf(&c->cell_cnt, &c->cell_cap, n, c->stream))) return r;
    k_emit_count<<<nblk(n, 256), 256, 0, c->stream>>>(c->m, sp, ep, n, c->hit_words, s.p[0], s.p[1], s.p[2], s.p[3], s.p[4], s.p[5],
                                                      c->cell_cnt);
    LAUNCH_CHECK(c);
    if ((r = espic_scan_u32(c, c->cell_cnt, n, c->dscal + 3))) return r;
    unsigned long long *h = (unsigned long long *)c->hpin;
    CK(cudaMemcpyAsync(h + 3, c->dscal + 3, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const long long total = (long long)h[3];
    *emitted = total;
    if (total == 0) return 0;
    if ((r = espic_species_reserve(c, target, t.np + total))) return r;
    k_emit_write<<<nblk(n, 256), 256, 0, c->stream>>>(c->m, sp, ep, n, c->ef4, c->cell_cnt, c->scan_pre, c->scan_coff,
                                                      s.p[0], s.p[1], s.p[2], s.p[3], s.p[4], s.p[5], t.np,
                                                      t.charge / t.mass, 0.5 * dt, t.mpw0,
                                                      t.p[0], t.p[1], t.p[2], t.p[3], t.p[4], t.p[5], t.p[6]);
    LAUNCH_CHECK(c);
    t.np += total;
    if (t.mpw0 > t.mpw_max) t.mpw_max = t.mpw0;
    t.acc_fresh = false;
    t.pushes_since_sort = 1 << 20;
    return 0;
}

extern "C" int espic_push_surface(espic_ctx *c, int sp, double dt, int neutrals_sp, int sput_sp,
                                  uint64_t seed, uint32_t stream, uint32_t step, long long emitted[2])
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    Species &s = c->sp[sp];
    if (emitted) emitted[0] = emitted[1] = 0;
    MIG_GUARD(c, s, "espic_push_surface");
    s.diag_valid = false;
    const bool charged = s.charge != 0;
    if (charged) {
        SP_CHECK(c, neutrals_sp);
        SP_CHECK(c, sput_sp);
        if (neutrals_sp == sp || sput_sp == sp) {
            espic_set_error("espic_push_surface: an ion species cannot emit into itself (species %d)", sp);
            return -1;
        }
        MIG_GUARD(c, c->sp[neutrals_sp], "espic_push_surface (emission target)");
        MIG_GUARD(c, c->sp[sput_sp], "espic_push_surface (emission target)");
        if (!(c->sp[neutrals_sp].mpw0 > 0) || !(c->sp[sput_sp].mpw0 > 0) || s.mpw0 / c->sp[neutrals_sp].mpw0 > 1e5) {
            espic_set_error("espic_push_surface: bad macroparticle weight ratio between species %d and its emission targets", sp);
            return -1;
        }
    }
    const long long n = s.np;
    s.acc_fresh = false;
    s.substep = true;
    if (n == 0) { s.n_settled = 0; return 0; }
    const long long nw = (n + 31) / 32;
    int r;
    if ((r = ensure_buf(&s.kill_words, &s.kill_cap, nw, c->stream))) return r;
    if (charged && (r = ensure_buf(&c->hit_words, &c->hit_words_cap, nw, c->stream))) return r;
    SurfPar p;
    p.charge = s.charge; p.mass = s.mass; p.dt = dt;
    p.n_settled = std::min(s.n_settled, n);
    p.v_th = sqrt(2 * 1.380648e-23 * 1000 / s.mass);       // sampleVth(1000): Const::K (ch4/World.h:18), T_sphere = 1000 K
    p.seed = seed; p.stream = stream; p.step = step;
    if (charged)
        k_push_surface<true><<<nblk(n, 256), 256, 0, c->stream>>>(c->m, c->ef4, s.p[0], s.p[1], s.p[2], s.p[3], s.p[4], s.p[5], s.p[6],
                                                                  n, p, s.kill_words, c->hit_words);
    else
        k_push_surface<false><<<nblk(n, 256), 256, 0, c->stream>>>(c->m, c->ef4, s.p[0], s.p[1], s.p[2], s.p[3], s.p[4], s.p[5], s.p[6],
                                                                   n, p, s.kill_words, nullptr);
    LAUNCH_CHECK(c);
    if (s.pushes_since_sort < (1 << 20)) s.pushes_since_sort++;
    if (charged) {
        EmitPar ep;
        ep.mpw_ratio = s.mpw0 / c->sp[neutrals_sp].mpw0;
        ep.mpw0 = s.mpw0;
        ep.sput_mpw0 = c->sp[sput_sp].mpw0;
        long long em = 0;
        if (neutrals_sp == sput_sp) {
            ep.mask = 3;
            if ((r = emit_pass(c, s, n, p, ep, neutrals_sp, dt, &em))) return r;
            if (emitted) emitted[0] = em;          // both kinds, in the reference's interleaved order
        } else {
            ep.mask = 1;
            if ((r = emit_pass(c, s, n, p, ep, neutrals_sp, dt, &em))) return r;
            if (emitted) emitted[0] = em;
            ep.mask = 2;
            if ((r = emit_pass(c, s, n, p, ep, sput_sp, dt, &em))) return r;
            if (emitted) emitted[1] = em;
        }
    }
    if ((r = compact_dead(c, s, n))) return r;
    s.n_settled = s.np;
    return 0;
}
