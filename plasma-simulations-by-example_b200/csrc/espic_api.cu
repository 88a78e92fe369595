// espic_api.cu -- context lifetime, geometry, field and particle transfer of the C ABI (include/espic.h).
#include "espic_internal.cuh"
#include <stdarg.h>
#include <algorithm>

static thread_local char g_err[512] = "";

void espic_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int espic_cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    espic_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return -(int)e - 1000;
}

extern "C" const char *espic_last_error(void) { return g_err; }

int espic_ensure(void **ptr, long long *cap, long long need, size_t elem, cudaStream_t s)
{
    if (need <= *cap && *ptr) return 0;
    // grow geometrically: sizes such as "number of particles removed this step" drift from call to call, and a
    // cudaFree/cudaMalloc pair synchronises the whole device (measured: up to 250 ms with tens of GB resident)
    long long ncap = std::max<long long>(2 * need, 1024);
    if (*ptr) { CK(cudaStreamSynchronize(s)); CK(cudaFree(*ptr)); *ptr = nullptr; }
    CK(cudaMalloc(ptr, (size_t)ncap * elem));
    *cap = ncap;
    return 0;
}

// ---- geometry kernels ---------------------------------------------------------------------------

// World::computeNodeVolumes (World.cpp:58-69)
__global__ void k_node_volumes(MeshC m, double *__restrict__ node_vol)
{
    long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (u >= m.nn) return;
    int i = (int)(u % m.ni), j = (int)((u / m.ni) % m.nj), k = (int)(u / ((long long)m.ni * m.nj));
    double V = m.dh[0] * m.dh[1] * m.dh[2];
    if (i == 0 || i == m.ni - 1) V *= 0.5;
    if (j == 0 || j == m.nj - 1) V *= 0.5;
    if (k == 0 || k == m.nk - 1) V *= 0.5;
    node_vol[u] = V;
}

// World::addSphere (World.cpp:87-105), node position from World::pos (World.h:84-94)
__global__ void k_add_sphere(MeshC m, double phi_sphere, int32_t *__restrict__ object_id, double *__restrict__ phi)
{
    long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (u >= m.nn) return;
    int i = (int)(u % m.ni), j = (int)((u / m.ni) % m.nj), k = (int)(u / ((long long)m.ni * m.nj));
    double x = m.x0[0] + m.dh[0] * (double)i;
    double y = m.x0[1] + m.dh[1] * (double)j;
    double z = m.x0[2] + m.dh[2] * (double)k;
    if (in_sphere(m, x, y, z)) { object_id[u] = 1; phi[u] = phi_sphere; }
}

// World::addInlet (World.cpp:108-115)
__global__ void k_add_inlet(MeshC m, int32_t *__restrict__ object_id, double *__restrict__ phi)
{
    long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (u >= (long long)m.ni * m.nj) return;
    object_id[u] = 2;
    phi[u] = 0;
}

static inline unsigned nblk(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// ef (API layout, 3 doubles per node) -> ef4 (gather layout, 4 doubles per node)
__global__ void k_repack_ef(long long nn, const double *__restrict__ ef, double *__restrict__ ef4)
{
    long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (u >= nn) return;
    double4 v = make_double4(ef[3 * u], ef[3 * u + 1], ef[3 * u + 2], 0.0);
    *reinterpret_cast<double4 *>(ef4 + 4 * u) = v;
}

int espic_repack_ef(espic_ctx *c)
{
    k_repack_ef<<<(unsigned)((c->m.nn + 255) / 256), 256, 0, c->stream>>>(c->m.nn, c->ef, c->ef4);
    LAUNCH_CHECK(c);
    return 0;
}

// ---- lifetime -------------------------------------------------------------------------------------

extern "C" int espic_create(espic_ctx **out, int ni, int nj, int nk, const double x0[3], const double xm[3], int device)
{
    if (!out || ni < 3 || nj < 3 || nk < 3) { espic_set_error("espic_create: need ni,nj,nk >= 3"); return -1; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        espic_set_error("espic_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return -2;
    }
    if (device < 0 || device >= ndev) { espic_set_error("espic_create: device %d of %d", device, ndev); return -3; }
    CK(cudaSetDevice(device));
    // ESPIC_SYNC_MODE=block|yield: how host threads wait in cudaStreamSynchronize (default: the driver's choice, which spins).
    // With one process per GPU and fewer host cores than waiting threads, spinning ranks steal each other's time slices.
    if (const char *sm = getenv("ESPIC_SYNC_MODE")) {
        if (!strcmp(sm, "block")) cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync);
        else if (!strcmp(sm, "yield")) cudaSetDeviceFlags(cudaDeviceScheduleYield);
        cudaGetLastError();      // (a context created with other flags keeps them on older drivers: not an error for us)
    }
    espic_ctx *c = new espic_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    MeshC &m = c->m;
    m.ni = ni; m.nj = nj; m.nk = nk;
    m.nn = (long long)ni * nj * nk;
    const int nn3[3] = {ni, nj, nk};
    for (int a = 0; a < 3; a++) {
        // World::setExtents (World.cpp:22-36)
        m.x0[a] = x0[a];
        m.xm[a] = xm[a];
        m.dh[a] = (xm[a] - x0[a]) / (nn3[a] - 1);
        m.rdh[a] = 1.0 / m.dh[a];
        c->xc[a] = (x0[a] + xm[a]) * 0.5;
        m.sc[a] = 0;
    }
    m.sr2 = 0;
    size_t nb = (size_t)m.nn * sizeof(double);
    CK(cudaMalloc(&c->phi, nb));
    CK(cudaMalloc(&c->rho, nb));
    CK(cudaMalloc(&c->ef, 3 * nb));
    CK(cudaMalloc(&c->ef4, 4 * nb));
    CK(cudaMalloc(&c->node_vol, nb));
    CK(cudaMalloc(&c->object_id, (size_t)m.nn * sizeof(int32_t)));
    CK(cudaMalloc(&c->dscal, 128 * sizeof(unsigned long long)));
    CK(cudaMallocHost(&c->hpin, 128 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(c->phi, 0, nb, c->stream));
    CK(cudaMemsetAsync(c->rho, 0, nb, c->stream));
    CK(cudaMemsetAsync(c->ef, 0, 3 * nb, c->stream));
    CK(cudaMemsetAsync(c->ef4, 0, 4 * nb, c->stream));
    CK(cudaMemsetAsync(c->object_id, 0, (size_t)m.nn * sizeof(int32_t), c->stream));
    CK(cudaMemsetAsync(c->dscal, 0, 128 * sizeof(unsigned long long), c->stream));
    k_node_volumes<<<nblk(m.nn, 256), 256, 0, c->stream>>>(m, c->node_vol);
    LAUNCH_CHECK(c);
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_stage, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_copied, cudaEventDisableTiming));
    *out = c;
    return 0;
}

extern "C" void espic_destroy(espic_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    espic_comm_destroy(c);
    espic_mg_destroy(c);
    espic_migrate_destroy(c);
    cudaFree(c->phi); cudaFree(c->rho); cudaFree(c->ef); cudaFree(c->ef4); cudaFree(c->node_vol); cudaFree(c->object_id);
    for (int s = 0; s < c->nsp; s++) {
        for (int q = 0; q < 7; q++) { cudaFree(c->sp[s].p[q]); cudaFree(c->sp[s].alt[q]); }
        cudaFree(c->sp[s].den); cudaFree(c->sp[s].den_ave); cudaFree(c->sp[s].acc); cudaFree(c->sp[s].mom); cudaFree(c->sp[s].mpc);
        cudaFree(c->sp[s].kill_words); cudaFree(c->sp[s].leave_words);
    }
    cudaFree(c->dead_words); cudaFree(c->hit_words); cudaFree(c->scan_pre); cudaFree(c->scan_coff); cudaFree(c->lists);
    cudaFree(c->red); cudaFree(c->dscal); cudaFree(c->cell_cnt); cudaFree(c->sort_key); cudaFree(c->sort_src); cudaFree(c->node_type);
    for (int q = 0; q < 8; q++) cudaFree(c->sv[q]);
    if (c->push_ev0) { cudaEventDestroy(c->push_ev0); cudaEventDestroy(c->push_ev1); }
    cudaFreeHost(c->hpin);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->ev_stage) { cudaEventDestroy(c->ev_stage); cudaEventDestroy(c->ev_snap); cudaEventDestroy(c->ev_copied); }
    cudaFree(c->stage); cudaFree(c->snap);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int espic_set_stream(espic_ctx *c, void *stream)
{
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    if (c->own_stream) { CK(cudaStreamDestroy(c->stream)); c->own_stream = false; }
    c->stream = (cudaStream_t)stream;
    return 0;
}

extern "C" int espic_sync(espic_ctx *c)
{
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" long long espic_kernel_launches(espic_ctx *c) { return c->launches; }

extern "C" int espic_get_mesh(espic_ctx *c, double dh[3], double xc[3])
{
    for (int a = 0; a < 3; a++) { dh[a] = c->m.dh[a]; xc[a] = c->xc[a]; }
    return 0;
}

// ---- geometry ---------------------------------------------------------------------------------------

extern "C" int espic_add_sphere(espic_ctx *c, const double ctr[3], double radius, double phi_sphere)
{
    CK(cudaSetDevice(c->device));
    for (int a = 0; a < 3; a++) c->m.sc[a] = ctr[a];
    c->m.sr2 = radius * radius;
    k_add_sphere<<<nblk(c->m.nn, 256), 256, 0, c->stream>>>(c->m, phi_sphere, c->object_id, c->phi);
    LAUNCH_CHECK(c);
    c->geom_version++;
    return 0;
}

extern "C" int espic_add_inlet(espic_ctx *c)
{
    CK(cudaSetDevice(c->device));
    k_add_inlet<<<nblk((long long)c->m.ni * c->m.nj, 256), 256, 0, c->stream>>>(c->m, c->object_id, c->phi);
    LAUNCH_CHECK(c);
    c->geom_version++;
    return 0;
}

// ---- fields -------------------------------------------------------------------------------------------

static int field_ptr(espic_ctx *c, int which, int sp, void **p, size_t *bytes)
{
    size_t nn = (size_t)c->m.nn;
    if (which >= ESPIC_DEN && (sp < 0 || sp >= c->nsp)) {
        espic_set_error("field: bad species %d", sp);
        return -1;
    }
    if (which >= ESPIC_VEL && which <= ESPIC_NWW_SUM) {
        if (espic_ensure_moments(c, sp)) return -1;
        double *m = c->sp[sp].mom;
        // layout: n_sum | nv_sum[3] | nuu | nvv | nww | vel[3] | T
        switch (which) {
            case ESPIC_N_SUM: *p = m; *bytes = nn * 8; return 0;
            case ESPIC_NV_SUM: *p = m + nn; *bytes = nn * 24; return 0;
            case ESPIC_NUU_SUM: *p = m + 4 * nn; *bytes = nn * 8; return 0;
            case ESPIC_NVV_SUM: *p = m + 5 * nn; *bytes = nn * 8; return 0;
            case ESPIC_NWW_SUM: *p = m + 6 * nn; *bytes = nn * 8; return 0;
            case ESPIC_VEL: *p = m + 7 * nn; *bytes = nn * 24; return 0;
            case ESPIC_T: *p = m + 10 * nn; *bytes = nn * 8; return 0;
        }
    }
    switch (which) {
        case ESPIC_PHI: *p = c->phi; *bytes = nn * 8; return 0;
        case ESPIC_RHO: *p = c->rho; *bytes = nn * 8; return 0;
        case ESPIC_EF: *p = c->ef; *bytes = nn * 24; return 0;
        case ESPIC_NODE_VOL: *p = c->node_vol; *bytes = nn * 8; return 0;
        case ESPIC_OBJECT_ID: *p = c->object_id; *bytes = nn * 4; return 0;
        case ESPIC_DEN: *p = c->sp[sp].den; *bytes = nn * 8; return 0;
        case ESPIC_DEN_AVE: *p = c->sp[sp].den_ave; *bytes = nn * 8; return 0;
        case ESPIC_MPC: {
            const size_t nc = (size_t)(c->m.ni - 1) * (c->m.nj - 1) * (c->m.nk - 1);
            if (!c->sp[sp].mpc) {
                if (cudaMalloc(&c->sp[sp].mpc, nc * 8) != cudaSuccess) { espic_set_error("field: out of memory (mpc)"); return -1; }
                cudaMemsetAsync(c->sp[sp].mpc, 0, nc * 8, c->stream);
            }
            *p = c->sp[sp].mpc; *bytes = nc * 8; return 0;
        }
    }
    espic_set_error("field: unknown id %d", which);
    return -1;
}

int espic_ensure_moments(espic_ctx *c, int sp)
{
    Species &s = c->sp[sp];
    if (s.mom) return 0;
    const size_t nb = (size_t)c->m.nn * 11 * sizeof(double);
    CK(cudaMalloc(&s.mom, nb));
    CK(cudaMemsetAsync(s.mom, 0, nb, c->stream));
    return 0;
}

extern "C" int espic_field_download(espic_ctx *c, int which, int sp, void *host)
{
    void *p; size_t b;
    CK(cudaSetDevice(c->device));
    if (field_ptr(c, which, sp, &p, &b)) return -1;
    CK(cudaMemcpyAsync(host, p, b, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// The same download without stalling the compute stream: the field is snapshotted device-to-device on the compute stream
// (microseconds), the snapshot travels to `host` (pinned memory, or the copy degrades to a synchronous one) on the copy
// stream while the next step computes.  espic_copy_sync waits for it; one transfer in flight at a time.
extern "C" int espic_field_download_async(espic_ctx *c, int which, int sp, void *host)
{
    void *p; size_t b;
    CK(cudaSetDevice(c->device));
    if (field_ptr(c, which, sp, &p, &b)) return -1;
    if (c->copy_pending) CK(cudaStreamWaitEvent(c->stream, c->ev_copied, 0));       // the snapshot buffer is still being read
    if (b > c->snap_cap) {
        CK(cudaStreamSynchronize(c->copy_stream));
        if (c->snap) CK(cudaFree(c->snap));
        CK(cudaMalloc(&c->snap, b));
        c->snap_cap = b;
    }
    CK(cudaMemcpyAsync(c->snap, p, b, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaEventRecord(c->ev_snap, c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->ev_snap, 0));
    CK(cudaMemcpyAsync(host, c->snap, b, cudaMemcpyDeviceToHost, c->copy_stream));
    CK(cudaEventRecord(c->ev_copied, c->copy_stream));
    c->copy_pending = true;
    return 0;
}

extern "C" int espic_copy_sync(espic_ctx *c)
{
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->copy_stream));
    c->copy_pending = false;
    return 0;
}

extern "C" int espic_field_upload(espic_ctx *c, int which, int sp, const void *host)
{
    void *p; size_t b;
    CK(cudaSetDevice(c->device));
    if (field_ptr(c, which, sp, &p, &b)) return -1;
    CK(cudaMemcpyAsync(p, host, b, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (which == ESPIC_OBJECT_ID) c->geom_version++;
    if (which == ESPIC_EF) { int r = espic_repack_ef(c); if (r) return r; }
    if (which == ESPIC_DEN) c->sp[sp].acc_fresh = false;
    return 0;
}

extern "C" int espic_field_devptr(espic_ctx *c, int which, int sp, void **dptr)
{
    size_t b;
    return field_ptr(c, which, sp, dptr, &b);
}

// ---- species ------------------------------------------------------------------------------------------

extern "C" int espic_species_create(espic_ctx *c, double mass, double charge, double mpw0, long long capacity)
{
    CK(cudaSetDevice(c->device));
    if (c->nsp >= ESPIC_MAX_SPECIES) { espic_set_error("too many species (max %d)", ESPIC_MAX_SPECIES); return -1; }
    int id = c->nsp;
    Species &s = c->sp[id];
    s = Species();
    s.mass = mass; s.charge = charge; s.mpw0 = mpw0;
    s.mpw_max = mpw0 > 0 ? mpw0 : 0;
    size_t nb = (size_t)c->m.nn * sizeof(double);
    CK(cudaMalloc(&s.den, nb));
    CK(cudaMalloc(&s.den_ave, nb));
    CK(cudaMalloc(&s.acc, nb));
    CK(cudaMemsetAsync(s.den, 0, nb, c->stream));
    CK(cudaMemsetAsync(s.den_ave, 0, nb, c->stream));
    c->nsp++;
    int r = espic_species_reserve(c, id, capacity > 0 ? capacity : 1024);
    if (r) return r;
    return id;
}

#define SP_CHECK(c, sp) do { if ((sp) < 0 || (sp) >= (c)->nsp) { espic_set_error("bad species id %d", (sp)); return -1; } } while (0)

extern "C" int espic_species_reserve(espic_ctx *c, int sp, long long capacity)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    Species &s = c->sp[sp];
    if (capacity <= s.cap) return 0;
    // grow geometrically once particles exist (a source appends a few thousand particles every step; reallocating seven
    // arrays per step costs ~40 ms of cudaMalloc/cudaFree), and round up so 128-bit vector access of whole warps never
    // leaves the allocation
    if (s.cap > 0 && s.np > 0) capacity = std::max(capacity, s.cap + s.cap / 2);
    long long ncap = (capacity + 1023) / 1024 * 1024;
    for (int q = 0; q < 7; q++) {
        double *np_ = nullptr;
        CK(cudaMalloc(&np_, (size_t)ncap * sizeof(double)));
        if (s.p[q] && s.np > 0)
            CK(cudaMemcpyAsync(np_, s.p[q], (size_t)s.np * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        if (s.p[q]) { CK(cudaStreamSynchronize(c->stream)); CK(cudaFree(s.p[q])); }
        s.p[q] = np_;
    }
    s.cap = ncap;
    return 0;
}

extern "C" long long espic_species_count(espic_ctx *c, int sp)
{
    SP_CHECK(c, sp);
    return c->sp[sp].np;
}

extern "C" int espic_species_upload(espic_ctx *c, int sp, const double *const comp[7], long long n, int append)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    Species &s = c->sp[sp];
    s.diag_valid = false;
    MIG_GUARD(c, s, "espic_species_upload");
    long long base = append ? s.np : 0;
    int r = espic_species_reserve(c, sp, base + n);
    if (r) return r;
    for (int q = 0; q < 7; q++)
        if (n > 0) CK(cudaMemcpyAsync(s.p[q] + base, comp[q], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (long long i = 0; i < n; i++) if (comp[6][i] > s.mpw_max) s.mpw_max = comp[6][i];
    s.np = base + n;
    if (!append) s.n_settled = s.np;     // a restored state: every particle finished its last step (ch4 Particle::dt = 0)
    s.acc_fresh = false;
    s.pushes_since_sort = 1 << 20;       // arbitrary order
    return 0;
}

extern "C" int espic_species_upload_device(espic_ctx *c, int sp, const double *const dcomp[7], long long n, double mpw_max, int append)
{
    SP_CHECK(c, sp);
    CK(cudaSetDevice(c->device));
    Species &s = c->sp[sp];
    // appending behind a packed migration is how arrivals come in (espic_migrate_pack -> upload_device(append) -> espic_migrate_finish)
    s.diag_valid = false;
    if (!(append && s.mig_stage == 2)) MIG_GUARD(c, s, "espic_species_upload_device");
    long long base = append ? s.np : 0;
    int r = espic_species_reserve(c, sp, base + n);
    if (r) return r;
    for (int q = 0; q < 7; q++)
        if (n > 0) CK(cudaMemcpyAsync(s.p[q] + base, dcomp[q], (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    if (mpw_max > s.mpw_max) s.mpw_max = mpw_max;
    s.np = base + n;
    if (!append) s.n_settled = s.np;     // a restored state: every particle finished its last step (ch4 Particle::dt = 0)
    s.acc_fresh = false;
    s.pushes_since_sort = 1 << 20;       // arbitrary order
    return 0;
}

extern "C" long long espic_species_download(espic_ctx *c, int sp, double *const comp[7], long long n_max)
{
    SP_CHECK(c, sp);
    if (cudaSetDevice(c->device) != cudaSuccess) return -1;
    Species &s = c->sp[sp];
    long long n = std::min(n_max, s.np);
    for (int q = 0; q < 7; q++)
        if (n > 0 && cudaMemcpyAsync(comp[q], s.p[q], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) {
            espic_set_error("species download failed");
            return -1;
        }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { espic_set_error("species download sync failed"); return -1; }
    return n;
}
