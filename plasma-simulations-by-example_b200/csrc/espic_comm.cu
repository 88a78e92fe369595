// espic_comm.cu -- multi-GPU plumbing: one process per GPU, particles sharded by index, the deposited density
// summed over ranks with NCCL over NVLink/NVSwitch (replaces ch9/MPI Field::updateBoundaries,
// ch9/MPI/include/Field.h:122-179).  libnccl.so.2 is opened lazily so single-GPU use needs no NCCL at all.
#include "espic_internal.cuh"
#include <dlfcn.h>

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclInt64 = 4, ncclUint64 = 5, ncclFloat64 = 8 };   // ncclDataType_t (nccl.h)
enum { ncclSum = 0, ncclMax = 2 };                         // ncclRedOp_t

struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)(void) = nullptr;
    int (*GroupEnd)(void) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load()
{
    if (g_nccl.h) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { espic_set_error("cannot open libnccl.so.2: %s", dlerror()); return -1; }
    g_nccl.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.Send = (int (*)(const void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclSend");
    g_nccl.Recv = (int (*)(void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclRecv");
    g_nccl.GroupStart = (int (*)(void))dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)(void))dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce) { espic_set_error("libnccl lacks required symbols"); return -1; }
    g_nccl.h = h;
    return 0;
}

#define NCK(call) do { int r_ = (call); if (r_ != 0) { espic_set_error("NCCL error %d (%s): %s", r_, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?", #call); return -2000 - r_; } } while (0)

extern "C" int espic_comm_unique_id(void *id128)
{
    if (nccl_load()) return -1;
    ncclUniqueId id;
    NCK(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

extern "C" int espic_comm_init(espic_ctx *c, int rank, int nranks, const void *id128)
{
    CK(cudaSetDevice(c->device));
    if (nranks <= 1) { c->rank = 0; c->nranks = 1; return 0; }
    if (nccl_load()) return -1;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    NCK(g_nccl.CommInitRank(&comm, nranks, id, rank));
    c->nccl = comm;
    c->rank = rank;
    c->nranks = nranks;
    return 0;
}

void espic_comm_destroy(espic_ctx *c)
{
    if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->nccl);
    c->nccl = nullptr;
}

// largest value of v over all ranks (used to agree on the fixed-point scale)
int espic_comm_max_double(espic_ctx *c, double *v)
{
    if (c->nranks <= 1 || !c->nccl) return 0;
    double *d = reinterpret_cast<double *>(c->dscal + 32);
    double *h = reinterpret_cast<double *>(c->hpin) + 32;
    h[0] = *v;
    CK(cudaMemcpyAsync(d, h, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NCK(g_nccl.AllReduce(d, d, 1, ncclFloat64, ncclMax, (ncclComm_t)c->nccl, c->stream));
    CK(cudaMemcpyAsync(h, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *v = h[0];
    return 0;
}

// sum the scatter accumulator over ranks before it is divided by the node volumes
int espic_comm_allreduce_acc(espic_ctx *c, Species &s)
{
    if (c->nranks <= 1 || !c->nccl) return 0;
    int dt = (s.acc_mode == ESPIC_DEPOSIT_FIXED) ? ncclInt64 : ncclFloat64;
    NCK(g_nccl.AllReduce(s.acc, s.acc, (size_t)c->m.nn, dt, ncclSum, (ncclComm_t)c->nccl, c->stream));
    return 0;
}

// in-place all-gather of `count` doubles per rank: rank r's block sits at buf + r*count
int espic_comm_allgather_doubles(espic_ctx *c, double *buf, size_t count)
{
    if (c->nranks <= 1 || !c->nccl) return 0;
    if (!g_nccl.AllGather) { espic_set_error("libnccl lacks ncclAllGather"); return -1; }
    NCK(g_nccl.AllGather(buf + (size_t)c->rank * count, buf, count, ncclFloat64, (ncclComm_t)c->nccl, c->stream));
    return 0;
}

// in-place all-gather of `bytes` bytes per rank (device buffer)
int espic_comm_allgather_bytes(espic_ctx *c, void *buf, size_t bytes)
{
    if (c->nranks <= 1 || !c->nccl) return 0;
    if (!g_nccl.AllGather) { espic_set_error("libnccl lacks ncclAllGather"); return -1; }
    NCK(g_nccl.AllGather((char *)buf + (size_t)c->rank * bytes, buf, bytes, 0 /* ncclInt8 */, (ncclComm_t)c->nccl, c->stream));
    return 0;
}

// particle migration (espic_migrate.cuh): one NCCL group moves every (source, destination) segment.  A segment is SoA
// [7][count]; component q of the segment from `peer` lands at recvp[peer][q], i.e. straight in the receiver's particle arrays.
// Both sides know every count from the all-gathered matrix, so each send has exactly one matching receive, in component order.
int espic_comm_exchange_segments(espic_ctx *c, int parts, int me, const double *const *sendp, const long long *sendn,
                                 double *(*recvp)[7], const long long *recvn)
{
    if (c->nranks <= 1 || !c->nccl) return 0;
    if (!g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd) { espic_set_error("libnccl lacks ncclSend/ncclRecv"); return -1; }
    ncclComm_t comm = (ncclComm_t)c->nccl;
    NCK(g_nccl.GroupStart());
    for (int peer = 0; peer < parts; peer++) {
        if (peer == me) continue;
        for (int q = 0; q < 7; q++) {
            if (sendn[peer] > 0) NCK(g_nccl.Send(sendp[peer] + (size_t)q * sendn[peer], (size_t)sendn[peer], ncclFloat64, peer, comm, c->stream));
            if (recvn[peer] > 0) NCK(g_nccl.Recv(recvp[peer][q], (size_t)recvn[peer], ncclFloat64, peer, comm, c->stream));
        }
    }
    NCK(g_nccl.GroupEnd());
    return 0;
}

