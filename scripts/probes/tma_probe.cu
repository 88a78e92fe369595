// Stand-alone probe of the TMA tile path used by espic_mg.cuh (development aid): one block loads a (34 x 10 x 1) FP64 box and a
// (36 x 10 x 1) FP32 box at possibly out-of-range coordinates through cp.async.bulk.tensor.3d and writes them back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe [variant]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %d (%s) at line %d\n", (int)e_, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct Args {
    int nx, ny, nz;
    double *out_d; float *out_f;
    int c0, c1, c2;
    int which, bytes;
    alignas(64) CUtensorMap md, mf;
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <bool GLOBAL_MAP>
__global__ void k_probe(const __grid_constant__ Args a, const CUtensorMap *gmaps)
{
    extern __shared__ __align__(128) unsigned char smem[];
    double *sd = reinterpret_cast<double *>(smem);
    float *sf = reinterpret_cast<float *>(smem + 2816);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + 2816 + 1536);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap *md = GLOBAL_MAP ? gmaps : &a.md, *mf = GLOBAL_MAP ? gmaps + 1 : &a.mf;
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(a.bytes) : "memory");
        if (a.which & 1) asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(smem_u32(sd)), "l"((unsigned long long)md), "r"(smem_u32(bar)), "r"(a.c0), "r"(a.c1), "r"(a.c2) : "memory");
        if (a.which & 2) asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(smem_u32(sf)), "l"((unsigned long long)mf), "r"(smem_u32(bar)), "r"(a.c0), "r"(a.c1), "r"(a.c2) : "memory");
    }
    unsigned done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    } while (!done);
    for (int t = threadIdx.x; t < 340; t += blockDim.x) a.out_d[t] = sd[t];
    for (int t = threadIdx.x; t < 360; t += blockDim.x) a.out_f[t] = sf[t];
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int nx = argc > 2 ? atoi(argv[2]) : 128, ny = argc > 3 ? atoi(argv[3]) : 128, nz = 8;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)p;
    printf("variant %d, mesh %d x %d x %d, encoder %p\n", variant, nx, ny, nz, p);
    const size_t n = (size_t)nx * ny * nz;
    std::vector<double> hd(n); std::vector<float> hf(n);
    for (size_t i = 0; i < n; i++) { hd[i] = 1.0 + i; hf[i] = 0.5f + (float)(i % 4096); }
    double *dd; float *df; double *od; float *of;
    CK(cudaMalloc(&dd, n * 8)); CK(cudaMalloc(&df, n * 4)); CK(cudaMalloc(&od, 340 * 8)); CK(cudaMalloc(&of, 360 * 4));
    CK(cudaMemcpy(dd, hd.data(), n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(df, hf.data(), n * 4, cudaMemcpyHostToDevice));
    Args a;
    a.nx = nx; a.ny = ny; a.nz = nz; a.out_d = od; a.out_f = of;
    a.c0 = argc > 4 ? atoi(argv[4]) : -1; a.c1 = argc > 5 ? atoi(argv[5]) : -1; a.c2 = argc > 6 ? atoi(argv[6]) : 2;
    a.which = argc > 7 ? atoi(argv[7]) : 3;
    const int bwd = argc > 8 ? atoi(argv[8]) : 34, bwf = argc > 9 ? atoi(argv[9]) : 36;
    a.bytes = ((a.which & 1) ? bwd * 10 * 8 : 0) + ((a.which & 2) ? bwf * 10 * 4 : 0);
    printf("which %d box widths %d %d coords %d %d %d\n", a.which, bwd, bwf, a.c0, a.c1, a.c2);
    for (int f64 = 0; f64 < 2; f64++) {
        const cuuint64_t es = f64 ? 8 : 4;
        const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
        const cuuint64_t strides[2] = {nx * es, (cuuint64_t)nx * ny * es};
        const cuuint32_t box[3] = {(cuuint32_t)(f64 ? bwd : bwf), 10, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(f64 ? &a.md : &a.mf, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, f64 ? (void *)dd : (void *)df,
                         dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode f64=%d -> %d\n", f64, (int)r);
    }
    CUtensorMap *gm;
    CK(cudaMalloc(&gm, 2 * sizeof(CUtensorMap)));
    CK(cudaMemcpy(gm, &a.md, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(gm + 1, &a.mf, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    const size_t smem = 2816 + 1536 + 64;
    if (variant == 0) k_probe<false><<<1, 128, smem>>>(a, gm);
    else if (variant == 1) k_probe<true><<<1, 128, smem>>>(a, gm);
    else {
        void *args[] = {&a, &gm};
        CK(cudaLaunchCooperativeKernel((void *)k_probe<false>, dim3(1), dim3(128), args, smem, 0));
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<double> rd(340); std::vector<float> rf(360);
    CK(cudaMemcpy(rd.data(), od, 340 * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rf.data(), of, 360 * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int y = 0; y < 10; y++)
        for (int x = 0; x < 36; x++) {
            const int gx = a.c0 + x, gy = a.c1 + y, gz = a.c2;
            const bool in = gx >= 0 && gx < nx && gy >= 0 && gy < ny && gz >= 0 && gz < nz;
            const size_t u = ((size_t)gz * ny + gy) * nx + gx;
            if ((a.which & 1) && x < bwd) { const double want = in ? hd[u] : 0.0; if (rd[y * bwd + x] != want) { if (bad < 5) printf("d(%d,%d): got %g want %g\n", x, y, rd[y * bwd + x], want); bad++; } }
            const float wf = in ? hf[u] : 0.f; if ((a.which & 2) && x < bwf && rf[y * bwf + x] != wf) { if (bad < 5) printf("f(%d,%d): got %g want %g\n", x, y, rf[y * bwf + x], wf); bad++; }
        }
    printf("variant %d: %s (%d mismatches)\n", variant, bad ? "FAIL" : "ok", bad);
    return bad != 0;
}
