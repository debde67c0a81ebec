// feasibility probe (not part of the product): SM partitioning with CUDA green contexts + runtime-API launches
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <set>
#define DRV(name) decltype(&name) p_##name = nullptr; { void* f = nullptr; cudaDriverEntryPointQueryResult q; \
    if (cudaGetDriverEntryPoint(#name, &f, cudaEnableDefault, &q) != cudaSuccess || !f) { printf("no %s\n", #name); return 1; } p_##name = (decltype(&name))f; }
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { printf("%s -> %d line %d\n", #x, (int)r_, __LINE__); return 1; } } while (0)
#define RT(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s -> %s line %d\n", #x, cudaGetErrorString(r_), __LINE__); return 1; } } while (0)
__global__ void who(unsigned* out, int spin) {
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long t0 = clock64(); while (clock64() - t0 < spin) {}
    if (threadIdx.x == 0) out[blockIdx.x] = smid;
}
int main() {
    RT(cudaSetDevice(0)); RT(cudaFree(0));
    DRV(cuDeviceGetDevResource) DRV(cuDevSmResourceSplitByCount) DRV(cuDevResourceGenerateDesc) DRV(cuGreenCtxCreate)
    DRV(cuGreenCtxStreamCreate) DRV(cuGreenCtxDestroy) DRV(cuDeviceGet)
    CUdevice dev; CK(p_cuDeviceGet(&dev, 0));
    CUdevResource all; CK(p_cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    printf("device SMs %u\n", all.sm.smCount);
    CUdevResource part, rest; unsigned n = 1;
    CK(p_cuDevSmResourceSplitByCount(&part, &n, &all, &rest, 0, 24));
    printf("split: groups %u, part %u SMs, rest %u SMs\n", n, part.sm.smCount, rest.sm.smCount);
    CUdevResourceDesc d0, d1; CK(p_cuDevResourceGenerateDesc(&d0, &part, 1)); CK(p_cuDevResourceGenerateDesc(&d1, &rest, 1));
    CUgreenCtx g0, g1; CK(p_cuGreenCtxCreate(&g0, d0, dev, CU_GREEN_CTX_DEFAULT_STREAM)); CK(p_cuGreenCtxCreate(&g1, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream s0, s1; CK(p_cuGreenCtxStreamCreate(&s0, g0, CU_STREAM_NON_BLOCKING, 0)); CK(p_cuGreenCtxStreamCreate(&s1, g1, CU_STREAM_NON_BLOCKING, 0));
    const int N = 4096; unsigned *o0, *o1; RT(cudaMalloc(&o0, N * 4)); RT(cudaMalloc(&o1, N * 4));
    cudaEvent_t e; RT(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    who<<<N, 128, 0, (cudaStream_t)s0>>>(o0, 20000); RT(cudaGetLastError());
    RT(cudaEventRecord(e, (cudaStream_t)s0));
    RT(cudaStreamWaitEvent((cudaStream_t)s1, e, 0));
    who<<<N, 128, 0, (cudaStream_t)s1>>>(o1, 20000); RT(cudaGetLastError());
    RT(cudaStreamSynchronize((cudaStream_t)s1)); RT(cudaStreamSynchronize((cudaStream_t)s0));
    static unsigned h0[N], h1[N]; RT(cudaMemcpy(h0, o0, N * 4, cudaMemcpyDeviceToHost)); RT(cudaMemcpy(h1, o1, N * 4, cudaMemcpyDeviceToHost));
    std::set<unsigned> a(h0, h0 + N), b(h1, h1 + N); int common = 0; for (unsigned x : a) common += b.count(x);
    printf("partition A used %zu SMs, partition B used %zu SMs, common %d\n", a.size(), b.size(), common);
    CK(p_cuGreenCtxDestroy(g0)); CK(p_cuGreenCtxDestroy(g1));
    printf("ok\n");
    return 0;
}
