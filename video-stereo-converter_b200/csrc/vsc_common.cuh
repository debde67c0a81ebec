// vsc_common.cuh — shared device helpers for the sm_100a SBS kernels.
//
// Float determinism contract: this translation unit is compiled with -fmad=false, so a*b+c is
// never contracted behind our back; every fused multiply-add below is an explicit fmaf()/fma().
// The operation order of each float stage is the one the CPU oracle documents, which in turn was
// identified against torch 2.11 / OpenCV 4.13 (see DESIGN.md "Float order").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vsc {

constexpr int kThreads = 256;

struct __align__(16) AxisTap {  // one output coordinate of F.interpolate(bilinear, align_corners=False)
    int i0, i1;
    float l0, l1;
};

struct GaussTaps {  // normalised 1-D Gaussian, k odd, k <= 31 (stereo_core.py:384, :430)
    int k;
    float g[31];
};

struct BilateralTaps {  // cv2.bilateralFilter circular window, radius <= 7 -> <= 149 taps
    int n;
    int radius;
    float w[152];
    signed char dy[152];
    signed char dx[152];
};

__device__ __forceinline__ int reflect_idx(int i, int n) {  // torch 'reflect' pad, |pad| < n
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

__device__ __forceinline__ int reflect101(int i, int n) {  // cv2 BORDER_REFLECT_101
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// order-preserving float <-> uint map for atomicMin/atomicMax on floats of either sign
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(unsigned u) {
    unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

// ---- CTA-wide byte copies with 128-bit vector accesses where alignment allows -----------------
// s must satisfy (s & 15) == (g & 15) so that both sides vectorise on the same 16-byte grid.
__device__ __forceinline__ void cta_copy_g2s(uint8_t* s, const uint8_t* __restrict__ g, int n) {
    const int tid = threadIdx.x + threadIdx.y * blockDim.x, nt = blockDim.x * blockDim.y;
    int head = (int)((16 - ((uintptr_t)g & 15)) & 15);
    if (head > n) head = n;
    for (int i = tid; i < head; i += nt) s[i] = g[i];
    const int nv = (n - head) >> 4;
    const uint4* gv = reinterpret_cast<const uint4*>(g + head);
    uint4* sv = reinterpret_cast<uint4*>(s + head);
    for (int i = tid; i < nv; i += nt) sv[i] = __ldg(gv + i);
    for (int i = head + (nv << 4) + tid; i < n; i += nt) s[i] = g[i];
}
__device__ __forceinline__ void cta_copy_s2g(uint8_t* __restrict__ g, const uint8_t* s, int n) {
    const int tid = threadIdx.x + threadIdx.y * blockDim.x, nt = blockDim.x * blockDim.y;
    int head = (int)((16 - ((uintptr_t)g & 15)) & 15);
    if (head > n) head = n;
    for (int i = tid; i < head; i += nt) g[i] = s[i];
    const int nv = (n - head) >> 4;
    uint4* gv = reinterpret_cast<uint4*>(g + head);
    const uint4* sv = reinterpret_cast<const uint4*>(s + head);
    for (int i = tid; i < nv; i += nt) gv[i] = sv[i];
    for (int i = head + (nv << 4) + tid; i < n; i += nt) g[i] = s[i];
}

// ---- deterministic pow for apply_depth_gamma (stereo_core.py:107) ------------------------------
// Same specification as the CPU oracle's orc_powf: table-driven log2 / exp2 in double built
// from IEEE-754 +,*,fma only, so CPU and GPU agree bit for bit; the result is the double value rounded to
// float (correctly rounded except for ~1e-7 of inputs).  `tab` = [1/c_i (128) | log2 c_i (128) | 2^(j/64) (64)],
// computed once on the host with the specification's fixed series (vsc_api.cu: pow_tables).
constexpr int kPowTabN = 320;
__device__ __forceinline__ double det_log2(double x, const double* tab) {      // x normal, positive
    const int hi = __double2hiint(x);
    const int e = ((hi >> 20) & 0x7ff) - 1023, i = (hi >> 13) & 127;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));    // [1, 2)
    const double r = fma(m, tab[i], -1.0);
    double q = 1.0 / 7.0;
    q = fma(q, r, -1.0 / 6.0);
    q = fma(q, r, 1.0 / 5.0);
    q = fma(q, r, -1.0 / 4.0);
    q = fma(q, r, 1.0 / 3.0);
    q = fma(q, r, -1.0 / 2.0);
    q = fma(q, r, 1.0);
    const double ln1p = __dmul_rn(r, q);
    return fma(ln1p, 1.4426950408889634074, __dadd_rn((double)e, tab[128 + i]));
}
__device__ __forceinline__ double det_exp2(double t, const double* tab) {
    const double k = rint(__dmul_rn(t, 64.0));
    const int ki = (int)k, j = ki & 63, n = (ki - j) / 64;
    const double f = __dmul_rn(__dadd_rn(t, -__dmul_rn(k, 0.015625)), 0.69314718055994530942);
    double p = 1.0 / 120.0;
    p = fma(p, f, 1.0 / 24.0);
    p = fma(p, f, 1.0 / 6.0);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    const double v = __dmul_rn(tab[256 + j], p);
    if (n < -1000 || n > 1000) return ldexp(v, n);      // would leave the normal range: let the library round
    return __hiloint2double(__double2hiint(v) + n * 1048576, __double2loint(v));      // exact scaling of a normal number
}
__device__ __forceinline__ float det_powf(float x, float g, const double* tab) {
    if (g == 2.0f) return __fmul_rn(x, x);
    if (g == 3.0f) return __fmul_rn(__fmul_rn(x, x), x);
    if (x == 1.0f) return 1.0f;
    return (float)det_exp2(__dmul_rn((double)g, det_log2((double)x, tab)), tab);
}

// ---- TMA bulk copy (cp.async.bulk, 1-D) of a byte range into shared memory ---------------------
// Rows of the frame are staged with the Blackwell copy engine instead of LSU loads: one elected
// thread arms an mbarrier with the byte count and issues the bulk copy, the CTA waits on the barrier.
// The engine needs 16-byte aligned addresses and sizes, so the copy starts at the enclosing aligned
// address and the caller reads at `smem + (g & 15)` (the same convention as cta_copy_g2s).  Up to 15
// bytes beyond the range are read: callers use the LSU path for the last row of a buffer.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* mbar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* mbar, unsigned phase) {
    unsigned done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_addr(mbar)), "r"(phase) : "memory");
    } while (!done);
}
// bytes the engine moves for the range [g, g+n): from the enclosing aligned address, rounded up to 16
__device__ __forceinline__ unsigned tma_span(const uint8_t* g, int n) {
    return (unsigned)((((uintptr_t)g & 15) + (unsigned)n + 15u) & ~15u);
}
// one thread: announce the total byte count of the copies that will complete on `mbar` (single arrival)
__device__ __forceinline__ void tma_expect(unsigned long long* mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(mbar)), "r"(bytes) : "memory");
}
// one thread: copy [g & ~15, roundup16(g + n)) to smem_base (16-byte aligned); completion is signalled on mbar
__device__ __forceinline__ void tma_copy_g2s(uint8_t* smem_base, const uint8_t* g, int n, unsigned long long* mbar) {
    const uint8_t* ga = reinterpret_cast<const uint8_t*>((uintptr_t)g & ~(uintptr_t)15);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(smem_base)), "l"(ga), "r"(tma_span(g, n)), "r"(smem_addr(mbar)) : "memory");
}

}  // namespace vsc
