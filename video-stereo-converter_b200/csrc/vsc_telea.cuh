// vsc_telea.cuh — exact GPU implementation of the reference's hole filling:
//   cv2.dilate(mask, ones(3,3)) + cv2.inpaint(img, mask, 3, INPAINT_TELEA)
//   (/root/reference/helper/stereo_core.py:436-457; OpenCV photo/inpaint.cpp).
//
// Telea's method is a fast-marching (priority-queue ordered) sweep that re-reads pixels it has just
// written, so its result depends on the global pop order.  That order only matters between pixels
// that can see each other (window radius 3 + 1 for gradients, outer distance ring radius 3), so the
// image is decomposed into independent *clusters* of holes:
//   1. telea_prepare_kernel   morphology on the warp kernel's hole bitmap (1 bit per pixel): M = dilate3x3(hole),
//                             band, outer ring, initial T and order words near M, 8x8-tile occupancy and
//                             "reaches the kept window" counts; one warp per 32x16 tile, no shared memory
//   2. tile CCL kernels       union-find connected components over occupied 8x8 tiles; holes in
//                             different components are >= 15 px apart (> 2*range+2 = 8, the analytic
//                             independence bound), so components can be marched independently
//   3. telea_march_kernel     (vsc_march.cuh) one CTA per cluster, persistent CTAs pull clusters, big ones first:
//                             arrival times and computation order in bulk-synchronous generations, then the colours
//                             as a dataflow over that order
// Clusters with no hole pixel inside the kept (convergence-cropped) column window are skipped:
// their pixels are never read by the back end.
#pragma once
#include "vsc_kernels.cuh"

namespace vsc {

// state byte per pixel: bits 0-1 f (Telea flags), bits 2-3 o (outer-ring flags), bit 4 initial band
enum : unsigned char { F_KNOWN = 0, F_BAND = 1, F_INSIDE = 2, F_MASK = 3, O_BAND = 1 << 2, O_INSIDE = 2 << 2,
                       O_CHANGE = 3 << 2, O_MASK = 3 << 2, ST_BAND0 = 1 << 4 };

constexpr int TG = 8;   // tile edge for clustering

struct TeleaView {
    uchar4* img;            // [Hs][Ws] in/out (alpha = validity from the warp)
    const unsigned* holes;  // [Hs][wb] hole bitmap (bit i of word j <-> column 32 j + i; 1 = alpha 0)
    uint8_t* st;            // [Hs][Ws]
    float* tt;              // [Hs][Ws]
    // tile grid
    unsigned char* tile_cnt;   // [th*tw] number of M|band|ring pixels (<= 64)
    unsigned char* tile_need;  // [th*tw] number of M pixels of the tile inside the kept window
    int* lab;                  // [th*tw] union-find parent, -1 = empty tile
    int* csize;                // [th*tw] per-root pixel count
    int* ctiles;               // [th*tw] per-root tile count
    int* cneed;                // [th*tw] per-root number of M pixels inside the kept window
    int* cslot;                // [th*tw] per-root cluster slot
    // clusters
    int* cl_qoff;  int* cl_toff;  int* cl_ntiles;  int* cl_size;  int* cl_fill;
    int* tile_list;            // [ntiles_active]
    unsigned long long* qkey[3];   // march scratch: 6 + 3 arrays of qcap 32-bit words (see telea_march_kernel)
    unsigned* qidx[3];
    unsigned* pstate;          // [Hs][Ws] order word per pixel: position in the reference's computation order (see march)
    int qcap;
    FrameScalars* fs;          // per-frame counters of the frame this view belongs to
    int vi;                    // 0 = left, 1 = right eye within that frame
    int keep_x0, keep_x1;      // columns the back end reads
};

constexpr int TELEA_MAX_VIEWS = 8;   // a march launch covers up to 4 frames x 2 eyes
struct TeleaArgs {
    TeleaView v[TELEA_MAX_VIEWS];
    int Hs, Ws, tw, th;
    int wb;                      // words per row of the hole bitmaps
    int nviews;
    unsigned long long* stats;   // optional [2 views][2 passes][16] counters (VSC_TELEA_STATS builds)
};

// ---- 1. morphology ----------------------------------------------------------------------------
// Input: the hole bitmap the warp kernel wrote (1 bit per pixel, 32 columns per word).  One WARP owns a tile
// of 32 columns x 16 rows: lane r holds image row Y0 - 5 + r of the tile (+5 apron) as one 64-bit word, so the
// 3x3 dilation, the 4-neighbour band and the 7x7 / 9x9 neighbourhood tests are a few shifts, ORs and
// shuffles per row.  No shared memory, no CTA barrier; tiles without a hole nearby (the common case) only
// clear their state bytes.
__device__ __forceinline__ unsigned long long hdil(unsigned long long m, int r) {   // horizontal dilation by r
    unsigned long long o = m;
    for (int d = 1; d <= r; d++) o |= (m << d) | (m >> d);
    return o;
}
__device__ __forceinline__ unsigned long long shfl_up64(unsigned long long v, int d) {
    return ((unsigned long long)__shfl_up_sync(0xffffffffu, (unsigned)(v >> 32), d) << 32) | __shfl_up_sync(0xffffffffu, (unsigned)v, d);
}
__device__ __forceinline__ unsigned long long shfl_down64(unsigned long long v, int d) {
    return ((unsigned long long)__shfl_down_sync(0xffffffffu, (unsigned)(v >> 32), d) << 32) | __shfl_down_sync(0xffffffffu, (unsigned)v, d);
}
__device__ __forceinline__ unsigned vdil1(unsigned v) {    // OR with the rows above and below (lane -1 / +1)
    return v | __shfl_up_sync(0xffffffffu, v, 1) | __shfl_down_sync(0xffffffffu, v, 1);
}
__device__ __forceinline__ unsigned vdil2(unsigned v) {    // OR with the rows two above and two below
    return v | __shfl_up_sync(0xffffffffu, v, 2) | __shfl_down_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ unsigned spread4(unsigned nib) { return (nib * 0x00204081u) & 0x01010101u; }   // bit i -> byte i

constexpr int PR_ROWS = 16;     // output rows per warp (two 8x8 tile rows); window rows = PR_ROWS + 10 <= 32 lanes
__global__ void __launch_bounds__(kThreads) telea_prepare_kernel(const __grid_constant__ TeleaArgs a) {
    const int v = blockIdx.z;
    const TeleaView& V = a.v[v];
    const int lane = threadIdx.x;
    const int X0 = blockIdx.x * 32, Y0 = (blockIdx.y * (kThreads / 32) + threadIdx.y) * PR_ROWS;
    if (Y0 >= a.Hs) return;
    // lane <-> image row Y0 - 5 + lane; bit i of a row word <-> image column X0 - 5 + i (42 bits used)
    const int y = Y0 - 5 + lane;
    const bool yin = lane < PR_ROWS + 10 && y >= 0 && y < a.Hs;
    unsigned long long h = 0;
    if (yin) {
        const unsigned* row = V.holes + (size_t)y * a.wb;
        const int j = blockIdx.x;
        const unsigned wm = j > 0 ? row[j - 1] : 0u, w0 = row[j], wp = j + 1 < a.wb ? row[j + 1] : 0u;
        h = ((unsigned long long)(wm >> 27) | ((unsigned long long)w0 << 5) | ((unsigned long long)wp << 37)) & ((1ull << 42) - 1ull);
    }
    const bool orow = lane >= 5 && lane < 5 + PR_ROWS && y < a.Hs;     // this lane's row is an output row
    unsigned m32 = 0, band32 = 0, ring32 = 0, n4 = 0;
    const bool any_hole = __any_sync(0xffffffffu, h != 0ull);
    if (any_hole) {
        // columns of the window that lie inside the image
        const int clo = max(0, 5 - X0), chi = min(42, a.Ws - (X0 - 5));
        const unsigned long long cm = ((1ull << chi) - 1ull) & ~((1ull << clo) - 1ull);
        const unsigned long long hd = hdil(h, 1);
        unsigned long long M = (shfl_up64(hd, 1) | hd | shfl_down64(hd, 1)) & cm;      // dilate3x3(hole)
        if (!(yin && lane >= 1 && lane <= PR_ROWS + 8)) M = 0;
        const unsigned long long band = ~M & (shfl_up64(M, 1) | shfl_down64(M, 1) | (M << 1) | (M >> 1)) & cm;
        const unsigned long long A3 = hdil(M, 3), A4 = A3 | (M << 4) | (M >> 4);
        const unsigned n3 = vdil2(vdil1((unsigned)(A3 >> 5)));              // 7x7 neighbourhood of M
        const unsigned n4v = vdil1(vdil2(vdil1((unsigned)(A4 >> 5))));      // 9x9 neighbourhood of M
        if (orow) {
            m32 = (unsigned)(M >> 5); band32 = (unsigned)(band >> 5);
            ring32 = n3 & ~m32 & ~band32 & (unsigned)(cm >> 5);
            n4 = n4v;
        }
    }
    // ---- state bytes (and initial T / dataflow word near M): 4 rows x 8 lanes x 4 columns per step -------
    const bool vec = (a.Ws & 3) == 0;
#pragma unroll
    for (int it = 0; it < PR_ROWS / 4; it++) {
        const int rr = it * 4 + (lane >> 3), srcl = rr + 5, c0 = (lane & 7) * 4;
        const unsigned mm = (__shfl_sync(0xffffffffu, m32, srcl) >> c0) & 0xfu;
        const unsigned bb = (__shfl_sync(0xffffffffu, band32, srcl) >> c0) & 0xfu;
        const unsigned rg = (__shfl_sync(0xffffffffu, ring32, srcl) >> c0) & 0xfu;
        const unsigned nn = (__shfl_sync(0xffffffffu, n4, srcl) >> c0) & 0xfu;
        const int yy = Y0 + rr, xx = X0 + c0;
        if (yy >= a.Hs || xx >= a.Ws) continue;
        const unsigned stw = spread4(mm) * F_INSIDE | spread4(rg) * O_INSIDE | spread4(bb) * ST_BAND0;
        const size_t p = (size_t)yy * a.Ws + xx;
        if (vec) *reinterpret_cast<unsigned*>(V.st + p) = stw;         // Ws % 4 == 0: xx + 3 < Ws and p % 4 == 0
        else if ((a.Ws & 1) == 0) {                                    // Ws even: p even, pixels come in pairs
            *reinterpret_cast<unsigned short*>(V.st + p) = (unsigned short)stw;
            if (xx + 2 < a.Ws) *reinterpret_cast<unsigned short*>(V.st + p + 2) = (unsigned short)(stw >> 16);
        } else for (int k = 0; k < 4 && xx + k < a.Ws; k++) V.st[p + k] = (unsigned char)(stw >> (8 * k));
        if (nn) {
            for (int k = 0; k < 4 && xx + k < a.Ws; k++)
                if ((nn >> k) & 1u) { V.tt[p + k] = ((bb >> k) & 1u) ? 0.f : 1.0e6f; V.pstate[p + k] = 0xffffffffu; }
        }
    }
    // ---- 8x8 tile occupancy: lanes 5..12 and 13..20 each hold the rows of one tile row ------------------
    if (lane >= 5 && lane < 5 + PR_ROWS) {
        const int klo = min(32, max(0, V.keep_x0 - X0)), khi = max(0, min(32, V.keep_x1 - X0));
        const unsigned keep = khi > klo ? ((khi == 32 ? 0xffffffffu : ((1u << khi) - 1u)) & ~((1u << klo) - 1u)) : 0u;
        const unsigned occ = m32 | band32 | ring32, need = m32 & keep;
        unsigned pc = 0, pn = 0;       // four per-tile counts packed into bytes (each <= 8 per row, <= 64 per tile)
#pragma unroll
        for (int t = 0; t < 4; t++) {
            pc |= (unsigned)__popc(occ & (0xffu << (8 * t))) << (8 * t);
            pn |= (unsigned)__popc(need & (0xffu << (8 * t))) << (8 * t);
        }
        const int half = (lane - 5) >> 3;
        const unsigned gm = half ? 0x001fe000u : 0x00001fe0u;
        pc = __reduce_add_sync(gm, pc);
        pn = __reduce_add_sync(gm, pn);
        if (((lane - 5) & 7) == 0) {
            const int ty = Y0 / TG + half;
            if (ty < a.th)
                for (int t = 0; t < 4; t++) {
                    const int tx = X0 / TG + t;
                    if (tx < a.tw) {
                        V.tile_cnt[ty * a.tw + tx] = (unsigned char)(pc >> (8 * t));
                        V.tile_need[ty * a.tw + tx] = (unsigned char)(pn >> (8 * t));
                    }
                }
        }
    }
}

// hole bitmap from a validity byte map (stage API / tests; the pipeline's warp kernel writes the bitmap itself)
__global__ void pack_holes_kernel(const uint8_t* __restrict__ valid, int Hs, int Ws, int wb, unsigned* __restrict__ holes) {
    const int lane = threadIdx.x & 31;
    const size_t nw = (size_t)Hs * wb;
    for (size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5; w < nw; w += ((size_t)gridDim.x * blockDim.x) >> 5) {
        const int y = (int)(w / wb), x = (int)(w - (size_t)y * wb) * 32 + lane;
        const unsigned bits = __ballot_sync(0xffffffffu, x < Ws && valid[(size_t)y * Ws + x] == 0);
        if (lane == 0) holes[w] = bits;
    }
}

// ---- 2. connected components over occupied tiles (8-connectivity, union-find) -------------------
__device__ __forceinline__ int uf_find(int* lab, int x) {
    int p = lab[x];
    while (p != x) { x = p; p = lab[x]; }
    return x;
}
__device__ __forceinline__ void uf_union(int* lab, int a, int b) {
    while (true) {
        a = uf_find(lab, a);
        b = uf_find(lab, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        const int old = atomicMin(&lab[a], b);
        if (old == a) return;
        a = old;
    }
}

__global__ void telea_ccl_init_kernel(const __grid_constant__ TeleaArgs a) {
    const int n = a.tw * a.th;
    for (int v = 0; v < a.nviews; v++) {
        const TeleaView& V = a.v[v];
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
            V.lab[t] = V.tile_cnt[t] ? t : -1;
            V.csize[t] = 0; V.ctiles[t] = 0; V.cneed[t] = 0; V.cslot[t] = -1;
        }
    }
}
__global__ void telea_ccl_merge_kernel(const __grid_constant__ TeleaArgs a) {
    const int n = a.tw * a.th;
    for (int v = 0; v < a.nviews; v++) {
        const TeleaView& V = a.v[v];
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
            if (V.lab[t] < 0) continue;
            const int ty = t / a.tw, tx = t - ty * a.tw;
            if (tx > 0 && V.tile_cnt[t - 1]) uf_union(V.lab, t, t - 1);
            if (ty > 0) {
                if (V.tile_cnt[t - a.tw]) uf_union(V.lab, t, t - a.tw);
                if (tx > 0 && V.tile_cnt[t - a.tw - 1]) uf_union(V.lab, t, t - a.tw - 1);
                if (tx + 1 < a.tw && V.tile_cnt[t - a.tw + 1]) uf_union(V.lab, t, t - a.tw + 1);
            }
        }
    }
}
__global__ void telea_ccl_flatten_kernel(const __grid_constant__ TeleaArgs a) {
    const int n = a.tw * a.th;
    for (int v = 0; v < a.nviews; v++) {
        const TeleaView& V = a.v[v];
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
            if (V.lab[t] < 0) continue;
            const int root = uf_find(V.lab, t);
            atomicAdd(&V.csize[root], (int)V.tile_cnt[t]);
            atomicAdd(&V.ctiles[root], 1);
            if (V.tile_need[t]) atomicAdd(&V.cneed[root], (int)V.tile_need[t]);
        }
    }
}
__global__ void telea_cluster_alloc_kernel(const __grid_constant__ TeleaArgs a) {
    const int n = a.tw * a.th;
    for (int v = 0; v < a.nviews; v++) {
        const TeleaView& V = a.v[v];
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
            if (V.lab[t] != t || !V.cneed[t]) continue;
            // big clusters get slots from the front, small ones from the back: the work queue starts the
            // long poles first
            const int ci = V.csize[t] >= 1024 ? atomicAdd(&V.fs->nbig[V.vi], 1) : n - 1 - atomicAdd(&V.fs->nsmall[V.vi], 1);
            V.cl_qoff[ci] = atomicAdd(&V.fs->qbump[V.vi], V.csize[t]);
            V.cl_toff[ci] = atomicAdd(&V.fs->tbump[V.vi], V.ctiles[t]);
            V.cl_ntiles[ci] = V.ctiles[t];
            V.cl_size[ci] = V.cneed[t];      // hole pixels that the back end will read
            V.cl_fill[ci] = 0;
            V.cslot[t] = ci;
        }
    }
}
__global__ void telea_cluster_fill_kernel(const __grid_constant__ TeleaArgs a) {
    const int n = a.tw * a.th;
    for (int v = 0; v < a.nviews; v++) {
        const TeleaView& V = a.v[v];
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
            if (V.lab[t] < 0) continue;
            const int root = uf_find(V.lab, t);
            const int ci = V.cslot[root];
            if (ci < 0) continue;
            const int pos = atomicAdd(&V.cl_fill[ci], 1);
            V.tile_list[V.cl_toff[ci] + pos] = t;
        }
    }
}

// ---- 3. per-cluster march: one CTA per cluster, one warp per task -------------------------------------
struct TapConst { signed char dk[32]; signed char dl[32]; float dst[32]; };
__constant__ TapConst c_taps;   // 28 taps of the radius-3 disc in k-major raster order (centre excluded)

struct TapTable { int dk[32]; int dl[32]; float dst[32]; };    // shared-memory copy of c_taps (constant memory would serialise per lane)

// FastMarching_solve of OpenCV's inpaint.cpp: the arrival time of a pixel from two of its 4-neighbours
// (t1 / t2 their arrival times, in1 / in2 whether they are still INSIDE, i.e. unknown)
__device__ __forceinline__ float fmm_solve(float t1, float t2, bool in1, bool in2) {
    const double a11 = t1, a22 = t2;
    const double m12 = a11 < a22 ? a11 : a22;
    double sol;
    if (!in1) {
        if (!in2) {
            const double d = __dadd_rn(a11, -a22);
            if (fabs(d) >= 1.0) sol = __dadd_rn(1.0, m12);
            else sol = __dmul_rn(__dadd_rn(__dadd_rn(a11, a22), sqrt(__dadd_rn(2.0, -__dmul_rn(d, d)))), 0.5);
        } else sol = __dadd_rn(1.0, a11);
    } else if (!in2) sol = __dadd_rn(1.0, a22);
    else sol = __dadd_rn(1.0, m12);
    return (float)sol;
}

}  // namespace vsc
