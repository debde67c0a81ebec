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
//   3. telea_cluster_kernel   one CTA per cluster (persistent CTAs pull clusters, big ones first) runs the
//                             sequential algorithm as an ordered task dataflow (outer ring FMM, then the
//                             inpainting FMM); the 28 window taps of a pixel are evaluated one per lane and
//                             accumulated in the reference's raster order
// The priority queue (sorted list with FIFO ties in OpenCV) is realised as generations: all queued
// entries with T in [Tmin, Tmin+0.7) are extracted, sorted by (T, push order) and popped in order;
// anything pushed meanwhile has T >= popped T + 1/sqrt(2) and therefore belongs to a later
// generation, so the pop order is identical to the reference's.
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
    unsigned long long* qkey[3];   // two ping-pong pools + the current generation
    unsigned* qidx[3];
    unsigned* pstate;          // [Hs][Ws] order word per pixel for the dataflow (which pop / task owns it, see march)
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

// Per-warp shared-memory window around the pixel being popped.  Every access of one pop (fast-marching
// solve + Telea weights + gradients) stays within 5 pixels of the popped position, so the window is
// loaded once per pop with independent, coalesced-by-row loads (one memory round trip) and all the
// dependent arithmetic then runs out of shared memory; results are written through to global memory.
constexpr int WIN_R = 5, WIN_D = 2 * WIN_R + 1, WIN_S = WIN_D + 1;
struct TapTable { int dk[32]; int dl[32]; float dst[32]; };
struct WarpWin {
    unsigned char st[WIN_D * WIN_S];
    float tt[WIN_D * WIN_D];
    unsigned img[WIN_D * WIN_D];
    float taps[28 * 10];
};

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

struct Marcher {
    const TeleaView& V;
    int Hs, Ws;
    int lane;
    WarpWin* w;
    const TapTable* tp;   // shared-memory copy of the tap constants (constant memory would serialise per lane)
    int wy0, wx0;   // image coordinates of window element (0,0)
    __device__ __forceinline__ bool inb(int y, int x) const { return y >= 0 && y < Hs && x >= 0 && x < Ws; }
    // (re)load the window of radius r <= WIN_R centred on (yc, xc); pixels outside the image read as the
    // 1-pixel KNOWN frame (t = 1e6) that cv2.inpaint adds around the image
    template <int R, bool WITH_IMG> __device__ __forceinline__ void load(int yc, int xc) {
        wy0 = yc - WIN_R; wx0 = xc - WIN_R;
        constexpr int D = 2 * R + 1, NIT = (D * D + 31) / 32;
        unsigned char sv[NIT]; float tv[NIT]; unsigned iv[NIT];
#pragma unroll
        for (int it = 0; it < NIT; it++) {       // issue every load before the first use
            const int idx = lane + 32 * it;
            const int ly = idx / D + (WIN_R - R), lx = idx % D + (WIN_R - R);
            const int y = wy0 + ly, x = wx0 + lx;
            sv[it] = 0; tv[it] = 1.0e6f; iv[it] = 0;
            if (idx < D * D && inb(y, x)) {
                const size_t p = (size_t)y * Ws + x;
                sv[it] = V.st[p]; tv[it] = V.tt[p];
                if (WITH_IMG) iv[it] = *reinterpret_cast<const unsigned*>(&V.img[p]);
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            const int idx = lane + 32 * it;
            if (idx < D * D) {
                const int ly = idx / D + (WIN_R - R), lx = idx % D + (WIN_R - R);
                w->st[ly * WIN_S + lx] = sv[it]; w->tt[ly * WIN_D + lx] = tv[it];
                if (WITH_IMG) w->img[ly * WIN_D + lx] = iv[it];
            }
        }
        __syncwarp();
    }
    __device__ __forceinline__ unsigned char S(int y, int x) const { return w->st[(y - wy0) * WIN_S + (x - wx0)]; }
    __device__ __forceinline__ float Traw(int y, int x) const { return w->tt[(y - wy0) * WIN_D + (x - wx0)]; }
    template <bool OUTER> __device__ __forceinline__ bool inside(int y, int x) const {
        const unsigned char s = S(y, x);
        return OUTER ? ((s & O_MASK) == O_INSIDE) : ((s & F_MASK) == F_INSIDE);
    }
    // T as the inpainting pass sees it: the outer pass' distances are negated (icvCalcFMM negate=true)
    __device__ __forceinline__ float T_main(int y, int x) const {
        const float t = Traw(y, x);
        return ((S(y, x) & O_MASK) == O_CHANGE) ? -t : t;
    }
    template <bool OUTER> __device__ __forceinline__ float solve(int y1, int x1, int y2, int x2) const {
        return fmm_solve(OUTER ? Traw(y1, x1) : T_main(y1, x1), OUTER ? Traw(y2, x2) : T_main(y2, x2),
                         inside<OUTER>(y1, x1), inside<OUTER>(y2, x2));
    }
    // the four corner solves run on lanes 0..3; every lane returns the minimum
    template <bool OUTER> __device__ __forceinline__ float min4(int y, int x) const {
        const int l = lane & 3;
        const float s = solve<OUTER>(y + ((l & 1) ? 1 : -1), x, y, x + ((l & 2) ? 1 : -1));
        const float a = fminf(s, __shfl_xor_sync(0xffffffffu, s, 1));
        return fminf(a, __shfl_xor_sync(0xffffffffu, a, 2));
    }
    __device__ __forceinline__ int pix(int y, int x, int c) const { return (w->img[(y - wy0) * WIN_D + (x - wx0)] >> (8 * c)) & 0xff; }
    // write-through updates by lane 0
    __device__ __forceinline__ void set_T(int y, int x, float t) const {
        V.tt[(size_t)y * Ws + x] = t; w->tt[(y - wy0) * WIN_D + (x - wx0)] = t;
    }
    __device__ __forceinline__ void set_S(int y, int x, unsigned char s) const {
        V.st[(size_t)y * Ws + x] = s; w->st[(y - wy0) * WIN_S + (x - wx0)] = s;
    }
    // icvTeleaInpaintFMM body for one pixel (y,x) whose T was just set to `dist`; warp-cooperative
    __device__ void inpaint(int y, int x, float dist) const {
        float* sm = w->taps;
        // gradT (warp-uniform)
        float gtx, gty;
        {
            const bool r = !inside<false>(y, x + 1), l = !inside<false>(y, x - 1);
            const bool d = !inside<false>(y + 1, x), u = !inside<false>(y - 1, x);
            if (r) gtx = l ? __fmul_rn(__fsub_rn(T_main(y, x + 1), T_main(y, x - 1)), 0.5f) : __fsub_rn(T_main(y, x + 1), dist);
            else gtx = l ? __fsub_rn(dist, T_main(y, x - 1)) : 0.f;
            if (d) gty = u ? __fmul_rn(__fsub_rn(T_main(y + 1, x), T_main(y - 1, x)), 0.5f) : __fsub_rn(T_main(y + 1, x), dist);
            else gty = u ? __fsub_rn(dist, T_main(y - 1, x)) : 0.f;
        }
        bool valid = false;
        float term[10];
        if (lane < 28) {
            const int dk = tp->dk[lane], dl = tp->dl[lane];
            const int ky = y + dk, kx = x + dl;
            if (inb(ky, kx) && !inside<false>(ky, kx)) {
                valid = true;
                const float ry = (float)(-dk), rx = (float)(-dl);
                const float dst = tp->dst[lane];
                const float lev = (float)__ddiv_rn(1.0, __dadd_rn(1.0, fabs((double)__fsub_rn(T_main(ky, kx), dist))));
                float dir = __fadd_rn(__fmul_rn(rx, gtx), __fmul_rn(ry, gty));
                if (fabs((double)dir) <= 0.01) dir = 0.000001f;
                const float wgt = fabsf(__fmul_rn(__fmul_rn(dst, lev), dir));
                const bool fr = !inside<false>(ky, kx + 1), fl = !inside<false>(ky, kx - 1);
                const bool fd = !inside<false>(ky + 1, kx), fu = !inside<false>(ky - 1, kx);
                const int km = ky + (ky == 0), kp = ky - (ky == Hs - 1);
                const int lm = kx + (kx == 0), lp = kx - (kx == Ws - 1);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float gix, giy;
                    if (fr) gix = fl ? __fmul_rn((float)(pix(km, lp + 1, c) - pix(km, lm - 1, c)), 2.0f)
                                     : (float)(pix(km, lp + 1, c) - pix(km, lm, c));
                    else gix = fl ? (float)(pix(km, lp, c) - pix(km, lm - 1, c)) : 0.f;
                    if (fd) giy = fu ? __fmul_rn((float)(pix(kp + 1, lm, c) - pix(km - 1, lm, c)), 2.0f)
                                     : (float)(pix(kp + 1, lm, c) - pix(km, lm, c));
                    else giy = fu ? (float)(pix(kp, lm, c) - pix(km - 1, lm, c)) : 0.f;
                    term[c] = __fmul_rn(wgt, (float)pix(ky, kx, c));
                    term[3 + c] = __fmul_rn(wgt, __fmul_rn(gix, rx));
                    term[6 + c] = __fmul_rn(wgt, __fmul_rn(giy, ry));
                }
                term[9] = wgt;
            }
        }
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        if (valid) {
#pragma unroll
            for (int q = 0; q < 10; q++) sm[lane * 10 + q] = term[q];
        }
        __syncwarp();
        // lanes 0..9 each accumulate one quantity in tap (raster) order: Ia[3], Jx[3], Jy[3], s
        // (adding +0 for an absent tap leaves the accumulator bit-identical: it is never -0)
        float acc = lane == 9 ? 1.0e-20f : 0.f;
        {
            const int col = lane < 10 ? lane : 0;
            const bool sub = lane >= 3 && lane < 9;
            float t[28];
#pragma unroll
            for (int L = 0; L < 28; L++) t[L] = sm[L * 10 + col];
#pragma unroll
            for (int L = 0; L < 28; L++) {
                const float tv = ((vm >> L) & 1u) ? t[L] : 0.f;
                acc = sub ? __fsub_rn(acc, tv) : __fadd_rn(acc, tv);
            }
        }
        const float s = __shfl_sync(0xffffffffu, acc, 9);
        const float jx = __shfl_sync(0xffffffffu, acc, min(lane + 3, 31));
        const float jy = __shfl_sync(0xffffffffu, acc, min(lane + 6, 31));
        int outc = 0;
        if (lane < 3) {
            const float ia_s = __fdiv_rn(acc, s);
            const float jsum = __fadd_rn(jx, jy);
            const float jn = __fadd_rn(__fmul_rn(jx, jx), __fmul_rn(jy, jy));
            const double den = __dadd_rn(sqrt((double)jn), (double)1.0e-20f);
            const double val = __dadd_rn(__dadd_rn((double)ia_s, __ddiv_rn((double)jsum, den)), (double)0.5f);
            const float sat = (float)val;
            outc = min(max(__float2int_rn(sat), 0), 255);
        }
        const unsigned c0 = __shfl_sync(0xffffffffu, outc, 0), c1 = __shfl_sync(0xffffffffu, outc, 1),
                       c2 = __shfl_sync(0xffffffffu, outc, 2);
        __syncwarp();
        if (lane == 0) {
            const int li = (y - wy0) * WIN_D + (x - wx0);
            const unsigned nv = (w->img[li] & 0xff000000u) | c0 | (c1 << 8) | (c2 << 16);
            w->img[li] = nv;
            V.img[(size_t)y * Ws + x] = make_uchar4((unsigned char)c0, (unsigned char)c1, (unsigned char)c2, (unsigned char)(nv >> 24));
        }
    }
};

// ---- CTA-wide helpers --------------------------------------------------------------------------------
__device__ __forceinline__ void cmpswap(unsigned long long* key, unsigned* idx, int i, int p) {
    const unsigned long long a = key[i], b = key[p];
    if (a > b) {
        key[i] = b; key[p] = a;
        const unsigned t = idx[i]; idx[i] = idx[p]; idx[p] = t;
    }
}
// CTA-wide bitonic sort of n (key,idx) pairs in global memory, ascending by key.  All compare-exchanges
// are ascending (the "flip" formulation), so the virtual +inf padding above n never has to move and pairs
// whose partner is >= n are simply skipped.
__device__ void block_sort(unsigned long long* key, unsigned* idx, int n) {
    if (n <= 1) return;
    int N = 1;
    while (N < n) N <<= 1;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = 2; k <= N; k <<= 1) {
        for (int i = tid; i < n; i += nt) {
            const int p = i ^ (k - 1);
            if (p > i && p < n) cmpswap(key, idx, i, p);
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int i = tid; i < n; i += nt) {
                const int p = i ^ j;
                if (p > i && p < n) cmpswap(key, idx, i, p);
            }
            __syncthreads();
        }
    }
}

// Warps per march CTA (template parameter NW of the kernel).  Measured on a B200 at 1080p: with many frames in
// flight 8 warps give the best throughput (4 / 6 / 8 / 12 warps: 572 / 585 / 604 / 559 frames/s); a single frame
// alone finishes sooner with 16 (22 ms vs 35 ms), which is what a context with one slot gets.
constexpr int TELEA_WARPS_THROUGHPUT = 8, TELEA_WARPS_LATENCY = 16;
#ifndef VSC_OUTER_LANES
#define VSC_OUTER_LANES 1     // outer-ring distance pass: one task per lane (0 = one task per warp, the first implementation)
#endif
// Two compute tasks of one generation whose pixels are closer than this (Chebyshev) run in queue order;
// farther apart they commute.  Inpainting a pixel reads flags / T / colours within 4 of it and writes only
// the pixel itself -> 4.  An outer-ring distance reads the 4-neighbours and writes the pixel -> 1.
constexpr int TELEA_DC_MAIN = 4, TELEA_DC_OUTER = 1;

constexpr int TELEA_RING = 1024;      // completion ring entries (> tasks in flight: 32 per warp in the outer pass)
constexpr int TELEA_MAXDEP = 32;      // compact dependency list per warp (one entry per lane); more (rare): per-lane polling
template <int NW> struct MarchShared {
    int npool, npool2, ncur, ntask, next_t, done_t, gbase, scan_total, need_left;
    unsigned tmin;
    unsigned tbase;                    // CTA-monotonic index of the current generation's first task
    int ci;
    int wsum[NW];
    unsigned ring[TELEA_RING];         // ring[J % RING] = J + 1 once task J has finished (monotonic per entry)
    unsigned wdep[NW][TELEA_MAXDEP];
    WarpWin win[NW];
#ifdef VSC_TELEA_STATS
    unsigned long long c_wait, c_pop, c_sort, c_part, c_total, n_pops, n_pix, n_gen, n_polls, c_load, c_inp, c_rel, c_min4;
#endif
};

// pstate word per pixel: (index << 2) | kind
//   pop  of the current generation : (G << 2) | 1      G = global pop index (rank in sorted order)
//   compute task                   : (J << 2) | 2      J = CTA-monotonic task index = order in which a sequential run
//                                                      performs the tasks (also the FIFO tie-break of the queue)
//   never touched                  : 0xffffffff
// Whether task J has finished is NOT recorded here but in the CTA's shared-memory completion ring (MarchShared::ring):
// a waiting warp polls shared memory, not global memory.
__device__ __forceinline__ bool ps_is_pop(unsigned v) { return (v & 3u) == 1u; }
__device__ __forceinline__ bool ps_is_task(unsigned v) { return (v & 3u) == 2u; }


#if VSC_OUTER_LANES
// Outer-ring pass: a distance task only reads the arrival times and flags of its 4-neighbours, so a LANE runs
// a task: a warp claims 32 consecutive tasks, every lane collects the earlier tasks of the generation among
// its 4 neighbours, and the warp iterates until all its lanes have run (the earliest unfinished task of the
// CTA never waits, so this terminates).  Same order, same arithmetic, ~100x fewer warp instructions per task.
// (Its own function, not inlined: the main pass keeps its register allocation.)
template <int NW>
__device__ __noinline__ void outer_lane_dataflow(const TeleaView& V, MarchShared<NW>& sh, int Hs, int Ws, int ntask, int carry,
                                                 const unsigned* next_i, unsigned long long* next_k) {
    const int lane = threadIdx.x & 31;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&sh.next_t, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= ntask) break;
        const int j = base + lane;
        const bool active = j < ntask;
        const unsigned tbase = sh.tbase, J = tbase + (unsigned)j;
        unsigned pn = 0;
        int y = 0, x = 0;
        // a task reads (and is read by) its 4-neighbours only: those with an earlier task of this generation are
        // its dependencies (q: 0 up, 1 down, 2 left, 3 right)
        unsigned dep[4];
#pragma unroll
        for (int q = 0; q < 4; q++) dep[q] = 0xffffffffu;
        if (active) {
            pn = next_i[carry + j];
            y = (int)(pn / (unsigned)Ws); x = (int)(pn - (unsigned)y * (unsigned)Ws);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int yy = y + (q == 0 ? -1 : (q == 1 ? 1 : 0)), xx = x + (q == 2 ? -1 : (q == 3 ? 1 : 0));
                if ((yy >= 0 && yy < Hs && xx >= 0 && xx < Ws)) {
                    const unsigned v = V.pstate[(size_t)yy * Ws + xx];      // stable during the dataflow
                    if (ps_is_task(v) && (v >> 2) >= tbase && (v >> 2) < J) dep[q] = v >> 2;
                }
            }
        }
        bool done = !active;
        while (true) {
            bool ready = !done;
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (dep[q] != 0xffffffffu) {
                    if (*(volatile unsigned*)&sh.ring[dep[q] & (TELEA_RING - 1)] >= dep[q] + 1u) dep[q] = 0xffffffffu;
                    else ready = false;
                }
            if (ready) {
                __threadfence_block();
                // arrival times / flags of the 4-neighbours; outside the image: the KNOWN frame with T = 1e6
                float tn[4]; bool in_[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int yy = y + (q == 0 ? -1 : (q == 1 ? 1 : 0)), xx = x + (q == 2 ? -1 : (q == 3 ? 1 : 0));
                    tn[q] = 1.0e6f; in_[q] = false;
                    if ((yy >= 0 && yy < Hs && xx >= 0 && xx < Ws)) {
                        const size_t p = (size_t)yy * Ws + xx;
                        tn[q] = V.tt[p];       // ordered after the ring reads by the fence above (CTA scope: same SM, same L1)
                        in_[q] = (V.st[p] & O_MASK) == O_INSIDE;
                    }
                }
                // the four corner solves in min4's pairing: (up,left) (down,left) | (up,right) (down,right)
                const float s0 = fmm_solve(tn[0], tn[2], in_[0], in_[2]), s1 = fmm_solve(tn[1], tn[2], in_[1], in_[2]);
                const float s2 = fmm_solve(tn[0], tn[3], in_[0], in_[3]), s3 = fmm_solve(tn[1], tn[3], in_[1], in_[3]);
                const float dist = fminf(fminf(s0, s1), fminf(s2, s3));
                V.tt[pn] = dist;
                V.st[pn] = (unsigned char)((V.st[pn] & ~O_MASK) | O_BAND);
                next_k[carry + j] = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)J;
                volatile unsigned* slot = &sh.ring[J & (TELEA_RING - 1)];
                while (J >= (unsigned)TELEA_RING && *slot < J - (unsigned)TELEA_RING + 1u) __nanosleep(20);
                __threadfence_block();
                *slot = J + 1u;
                done = true;
            }
            const unsigned left = __ballot_sync(0xffffffffu, !done);
            if (!left) break;
            if (!__any_sync(0xffffffffu, ready)) __nanosleep(40);     // nobody moved: wait for another warp
        }
#ifdef VSC_TELEA_STATS
        if (lane == 0) atomicAdd(&sh.n_pix, (unsigned long long)min(32, ntask - base));
#endif
    }
}
#endif

// One fast-marching pass over one cluster, executed by a whole CTA.
//  * the queue is processed in generations (see file header); each generation is sorted CTA-wide.
//  * popping an entry computes those 4-neighbours that are still INSIDE; each such pixel is computed by the
//    FIRST popped neighbour (its owner).  Ownership only depends on the sorted order, so the list of compute
//    tasks of a generation, ordered by (owner rank, neighbour index) = the order in which a sequential run
//    performs them, is built up front in parallel.
//  * the tasks then run as a dataflow: warps claim tasks in order and a task starts once every earlier task
//    of its generation within TELEA_DC_* pixels has finished.  Tasks farther apart commute, so the result is
//    identical to the sequential order.  The dependencies of a task are read once from pstate (which pixel
//    belongs to which task), compacted into a short per-warp list, and then polled in the shared-memory
//    completion ring - a handful of instructions per poll instead of a sweep over global memory.
//  * the FIFO tie-break of the reference's queue is the task index J, i.e. the sequential push order.
template <bool OUTER, int NW>
__device__ void march(Marcher& mc, const TeleaView& V, MarchShared<NW>& sh, int qoff, int ntiles, const int* tiles, int tw,
                      int keep_x0, int keep_x1, unsigned long long* stats) {
    unsigned long long* pk[2] = {V.qkey[0] + qoff, V.qkey[1] + qoff};
    unsigned* pi[2] = {V.qidx[0] + qoff, V.qidx[1] + qoff};
    unsigned long long* cur_k = V.qkey[2] + qoff;
    unsigned* cur_i = V.qidx[2] + qoff;
    const int Ws = mc.Ws, Hs = mc.Hs, lane = mc.lane, wid = threadIdx.x >> 5, tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) { sh.npool = 0; sh.npool2 = 0; sh.ncur = 0; sh.ntask = 0; sh.next_t = 0; sh.done_t = 0; sh.tmin = 0xffffffffu; if (OUTER) sh.gbase = 1; }
#ifdef VSC_TELEA_STATS
    if (tid == 0) { sh.c_wait = sh.c_pop = sh.c_sort = sh.c_part = sh.n_pops = sh.n_pix = sh.n_gen = sh.n_polls = sh.c_load = sh.c_inp = sh.c_rel = sh.c_min4 = 0; }
    const long long t_start = clock64();
    long long t_mark;
#define STAT_MARK() t_mark = clock64()
#define STAT_ADD(field) do { if (lane == 0) atomicAdd(&sh.field, (unsigned long long)(clock64() - t_mark)); } while (0)
#define STAT_ADD0(field) do { if (tid == 0) atomicAdd(&sh.field, (unsigned long long)(clock64() - t_mark)); } while (0)
#define STAT_INC(field, n) do { if (lane == 0) atomicAdd(&sh.field, (unsigned long long)(n)); } while (0)
#define STAT_T0() const long long t_sub = clock64()
#define STAT_T1(field) do { if (lane == 0) atomicAdd(&sh.field, (unsigned long long)(clock64() - t_sub)); } while (0)
#else
#define STAT_MARK()
#define STAT_ADD(field)
#define STAT_ADD0(field)
#define STAT_INC(field, n)
#define STAT_T0()
#define STAT_T1(field)
#endif
    __syncthreads();
    // initial queue: the band pixels, T = 0, ordered by raster position (= linear index in the key)
    for (int ti = wid; ti < ntiles; ti += NW) {
        const int t = tiles[ti];
        const int ty = t / tw, tx = t - ty * tw;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int y = ty * TG + (lane >> 3) + 4 * h, x = tx * TG + (lane & 7);
            bool isb = false;
            unsigned p = 0;
            if (y < Hs && x < Ws) { p = (unsigned)y * (unsigned)Ws + (unsigned)x; isb = (V.st[p] & ST_BAND0) != 0; }
            const unsigned bm = __ballot_sync(0xffffffffu, isb);
            int base = 0;
            if (lane == 0 && bm) base = atomicAdd(&sh.npool, __popc(bm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (isb) {
                const int pos = base + __popc(bm & ((1u << lane) - 1));
                pk[0][pos] = (unsigned long long)p;
                pi[0][pos] = p;
            }
        }
    }
    __syncthreads();
    int src = 0;
    while (true) {
        const int npool = sh.npool;
        if (npool == 0) break;
        unsigned long long* pool_k = pk[src]; unsigned* pool_i = pi[src];
        unsigned long long* next_k = pk[src ^ 1]; unsigned* next_i = pi[src ^ 1];
        STAT_MARK();
        // ---- generation = entries with T < Tmin + 0.7 ---------------------------------------------------
        unsigned tmin = 0xffffffffu;
        for (int i = tid; i < npool; i += nt) tmin = min(tmin, (unsigned)(pool_k[i] >> 32));
        tmin = __reduce_min_sync(0xffffffffu, tmin);
        if (lane == 0) atomicMin(&sh.tmin, tmin);
        __syncthreads();
        const float thr = __uint_as_float(sh.tmin) + 0.7f;
        for (int i = tid; i < npool; i += nt) {
            const unsigned long long k = pool_k[i];
            const unsigned p = pool_i[i];
            if (__uint_as_float((unsigned)(k >> 32)) < thr) {
                const int pos = atomicAdd(&sh.ncur, 1);
                cur_k[pos] = k; cur_i[pos] = p;
            } else {
                const int pos = atomicAdd(&sh.npool2, 1);
                next_k[pos] = k; next_i[pos] = p;
            }
        }
        __syncthreads();
        const int ncur = sh.ncur, gbase = sh.gbase, carry = sh.npool2;
        STAT_ADD0(c_part);
        STAT_MARK();
        block_sort(cur_k, cur_i, ncur);
        __syncthreads();
        STAT_ADD0(c_sort);
        if (tid == 0) { STAT_INC(n_gen, 1); STAT_INC(n_pops, ncur); }
        STAT_MARK();
        // ---- mark the pops of this generation (and the outer pass' CHANGE flag, which nothing orders) -------
        for (int e = tid; e < ncur; e += nt) {
            const unsigned p = cur_i[e];
            V.pstate[p] = ((unsigned)(gbase + e) << 2) | 1u;
            if (OUTER) V.st[p] = (V.st[p] & ~O_MASK) | O_CHANGE;
        }
        __syncthreads();
        // ---- ownership: which INSIDE neighbours does pop e compute?  (mask of q in cur_k[e]) ---------------
        for (int e0 = 0; e0 < ncur; e0 += nt) {
            const int e = e0 + tid;
            unsigned own = 0;
            if (e < ncur) {
                const unsigned p = cur_i[e];
                const int yy = (int)(p / (unsigned)Ws), xx = (int)(p - (unsigned)yy * (unsigned)Ws);
                const unsigned G = (unsigned)(gbase + e);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int y = yy + (q == 0 ? -1 : (q == 2 ? 1 : 0)), x = xx + (q == 1 ? -1 : (q == 3 ? 1 : 0));
                    if (!mc.inb(y, x)) continue;
                    const unsigned char s = V.st[(size_t)y * Ws + x];
                    if (!(OUTER ? ((s & O_MASK) == O_INSIDE) : ((s & F_MASK) == F_INSIDE))) continue;
                    bool first = true;     // no 4-neighbour of (y,x) is popped earlier in this generation
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const int y2 = y + (r == 0 ? -1 : (r == 2 ? 1 : 0)), x2 = x + (r == 1 ? -1 : (r == 3 ? 1 : 0));
                        if (!mc.inb(y2, x2)) continue;
                        const unsigned v = V.pstate[(size_t)y2 * Ws + x2];
                        if (ps_is_pop(v) && (v >> 2) >= (unsigned)gbase && (v >> 2) < G) first = false;
                    }
                    if (first) own |= 1u << q;
                }
                cur_k[e] = own;
            }
            // exclusive scan of popc(own) over the generation, chunk by chunk
            const int c = __popc(own);
            int inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
            if (lane == 31) sh.wsum[wid] = inc;
            __syncthreads();
            int woff = 0;
            for (int w2 = 0; w2 < wid; w2++) woff += sh.wsum[w2];
            const int off = sh.ntask + woff + inc - c;
            if (e < ncur && own) {
                const unsigned p = cur_i[e];
                const int yy = (int)(p / (unsigned)Ws), xx = (int)(p - (unsigned)yy * (unsigned)Ws);
                int j = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (!(own & (1u << q))) continue;
                    const int y = yy + (q == 0 ? -1 : (q == 2 ? 1 : 0)), x = xx + (q == 1 ? -1 : (q == 3 ? 1 : 0));
                    const unsigned pn = (unsigned)y * (unsigned)Ws + (unsigned)x;
                    next_i[carry + off + j] = pn;         // key (T, J) is filled in when the task runs
                    V.pstate[pn] = ((sh.tbase + (unsigned)(off + j)) << 2) | 2u;
                    j++;
                }
            }
            __syncthreads();
            if (tid == nt - 1) sh.ntask = off + c;     // last thread holds the inclusive total of this chunk
            __syncthreads();
        }
        const int ntask = sh.ntask;
        STAT_ADD0(c_part);
        // ---- dataflow over the ordered tasks -------------------------------------------------------------
#if VSC_OUTER_LANES
        if (OUTER) outer_lane_dataflow<NW>(V, sh, Hs, Ws, ntask, carry, next_i, next_k);
#endif
        while (!(VSC_OUTER_LANES && OUTER)) {
            int j = 0;
            if (lane == 0) j = atomicAdd(&sh.next_t, 1);
            j = __shfl_sync(0xffffffffu, j, 0);
            if (j >= ntask) break;
            const unsigned pn = next_i[carry + j];
            const unsigned tbase = sh.tbase, J = tbase + (unsigned)j;
            const int y = (int)(pn / (unsigned)Ws), x = (int)(pn - (unsigned)y * (unsigned)Ws);
            STAT_MARK();
            {   // wait for every earlier task of this generation within TELEA_DC
                constexpr int DC = OUTER ? TELEA_DC_OUTER : TELEA_DC_MAIN;
                constexpr int D = 2 * DC + 1, NIT = (D * D + 31) / 32;
                unsigned* wd = sh.wdep[wid];
                unsigned depr[NIT];
                int nd = 0;
#pragma unroll
                for (int r = 0; r < NIT; r++) {
                    const int idx = lane + 32 * r;
                    const int yy = y + idx / D - DC, xx = x + idx % D - DC;
                    unsigned dep = 0xffffffffu;
                    if (idx < D * D && mc.inb(yy, xx)) {
                        const unsigned v = V.pstate[(size_t)yy * Ws + xx];      // stable during the dataflow
                        if (ps_is_task(v) && (v >> 2) >= tbase && (v >> 2) < J) dep = v >> 2;
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, dep != 0xffffffffu);
                    const int pos = nd + __popc(bal & ((1u << lane) - 1u));
                    if (dep != 0xffffffffu && pos < TELEA_MAXDEP) wd[pos] = dep;
                    nd += __popc(bal);
                    depr[r] = dep;
                }
                __syncwarp();
                STAT_INC(n_polls, 1);
                if (nd > 0 && nd <= TELEA_MAXDEP) {
                    const unsigned mine = lane < nd ? wd[lane] : 0xffffffffu;
                    const volatile unsigned* slot = &sh.ring[mine & (TELEA_RING - 1)];
                    bool pend = lane < nd;
                    while (true) {
                        if (pend) pend = *slot < mine + 1u;
                        if (!__any_sync(0xffffffffu, pend)) break;
                        // Only the oldest unfinished tasks are on the critical path: poll them eagerly; the further a
                        // claimed task is behind the completion front, the longer it sleeps (waiting warps must not
                        // steal issue slots from the working ones).
                        const int behind = j - *(volatile int*)&sh.done_t;
                        __nanosleep(behind <= 2 ? 40 : min(behind * 200, 4000));
                        STAT_INC(n_polls, 1);
                    }
                    __threadfence_block();
                } else if (nd > 0) {       // more dependencies than the list holds: every lane polls its own
                    while (true) {
                        bool pend = false;
#pragma unroll
                        for (int r = 0; r < NIT; r++)
                            if (depr[r] != 0xffffffffu) {
                                if (*(volatile unsigned*)&sh.ring[depr[r] & (TELEA_RING - 1)] >= depr[r] + 1u) depr[r] = 0xffffffffu;
                                else pend = true;
                            }
                        if (!__any_sync(0xffffffffu, pend)) break;
                        __nanosleep(100);
                    }
                    __threadfence_block();
                }
            }
            STAT_ADD(c_wait);
            STAT_MARK();
            { STAT_T0(); mc.template load<OUTER ? 1 : 4, !OUTER>(y, x); STAT_T1(c_load); }
            float dist;
            { STAT_T0(); dist = mc.min4<OUTER>(y, x); STAT_T1(c_min4); }
            __syncwarp();
            if (lane == 0) mc.set_T(y, x, dist);
            __syncwarp();
            STAT_INC(n_pix, 1);
#ifndef VSC_EXPERIMENT_NO_INPAINT
            if (!OUTER) { STAT_T0(); mc.inpaint(y, x, dist); STAT_T1(c_inp); }
#endif
            if (lane == 0) {
                const unsigned char s = mc.S(y, x);
                mc.set_S(y, x, OUTER ? ((s & ~O_MASK) | O_BAND) : ((s & ~F_MASK) | F_BAND));
                next_k[carry + j] = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)J;
                if (!OUTER && x >= keep_x0 && x < keep_x1) atomicSub(&sh.need_left, 1);
            }
            __syncwarp();
            {   // publish: results first, then the ring entry.  An entry is only overwritten once its previous occupant
                // (J - RING, claimed long ago) has finished, so "ring[J % RING] >= J + 1" always means "J is done".
                STAT_T0();
                if (lane == 0) {
                    volatile unsigned* slot = &sh.ring[J & (TELEA_RING - 1)];
                    while (J >= (unsigned)TELEA_RING && *slot < J - (unsigned)TELEA_RING + 1u) __nanosleep(20);
                    __threadfence_block();
                    *slot = J + 1u;
                    atomicAdd(&sh.done_t, 1);
                }
                __syncwarp();
                STAT_T1(c_rel);
            }
            STAT_ADD(c_pop);
        }
        __syncthreads();
        if (tid == 0) {
            sh.gbase = gbase + ncur; sh.npool = carry + ntask; sh.npool2 = 0; sh.ncur = 0; sh.ntask = 0; sh.next_t = 0; sh.done_t = 0; sh.tmin = 0xffffffffu;
            sh.tbase += (unsigned)ntask;
        }
        src ^= 1;
        __syncthreads();
        // Everything still queued has a larger T than every pixel computed so far and can therefore not influence
        // them; once all hole pixels inside the kept window are filled, the rest of the cluster is never read.
        if (!OUTER && sh.need_left <= 0) break;
    }
#ifdef VSC_TELEA_STATS
    if (tid == 0 && stats) {
        const unsigned long long tot = (unsigned long long)(clock64() - t_start);
        atomicAdd(&stats[0], sh.c_wait); atomicAdd(&stats[1], sh.c_pop); atomicAdd(&stats[2], sh.c_sort);
        atomicAdd(&stats[3], sh.c_part); atomicAdd(&stats[4], tot); atomicAdd(&stats[5], sh.n_pops);
        atomicAdd(&stats[6], sh.n_pix); atomicAdd(&stats[7], sh.n_gen); atomicAdd(&stats[8], sh.n_polls);
        atomicAdd(&stats[9], 1ull);
        if (tot > stats[10]) {   // record of the slowest cluster (racy but good enough for a profile)
            stats[10] = tot; stats[11] = sh.n_pops; stats[12] = sh.c_load; stats[13] = sh.c_min4; stats[14] = sh.c_inp; stats[15] = sh.c_rel;
        }
    }
#endif
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, 1024 / (NW * 32)) telea_cluster_kernel(const __grid_constant__ TeleaArgs a) {
    __shared__ MarchShared<NW> sh;
    __shared__ TapTable tp;
    if (threadIdx.x < 32) { tp.dk[threadIdx.x] = c_taps.dk[threadIdx.x]; tp.dl[threadIdx.x] = c_taps.dl[threadIdx.x]; tp.dst[threadIdx.x] = c_taps.dst[threadIdx.x]; }
    for (int i = threadIdx.x; i < TELEA_RING; i += blockDim.x) sh.ring[i] = 0u;
    if (threadIdx.x == 0) sh.tbase = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int v = blockIdx.x;       // view-major launch order: the first CTAs to start take each view's biggest cluster
    const TeleaView& V = a.v[v];
    const int nbig = V.fs->nbig[V.vi], ncl = nbig + V.fs->nsmall[V.vi];
    if (V.fs->qbump[V.vi] > V.qcap) {   // scratch too small: report and leave the frame to the host retry
        if (threadIdx.x == 0 && blockIdx.y == 0) atomicMax(&V.fs->overflow, V.fs->qbump[V.vi]);
        return;
    }
    Marcher mc{V, a.Hs, a.Ws, lane, &sh.win[threadIdx.x >> 5], &tp, 0, 0};
    const int cap = a.tw * a.th;
    while (true) {
        if (threadIdx.x == 0) sh.ci = atomicAdd(&V.fs->next[V.vi], 1);
        __syncthreads();
        const int i = sh.ci;
        if (i >= ncl) break;
        const int ci = i < nbig ? i : cap - 1 - (i - nbig);     // big clusters are queued first
        const int qoff = V.cl_qoff[ci], ntiles = V.cl_ntiles[ci];
        const int* tiles = V.tile_list + V.cl_toff[ci];
        if (threadIdx.x == 0) sh.need_left = V.cl_size[ci];
#ifdef VSC_EXPERIMENT_SLEEP_MARCH
        {   // experiment: hold the CTA's resources for about as long as the march would take, without issuing work
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            const unsigned long long dur = (unsigned long long)V.cl_size[ci] * VSC_EXPERIMENT_SLEEP_MARCH;
            do { __nanosleep(100000); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < dur);
            __syncthreads();
            continue;
        }
#endif
        march<true, NW>(mc, V, sh, qoff, ntiles, tiles, a.tw, V.keep_x0, V.keep_x1, a.stats ? a.stats + ((v & 1) * 2 + 0) * 16 : nullptr);
        __syncthreads();
        march<false, NW>(mc, V, sh, qoff, ntiles, tiles, a.tw, V.keep_x0, V.keep_x1, a.stats ? a.stats + ((v & 1) * 2 + 1) * 16 : nullptr);
        __syncthreads();
    }
}

}  // namespace vsc
