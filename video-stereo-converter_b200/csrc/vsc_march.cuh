// vsc_march.cuh — the march of the exact GPU Telea (cv2.inpaint(..., 3, INPAINT_TELEA),
// /root/reference/helper/stereo_core.py:457; OpenCV photo/inpaint.cpp), one CTA per cluster of holes
// (clusters: vsc_telea.cuh).  The reference pops a sorted list one pixel at a time and re-reads pixels it has just
// written; here the same result is produced in two stages:
//
//   A. ORDER.  The arrival times T and the order in which hole pixels are computed depend on the mask alone.  The two
//      fast-marching sweeps of the reference (outer distance ring, then the holes) are independent - a ring pixel and a
//      hole pixel are never 4-neighbours, the band separates them - and run as ONE march in bulk-synchronous GENERATIONS:
//        bucket g   = queue entries with floor(T / 0.7) == g.  A pixel computed while bucket g is popped has
//                     T in [popped T + 1/sqrt(2), popped T + 1], i.e. it belongs to bucket g+1 or g+2: three rotating
//                     lists replace the sorted list and bucket g is complete before it is popped
//        pop order  = (T, push order): lists are appended in push order, so a STABLE radix sort by T alone is exact
//        ownership  = a pixel is computed by its FIRST popped 4-neighbour: atomicMin of (pop rank * 4 + q)
//        task order = (owner's pop rank, q) = the sequential computation / push order; a prefix sum numbers the tasks J
//        distances  = task J sees a neighbour as known iff it is outside the domain, older than this generation, or a
//                     task J' < J of it.  Only the T of those same-generation neighbours is unknown up front; every
//                     thread re-evaluates its tasks until nothing changes - the fixed point is unique because the
//                     dependencies follow J (measured: <= 7 sweeps; the test suite pins this schedule with a sequential CPU model)
//      No queue, no per-pixel waiting: a generation is a handful of data-parallel phases.
//   B. COLOURS.  "Pixel q was known when pixel i was computed" is ord[q] < i, so the colours need no flags, no
//      generations and no CTA barriers: they are a dataflow over the recorded order.  Every hole pixel counts the earlier
//      hole pixels in its 9x9 window; pixels with count 0 are READY.  A warp only ever takes a ready pixel (window, 28
//      weights with one tap per lane, 9 sums in the reference's raster order, division / square root) and then decrements
//      the counters of the later hole pixels around it; the first one that reaches zero continues on the same warp - a
//      crack is a chain, there is no hand-over - the others go to the ready list.  See march_colour.
#pragma once
#include "vsc_telea.cuh"

namespace vsc {

constexpr float MARCH_BUCKET_INV = 1.4285714285714286f; // 1 / 0.7 (bucket width < 1/sqrt(2), see header)
constexpr int MARCH_SMEM_COUNTERS = 32768;
#ifndef VSC_MARCH_B_WARPS
#define VSC_MARCH_B_WARPS 12
#endif
constexpr int MARCH_B_WARPS = VSC_MARCH_B_WARPS;      // warps of a CTA that work on the colour stage

struct MarchWin {            // per-warp 9x9 window around the pixel being inpainted
    unsigned img[81];
    float tt[81];
    unsigned char kn[84];    // 1 = known when this pixel is computed (outside the image: the KNOWN frame)
    float taps[28 * 10];
};

template <int NW> struct MarchSh {
    union {
        unsigned hist[NW][256];                  // stage A: per-warp radix histograms
        unsigned char cnt[MARCH_SMEM_COUNTERS];  // stage B: dependency counters (bytes) of a cluster with that many tasks at most
    } u;
    unsigned tot[256];
    int wsum[32];
    int cnt[3];
    int rcnt[3];             // ring entries per bucket list (the early stop must not leave ring pixels behind)
    int need_left, ci, nband, flag;
    unsigned kmin, kmax;
    int rq_head, rq_tail, ndone;              // stage B: ready list cursors, finished tasks
#ifdef VSC_TELEA_STATS
    int st_gens, st_sweeps; long long st_sort, st_claim, st_sweep, st_push;
    unsigned long long b_ph[7], b_n, b_cont;
#endif
    TapTable tp;
    MarchWin win[NW];
};

struct MarchScratch {        // the cluster's slices of the per-view scratch arrays
    unsigned* band;          // initial band in raster order (generation 0 of both sweeps)
    unsigned* L[3];          // rotating bucket lists
    unsigned* sa;            // sort partner of the current list
    unsigned* ka; unsigned* kb;   // sort keys (ping-pong); ka doubles as the ownership masks of a generation
    unsigned* tl;            // tasks of the current generation, in order
    unsigned* seq;           // hole tasks of all generations, in order (stage B walks it)
};

template <int NW> __device__ __forceinline__ int block_excl_scan(int v, int* wsum, int& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) { const int s = wsum[w]; if (w < wid) woff += s; tot += s; }
    total = tot;
    __syncthreads();
    return woff + inc - v;
}

// One stable counting pass of an LSD radix sort on bits [shift, shift+8) of (key - kmin).  A warp owns a contiguous
// segment of the input, so equal digits keep their order without any cross-warp atomics.
template <int NW>
__device__ void radix_pass(const unsigned* ks, const unsigned* vs, unsigned* kd, unsigned* vd, int n, int shift, unsigned kmin, MarchSh<NW>& sh) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < NW * 256; i += NW * 32) (&sh.u.hist[0][0])[i] = 0u;
    __syncthreads();
    const int seg = (((n + NW - 1) / NW) + 31) & ~31;
    const int b0 = min(n, wid * seg), b1 = min(n, b0 + seg);
    for (int base = b0; base < b1; base += 32) {
        const int i = base + lane;
        const bool valid = i < b1;
        const unsigned d = valid ? (((ks[i] - kmin) >> shift) & 255u) : (256u + (unsigned)lane);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && lane == __ffs(peers) - 1) sh.u.hist[wid][d] += (unsigned)__popc(peers);
        __syncwarp();
    }
    __syncthreads();
    if (tid < 256) {       // per digit: exclusive offsets over the warps, and the digit total
        unsigned s = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) { const unsigned t = sh.u.hist[w][tid]; sh.u.hist[w][tid] = s; s += t; }
        sh.tot[tid] = s;
    }
    __syncthreads();
    if (wid == 0) {        // exclusive scan of the 256 digit totals
        unsigned loc[8], s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { loc[k] = sh.tot[lane * 8 + k]; s += loc[k]; }
        unsigned inc = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        unsigned ex = inc - s;
#pragma unroll
        for (int k = 0; k < 8; k++) { sh.tot[lane * 8 + k] = ex; ex += loc[k]; }
    }
    __syncthreads();
    for (int base = b0; base < b1; base += 32) {
        const int i = base + lane;
        const bool valid = i < b1;
        const unsigned key = valid ? ks[i] : 0u, val = valid ? vs[i] : 0u;
        const unsigned d = valid ? (((key - kmin) >> shift) & 255u) : (256u + (unsigned)lane);
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid) {
            const unsigned pos = sh.tot[d] + sh.u.hist[wid][d] + (unsigned)__popc(peers & ((1u << lane) - 1u));
            kd[pos] = key; vd[pos] = val;
        }
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) sh.u.hist[wid][d] += (unsigned)__popc(peers);
        __syncwarp();
    }
    __syncthreads();
}
// sorts (k0,v0)[0,n) by the low nbits of (key - kmin); returns 0 if the result is in (k0,v0), 1 if in (k1,v1)
template <int NW>
__device__ int radix_sort(unsigned* k0, unsigned* v0, unsigned* k1, unsigned* v1, int n, int nbits, unsigned kmin, MarchSh<NW>& sh) {
    int w = 0;
    for (int shift = 0; shift < nbits; shift += 8) {
        radix_pass<NW>(w ? k1 : k0, w ? v1 : v0, w ? k0 : k1, w ? v0 : v1, n, shift, kmin, sh);
        w ^= 1;
    }
    return w;
}

#ifdef VSC_TELEA_STATS
#define MSTAT(stmt_) do { if (threadIdx.x == 0) { stmt_; } } while (0)
#define MSTAT_ANY(stmt_) stmt_
#define MSTAT_T0() long long mst_t = clock64()
#define MSTAT_T1(f) do { if (threadIdx.x == 0) { const long long n_ = clock64(); sh.f += n_ - mst_t; mst_t = n_; } } while (0)
#else
#define MSTAT(stmt_)
#define MSTAT_ANY(stmt_)
#define MSTAT_T0()
#define MSTAT_T1(f)
#endif

// the reference's neighbour order: up, left, down, right
__device__ __forceinline__ int nb_dy(int q) { return q == 0 ? -1 : (q == 2 ? 1 : 0); }
__device__ __forceinline__ int nb_dx(int q) { return q == 1 ? -1 : (q == 3 ? 1 : 0); }

// ---- stage A: the fast march over the outer distance ring AND the holes ------------------------------------------
// The two sweeps of the reference are independent (a ring pixel and a hole pixel are never 4-neighbours, the band
// separates them), start from the same band and use the same buckets: they run as ONE march.  The order restricted to
// the hole pixels is the hole sweep's own order.  Order word (V.pstate) of a computed pixel = its task index J.
// Returns the number of hole tasks; sc.seq[0..) lists them in order.
__device__ __forceinline__ bool march_dom(unsigned char s) { return (s & F_MASK) == F_INSIDE || (s & O_MASK) == O_INSIDE; }

struct MarchEval { unsigned p; float tn[4]; bool in_[4]; float own; };
__device__ __forceinline__ void march_eval_load(const TeleaView& V, MarchEval& e, unsigned p, unsigned J, int Hs, int Ws) {
    e.p = p;
    const int y = (int)(p / (unsigned)Ws), x = (int)(p - (unsigned)y * (unsigned)Ws);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int yy = y + nb_dy(q), xx = x + nb_dx(q);
        e.tn[q] = 1.0e6f; e.in_[q] = false;           // outside the image: cv2's KNOWN frame with T = 1e6
        if (yy >= 0 && yy < Hs && xx >= 0 && xx < Ws) {
            const size_t nb = (size_t)yy * Ws + xx;
            e.tn[q] = V.tt[nb];
            const unsigned char s = V.st[nb];
            const unsigned o = V.pstate[nb];        // initialised wherever a neighbour of a domain pixel can lie
            e.in_[q] = march_dom(s) && o >= J;
        }
    }
    e.own = V.tt[p];
}
__device__ __forceinline__ float march_eval_min4(const MarchEval& e) {
    // min4's pairing: (up,left) (down,left) (up,right) (down,right)
    const float s0 = fmm_solve(e.tn[0], e.tn[1], e.in_[0], e.in_[1]), s1 = fmm_solve(e.tn[2], e.tn[1], e.in_[2], e.in_[1]);
    const float s2 = fmm_solve(e.tn[0], e.tn[3], e.in_[0], e.in_[3]), s3 = fmm_solve(e.tn[2], e.tn[3], e.in_[2], e.in_[3]);
    return fminf(fminf(s0, s1), fminf(s2, s3));
}

template <int NW>
__device__ int march_order(const TeleaView& V, MarchSh<NW>& sh, const MarchScratch& sc, int nband, int Hs, int Ws, int keep_x0, int keep_x1) {
    const int tid = threadIdx.x, nt = NW * 32, lane = tid & 31;
    if (tid == 0) { sh.cnt[0] = sh.cnt[1] = sh.cnt[2] = 0; sh.rcnt[0] = sh.rcnt[1] = sh.rcnt[2] = 0; }
    __syncthreads();
    unsigned tbase = 0;
    int mtot = 0;            // hole tasks so far
    for (int g = 0; g < (1 << 20); g++) {
        const int n = g == 0 ? nband : sh.cnt[g % 3];
        if (n == 0) {
            if (g == 0 || (sh.cnt[(g + 1) % 3] == 0 && sh.cnt[(g + 2) % 3] == 0)) break;
            continue;
        }
        unsigned* cur = g == 0 ? sc.band : sc.L[g % 3];
        const unsigned* S = cur;
        MSTAT(sh.st_gens++);
        MSTAT_T0();
        if (g > 0) {      // pop order of the bucket: stable sort by T (generation 0 is the band in raster order, all T = 0)
            if (tid == 0) { sh.kmin = 0xffffffffu; sh.kmax = 0u; }
            __syncthreads();
            unsigned kmn = 0xffffffffu, kmx = 0u;
            for (int e = tid; e < n; e += nt) {
                const unsigned k = __float_as_uint(V.tt[cur[e]]);       // T > 0: bit order = value order
                sc.ka[e] = k; kmn = min(kmn, k); kmx = max(kmx, k);
            }
            kmn = __reduce_min_sync(0xffffffffu, kmn); kmx = __reduce_max_sync(0xffffffffu, kmx);
            if (lane == 0) { atomicMin(&sh.kmin, kmn); atomicMax(&sh.kmax, kmx); }
            __syncthreads();
            const unsigned kmin = sh.kmin, span = sh.kmax - kmin;
            const int nbits = span ? 32 - __clz(span) : 0;
            if (radix_sort<NW>(sc.ka, cur, sc.kb, sc.sa, n, nbits, kmin, sh)) S = sc.sa;
        }
        MSTAT_T1(st_sort);
        // claims: the first popped neighbour (lowest pop rank, then lowest q) computes a pixel.  All loads of an entry are
        // issued before the first use (the order words are initialised wherever a neighbour of a queue entry can lie)
        for (int e = tid; e < n; e += nt) {
            const unsigned p = S[e];
            const int y = (int)(p / (unsigned)Ws), x = (int)(p - (unsigned)y * (unsigned)Ws);
            size_t nbq[4]; bool ok[4]; unsigned char sq[4]; unsigned oq[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int yy = y + nb_dy(q), xx = x + nb_dx(q);
                ok[q] = yy >= 0 && yy < Hs && xx >= 0 && xx < Ws;
                nbq[q] = ok[q] ? (size_t)yy * Ws + xx : (size_t)p;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) { sq[q] = V.st[nbq[q]]; oq[q] = V.pstate[nbq[q]]; }
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (ok[q] && march_dom(sq[q]) && oq[q] >= 0x80000000u) atomicMin(&V.pstate[nbq[q]], 0x80000000u + (unsigned)(e * 4 + q));
        }
        __syncthreads();
        // ownership (a thread takes a contiguous range of pops so that one prefix sum orders all tasks)
        const int per = (n + nt - 1) / nt, e0 = min(n, tid * per), e1 = min(n, e0 + per);
        int c = 0;
        for (int e = e0; e < e1; e++) {
            const unsigned p = S[e];
            const int y = (int)(p / (unsigned)Ws), x = (int)(p - (unsigned)y * (unsigned)Ws);
            size_t nbq[4]; bool ok[4]; unsigned char sq[4]; unsigned oq[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int yy = y + nb_dy(q), xx = x + nb_dx(q);
                ok[q] = yy >= 0 && yy < Hs && xx >= 0 && xx < Ws;
                nbq[q] = ok[q] ? (size_t)yy * Ws + xx : (size_t)p;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) { sq[q] = V.st[nbq[q]]; oq[q] = __ldcg(&V.pstate[nbq[q]]); }      // L2: sees the atomics
            unsigned own = 0;
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (ok[q] && march_dom(sq[q]) && oq[q] == 0x80000000u + (unsigned)(e * 4 + q)) own |= 1u << q;
            sc.ka[e] = own;
            c += __popc(own);
        }
        int ntask;
        const int off = block_excl_scan<NW>(c, sh.wsum, ntask);
        unsigned* TL = sc.tl;
        {
            int j = off;
            for (int e = e0; e < e1; e++) {
                const unsigned own = sc.ka[e];
                if (!own) continue;
                const unsigned p = S[e];
                const int y = (int)(p / (unsigned)Ws), x = (int)(p - (unsigned)y * (unsigned)Ws);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (!(own & (1u << q))) continue;
                    const unsigned nb = (unsigned)((y + nb_dy(q)) * Ws + (x + nb_dx(q)));
                    TL[j] = nb;
                    V.pstate[nb] = tbase + (unsigned)j;
                    j++;
                }
            }
        }
        __syncthreads();
        MSTAT_T1(st_claim);
        // distances: re-evaluate until the (unique) fixed point; two tasks per step keep more loads in flight
        while (true) {
            int ch = 0;
            MSTAT(sh.st_sweeps++);
            for (int j = tid; j < ntask; j += 2 * nt) {
                MarchEval ea, eb;
                const bool two = j + nt < ntask;
                march_eval_load(V, ea, TL[j], tbase + (unsigned)j, Hs, Ws);
                if (two) march_eval_load(V, eb, TL[j + nt], tbase + (unsigned)(j + nt), Hs, Ws);
                const float da = march_eval_min4(ea);
                if (da != ea.own) { V.tt[ea.p] = da; ch = 1; }
                if (two) {
                    const float db = march_eval_min4(eb);
                    if (db != eb.own) { V.tt[eb.p] = db; ch = 1; }
                }
            }
            if (!__syncthreads_or(ch)) break;
        }
        MSTAT_T1(st_sweep);
        // push in task order: a stable split into the next two buckets; hole tasks are also appended to seq
        const int b1i = (g + 1) % 3, b2i = (g + 2) % 3;
        const int tper = (ntask + nt - 1) / nt, j0 = min(ntask, tid * tper), j1 = min(ntask, j0 + tper);
        int c1 = 0, cm = 0, inwin = 0, r1 = 0, r2 = 0;
        for (int j = j0; j < j1; j++) {
            const unsigned p = TL[j];
            const int b = (int)floorf(__fmul_rn(V.tt[p], MARCH_BUCKET_INV));
            const bool hole = (V.st[p] & F_MASK) == F_INSIDE;
            if (b == g + 1) { c1++; r1 += hole ? 0 : 1; }
            else { r2 += hole ? 0 : 1; if (b != g + 2) sh.flag = 1; }      // cannot happen (header); makes the host fail loudly
            if (hole) { cm++; const int x = (int)(p % (unsigned)Ws); inwin += (x >= keep_x0 && x < keep_x1) ? 1 : 0; }
        }
        int total1, totalm;
        const int off1 = block_excl_scan<NW>(c1, sh.wsum, total1);
        const int offm = block_excl_scan<NW>(cm, sh.wsum, totalm);
        {
            const int base1 = sh.cnt[b1i];
            unsigned* L1 = sc.L[b1i] + base1;
            unsigned* L2 = sc.L[b2i];
            unsigned* SQ = sc.seq + mtot;
            int p1 = off1, p2 = j0 - off1, pm = offm;
            for (int j = j0; j < j1; j++) {
                const unsigned p = TL[j];
                const int b = (int)floorf(__fmul_rn(V.tt[p], MARCH_BUCKET_INV));
                if (b == g + 1) L1[p1++] = p; else L2[p2++] = p;
                if ((V.st[p] & F_MASK) == F_INSIDE) SQ[pm++] = p;
            }
            inwin = __reduce_add_sync(0xffffffffu, inwin);
            r1 = __reduce_add_sync(0xffffffffu, r1); r2 = __reduce_add_sync(0xffffffffu, r2);
            if (lane == 0) {
                if (inwin) atomicSub(&sh.need_left, inwin);
                if (r1) atomicAdd(&sh.rcnt[b1i], r1);
                if (r2) atomicAdd(&sh.rcnt[b2i], r2);
            }
            __syncthreads();
            if (tid == 0) {
                sh.cnt[b1i] = base1 + total1; sh.cnt[b2i] = ntask - total1; sh.cnt[g % 3] = 0; sh.rcnt[g % 3] = 0;
            }
        }
        tbase += (unsigned)ntask;
        mtot += totalm;
        __syncthreads();
        MSTAT_T1(st_push);
        // Everything still queued is farther from the hole boundary than every pixel computed so far and cannot influence
        // them; once all hole pixels inside the kept window are computed (and the whole ring is) the rest of the cluster
        // is never read.
        if (sh.need_left <= 0 && sh.rcnt[b1i] == 0 && sh.rcnt[b2i] == 0) break;
    }
    return mtot;
}

// ---- stage B: colours in the recorded order --------------------------------------------------------------------
// Inpainting pixel i reads the colours of the EARLIER hole pixels (order word < i) inside its 9x9 window (28 taps of the
// radius-3 disc plus their 4-neighbours for the image gradients).  Stage B runs that dependency graph as a dataflow:
//   * every task gets a counter = number of earlier hole pixels in its window; tasks with counter 0 seed the ready list
//   * a warp takes a READY task, so it never waits for another task: load the window, weights, sums, colour
//   * finishing a task decrements the counters of the later hole pixels in its window; the first one that reaches zero is
//     taken over by the same warp at once (a crack is a chain: no hand-over latency), the others go to the ready list
// Any topological order gives the reference's result, because each colour depends on its predecessors' colours only.
// Idle warps look at the ready list every few hundred nanoseconds; they are not on anybody's critical path.
// Counters are bytes in shared memory (a window has 80 other pixels) when the cluster's tasks fit, else words in global
// memory.  Colours written by other warps are read with L1-bypassing loads after the counter handshake.
constexpr unsigned MARCH_NONE = 0xffffffffu;
template <int NW>
__device__ void march_colour(const TeleaView& V, MarchSh<NW>& sh, const MarchScratch& sc, int ntask, int Hs, int Ws) {
    const int tid = threadIdx.x, nt = NW * 32, lane = tid & 31, wid = tid >> 5;
    if (ntask <= 0) return;
    MarchWin& w = sh.win[wid];
    const unsigned* seq = sc.seq;
    unsigned* const rq = sc.L[0];             // ready list: every task is appended at most once
    unsigned* const gcnt = sc.L[1];           // counters when they do not fit into shared memory
    unsigned char* const scnt = reinterpret_cast<unsigned char*>(&sh.u);
    const bool small = ntask <= MARCH_SMEM_COUNTERS;
    // ---- order words become indices into seq; ready list and counters are reset
    for (int i = tid; i < ntask; i += nt) { V.pstate[seq[i]] = (unsigned)i; rq[i] = MARCH_NONE; }
    if (tid == 0) { sh.rq_head = 0; sh.rq_tail = 0; sh.ndone = 0; }
    __syncthreads();
    // ---- counters: one warp per task
    for (int i = wid; i < ntask; i += NW) {
        const unsigned p = seq[i];
        const int y = (int)(p / (unsigned)Ws), x = (int)(p - (unsigned)y * (unsigned)Ws);
        int c = 0;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int idx = lane + 32 * r;
            bool dep = false;
            if (idx < 81 && idx != 40) {
                const int yy = y + idx / 9 - 4, xx = x + idx % 9 - 4;
                if (yy >= 0 && yy < Hs && xx >= 0 && xx < Ws) {
                    const size_t q = (size_t)yy * Ws + xx;
                    dep = (V.st[q] & F_MASK) == F_INSIDE && V.pstate[q] < (unsigned)i;
                }
            }
            c += __popc(__ballot_sync(0xffffffffu, dep));
        }
        if (lane == 0) {
            if (small) scnt[i] = (unsigned char)c; else gcnt[i] = (unsigned)c;
            if (c == 0) rq[atomicAdd(&sh.rq_tail, 1)] = (unsigned)i;
        }
    }
    __syncthreads();
#ifdef VSC_TELEA_STATS
    long long bts[7], bt_end = clock64();
#define BSTAT(k) bts[k] = clock64()
#else
#define BSTAT(k)
#endif
    unsigned next = MARCH_NONE;               // task handed over by the one I just finished
    unsigned prevp = 0, prevc = 0;            // its pixel and colour (my own store may not have reached L2 yet)
    // A cluster rarely offers more than a handful of independent chains at a time; warps beyond that would only poll
    // the ready list (and take issue slots from the frames that share the GPU).  They wait at the final barrier.
    // (wide disocclusions - clusters of tens of thousands of pixels - do offer more: every warp works on those)
    const int workers = ntask > MARCH_SMEM_COUNTERS ? NW : min(NW, max(2, min(MARCH_B_WARPS, ntask / 64)));
    while (wid < workers) {
        unsigned i = next;
        if (i == MARCH_NONE) {
            if (lane == 0) {
                const int h = atomicAdd(&sh.rq_head, 1);
                while (true) {
                    if (h < *(volatile int*)&sh.rq_tail) {
                        i = *reinterpret_cast<volatile unsigned*>(&rq[h]);
                        if (i != MARCH_NONE) break;
                    } else if (*(volatile int*)&sh.ndone >= ntask) break;
                    __nanosleep(200);
                }
            }
            i = __shfl_sync(0xffffffffu, i, 0);
            if (i == MARCH_NONE) break;
        }
        __threadfence_block();                 // counters / ready list first, then the colours they announce
        BSTAT(0);
        const unsigned p = seq[i];
        const int y = (int)(p / (unsigned)Ws), x = (int)(p - (unsigned)y * (unsigned)Ws);
        unsigned po[3];                        // order words of the hole pixels of the window (NONE elsewhere)
        {
            // all twelve loads of a lane are issued before the first use (one memory round trip); positions outside the
            // image read a clamped address and are discarded.  Order words are initialised wherever a window can reach.
            size_t qa[3]; bool inimg[3];
            unsigned char sv[3]; float tv[3]; unsigned cv[3], ov[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const int idx = min(lane + 32 * r, 80);
                const int yy = y + idx / 9 - 4, xx = x + idx % 9 - 4;
                inimg[r] = yy >= 0 && yy < Hs && xx >= 0 && xx < Ws;
                qa[r] = (size_t)min(max(yy, 0), Hs - 1) * Ws + min(max(xx, 0), Ws - 1);
            }
#pragma unroll
            for (int r = 0; r < 3; r++) {
                sv[r] = V.st[qa[r]]; tv[r] = V.tt[qa[r]]; ov[r] = V.pstate[qa[r]];
                cv[r] = __ldcg(reinterpret_cast<const unsigned*>(&V.img[qa[r]]));
            }
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const int idx = lane + 32 * r;
                po[r] = MARCH_NONE;
                if (idx < 81) {
                    unsigned c = 0; float t = 1.0e6f; unsigned char kn = 1;
                    if (inimg[r]) {
                        t = tv[r]; c = cv[r];
                        if ((sv[r] & F_MASK) == F_INSIDE) {
                            kn = ov[r] < i;
                            po[r] = ov[r];
                            if ((unsigned)qa[r] == prevp && next != MARCH_NONE) c = prevc;
                        }
                    }
                    w.kn[idx] = kn; w.tt[idx] = t; w.img[idx] = c;
                }
            }
        }
        __syncwarp();
        BSTAT(1);
#define MWI(yy, xx) (((yy) - y + 4) * 9 + ((xx) - x + 4))
        const float dist = w.tt[MWI(y, x)];
        float gtx, gty;
        {
            const bool r = w.kn[MWI(y, x + 1)], l = w.kn[MWI(y, x - 1)], d = w.kn[MWI(y + 1, x)], u = w.kn[MWI(y - 1, x)];
            if (r) gtx = l ? __fmul_rn(__fsub_rn(w.tt[MWI(y, x + 1)], w.tt[MWI(y, x - 1)]), 0.5f) : __fsub_rn(w.tt[MWI(y, x + 1)], dist);
            else gtx = l ? __fsub_rn(dist, w.tt[MWI(y, x - 1)]) : 0.f;
            if (d) gty = u ? __fmul_rn(__fsub_rn(w.tt[MWI(y + 1, x)], w.tt[MWI(y - 1, x)]), 0.5f) : __fsub_rn(w.tt[MWI(y + 1, x)], dist);
            else gty = u ? __fsub_rn(dist, w.tt[MWI(y - 1, x)]) : 0.f;
        }
        // ---- icvTeleaInpaintFMM body: one tap per lane, then sums in the reference's raster order
        bool valid = false;
        if (lane < 28) {
            const int dk = sh.tp.dk[lane], dl = sh.tp.dl[lane];
            const int ky = y + dk, kx = x + dl;
            if (ky >= 0 && ky < Hs && kx >= 0 && kx < Ws && w.kn[MWI(ky, kx)]) {
                valid = true;
                const float ry = (float)(-dk), rx = (float)(-dl);
                const float lev = (float)__drcp_rn(__dadd_rn(1.0, fabs((double)__fsub_rn(w.tt[MWI(ky, kx)], dist))));     // 1/x correctly rounded = 1.0/x
                float dir = __fadd_rn(__fmul_rn(rx, gtx), __fmul_rn(ry, gty));
                if (fabsf(dir) <= 0.01f) dir = 0.000001f;     // 0.01f is the largest float <= 0.01: same as the reference's double compare
                const float wgt = fabsf(__fmul_rn(__fmul_rn(sh.tp.dst[lane], lev), dir));
                const bool fr = w.kn[MWI(ky, kx + 1)], fl = w.kn[MWI(ky, kx - 1)], fd = w.kn[MWI(ky + 1, kx)], fu = w.kn[MWI(ky - 1, kx)];
                const int km = ky + (ky == 0), kp = ky - (ky == Hs - 1);
                const int lm = kx + (kx == 0), lp = kx - (kx == Ws - 1);
                int ixa = 0, ixb = 0, iya = 0, iyb = 0, mx = 0, my = 0;     // mx / my: 0 none, 1 one-sided, 2 central (x2)
                if (fr) { mx = fl ? 2 : 1; ixa = MWI(km, lp + 1); ixb = fl ? MWI(km, lm - 1) : MWI(km, lm); }
                else if (fl) { mx = 1; ixa = MWI(km, lp); ixb = MWI(km, lm - 1); }
                if (fd) { my = fu ? 2 : 1; iya = MWI(kp + 1, lm); iyb = fu ? MWI(km - 1, lm) : MWI(km, lm); }
                else if (fu) { my = 1; iya = MWI(kp, lm); iyb = MWI(km - 1, lm); }
                const unsigned pc = w.img[MWI(ky, kx)];
                const unsigned pxa = w.img[ixa], pxb = w.img[ixb], pya = w.img[iya], pyb = w.img[iyb];
                float* sm = w.taps + lane * 10;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const int sh8 = 8 * c;
                    float gix = 0.f, giy = 0.f;
                    if (mx) { gix = (float)((int)((pxa >> sh8) & 0xffu) - (int)((pxb >> sh8) & 0xffu)); if (mx == 2) gix = __fmul_rn(gix, 2.0f); }
                    if (my) { giy = (float)((int)((pya >> sh8) & 0xffu) - (int)((pyb >> sh8) & 0xffu)); if (my == 2) giy = __fmul_rn(giy, 2.0f); }
                    sm[c] = __fmul_rn(wgt, (float)((pc >> sh8) & 0xffu));
                    sm[3 + c] = __fmul_rn(wgt, __fmul_rn(gix, rx));
                    sm[6 + c] = __fmul_rn(wgt, __fmul_rn(giy, ry));
                }
                sm[9] = wgt;
            }
        }
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        BSTAT(2);
        // lanes 0..9 each accumulate one quantity in tap order: Ia[3], Jx[3], Jy[3], s
        // (adding +0 for an absent tap leaves the accumulator bit-identical: it is never -0)
        float acc = lane == 9 ? 1.0e-20f : 0.f;
        {
            const int col = lane < 10 ? lane : 0;
            const bool sub = lane >= 3 && lane < 9;
            float t[28];
#pragma unroll
            for (int L = 0; L < 28; L++) t[L] = w.taps[L * 10 + col];
#pragma unroll
            for (int L = 0; L < 28; L++) {
                const float tv = ((vm >> L) & 1u) ? t[L] : 0.f;
                acc = sub ? __fsub_rn(acc, tv) : __fadd_rn(acc, tv);
            }
        }
        BSTAT(3);
        const float s = __shfl_sync(0xffffffffu, acc, 9);
        const float jx = __shfl_sync(0xffffffffu, acc, min(lane + 3, 31));
        const float jy = __shfl_sync(0xffffffffu, acc, min(lane + 6, 31));
        int outc = 0;
        if (lane < 3) {
            // sat = Ia/s + (Jx+Jy)/(sqrt(Jx^2+Jy^2)+1e-20) + 0.5, rounded to the nearest integer.  Fast path: approximate it
            // in float (error < 3e-4, see below); unless that lands within 2e-3 of a rounding boundary the integer is
            // decided.  Otherwise (and for degenerate sums) the reference's float / double sequence runs.
            //   error budget at |value| <= 256: acc * rcp(s) 3 ulp (9e-5), rsqrt 2^-22 relative of <= 1.42 (7e-7), three
            //   float additions (5e-5), reference's own float rounding of Ia/s and of the result (3e-5).
            const float jn = __fadd_rn(__fmul_rn(jx, jx), __fmul_rn(jy, jy));
            bool decided = false;
            if (s > 1.0e-6f && jn > 1.0e-12f && jn < 1.0e30f && fabsf(acc) < 256.0f * s) {
                float rs, rq2;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(s));
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rq2) : "f"(jn));
                const float approx = __fadd_rn(__fadd_rn(__fmul_rn(acc, rs), __fmul_rn(__fadd_rn(jx, jy), rq2)), 0.5f);
                const float fl = floorf(approx), fr = __fsub_rn(approx, fl);      // rint(v) = floor(v + 0.5) away from the ties
                if (fabsf(__fsub_rn(fr, 0.5f)) > 2.0e-3f) {
                    outc = min(max((int)fl + (fr > 0.5f ? 1 : 0), 0), 255);
                    decided = true;
                }
            }
            if (!decided) {
                const float ia_s = __fdiv_rn(acc, s);
                const float jsum = __fadd_rn(jx, jy);
                const double den = __dadd_rn(sqrt((double)jn), (double)1.0e-20f);
                const double val = __dadd_rn(__dadd_rn((double)ia_s, __ddiv_rn((double)jsum, den)), (double)0.5f);
                const float sat = (float)val;
                outc = min(max(__float2int_rn(sat), 0), 255);
            }
        }
        const unsigned c0 = __shfl_sync(0xffffffffu, outc, 0), c1 = __shfl_sync(0xffffffffu, outc, 1),
                       c2 = __shfl_sync(0xffffffffu, outc, 2);
        const unsigned nv = (w.img[MWI(y, x)] & 0xff000000u) | c0 | (c1 << 8) | (c2 << 16);
        BSTAT(4);
#undef MWI
        if (lane == 0) {
            __stcg(reinterpret_cast<unsigned*>(&V.img[p]), nv);
            __threadfence_block();             // the colour is on its way to L2 before anybody can see a counter drop
        }
        __syncwarp();
        // ---- release the later hole pixels of the window
        unsigned ready[3];
        bool any = false;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            ready[r] = MARCH_NONE;
            const unsigned o = po[r];
            if (o != MARCH_NONE && o > i) {
                unsigned left;
                if (small) {
                    const unsigned sh8 = 8u * (o & 3u);
                    const unsigned old = atomicSub(reinterpret_cast<unsigned*>(scnt) + (o >> 2), 1u << sh8);
                    left = ((old >> sh8) & 0xffu) - 1u;
                } else left = atomicSub(&gcnt[o], 1u) - 1u;
                if (left == 0u) { ready[r] = o; any = true; }
            }
        }
        BSTAT(5);
        // the lowest ready pixel continues on this warp, the others are published
        unsigned mine = min(ready[0], min(ready[1], ready[2]));
        mine = __reduce_min_sync(0xffffffffu, mine);
        if (any) {
#pragma unroll
            for (int r = 0; r < 3; r++)
                if (ready[r] != MARCH_NONE && ready[r] != mine) {
                    const int t = atomicAdd(&sh.rq_tail, 1);
                    *reinterpret_cast<volatile unsigned*>(&rq[t]) = ready[r];
                }
        }
        if (lane == 0) atomicAdd(&sh.ndone, 1);
        next = mine; prevp = p; prevc = nv;
        __syncwarp();
        BSTAT(6);
#ifdef VSC_TELEA_STATS
        if (lane == 0) {
            atomicAdd(&sh.b_n, 1ull); if (mine != MARCH_NONE) atomicAdd(&sh.b_cont, 1ull);
            for (int k = 0; k < 6; k++) atomicAdd(&sh.b_ph[k], (unsigned long long)(bts[k + 1] - bts[k]));
            atomicAdd(&sh.b_ph[6], (unsigned long long)(bts[0] - bt_end));
            bt_end = bts[6];
        }
#endif
    }
    __syncthreads();
}

// One CTA per cluster (persistent CTAs pull clusters, big ones first).
template <int NW>
__global__ void __launch_bounds__(NW * 32, NW <= 16 ? 2 : 1) telea_march_kernel(const __grid_constant__ TeleaArgs a) {
    extern __shared__ __align__(16) unsigned char march_smem[];
    MarchSh<NW>& sh = *reinterpret_cast<MarchSh<NW>*>(march_smem);
    const int tid = threadIdx.x, nt = NW * 32, lane = tid & 31, wid = tid >> 5;
    if (tid < 32) { sh.tp.dk[tid] = c_taps.dk[tid]; sh.tp.dl[tid] = c_taps.dl[tid]; sh.tp.dst[tid] = c_taps.dst[tid]; }
    if (tid == 0) sh.flag = 0;
    const int v = blockIdx.x;       // view-major launch order: the first CTAs to start take each view's biggest cluster
    const TeleaView& V = a.v[v];
    const int nbig = V.fs->nbig[V.vi], ncl = nbig + V.fs->nsmall[V.vi];
    if (V.fs->qbump[V.vi] > V.qcap) {   // scratch too small: report and leave the frame to the host retry
        if (tid == 0 && blockIdx.y == 0) atomicMax(&V.fs->overflow, V.fs->qbump[V.vi]);
        return;
    }
    const int Hs = a.Hs, Ws = a.Ws, cap = a.tw * a.th;
    unsigned* const scr0 = reinterpret_cast<unsigned*>(V.qkey[0]);     // 6 qcap-sized u32 arrays
    unsigned* const scr1 = V.qidx[0];                                  // 3 more
    __syncthreads();
    while (true) {
        if (tid == 0) sh.ci = atomicAdd(&V.fs->next[V.vi], 1);
        __syncthreads();
        const int i = sh.ci;
        if (i >= ncl) break;
        const int ci = i < nbig ? i : cap - 1 - (i - nbig);     // big clusters are queued first
        const int qoff = V.cl_qoff[ci], ntiles = V.cl_ntiles[ci];
        const int* tiles = V.tile_list + V.cl_toff[ci];
        MarchScratch sc;
        sc.band = scr0 + qoff; sc.L[0] = scr0 + (size_t)V.qcap + qoff; sc.L[1] = scr0 + 2 * (size_t)V.qcap + qoff;
        sc.L[2] = scr0 + 3 * (size_t)V.qcap + qoff; sc.sa = scr0 + 4 * (size_t)V.qcap + qoff; sc.ka = scr0 + 5 * (size_t)V.qcap + qoff;
        sc.kb = scr1 + qoff; sc.tl = scr1 + (size_t)V.qcap + qoff; sc.seq = scr1 + 2 * (size_t)V.qcap + qoff;
        if (tid == 0) { sh.need_left = V.cl_size[ci]; sh.nband = 0; sh.kmin = 0xffffffffu; sh.kmax = 0u; }
#ifdef VSC_TELEA_STATS
        long long ck[5];
        if (tid == 0) { sh.st_gens = sh.st_sweeps = 0; sh.st_sort = sh.st_claim = sh.st_sweep = sh.st_push = 0; ck[0] = clock64();
                        for (int k = 0; k < 7; k++) sh.b_ph[k] = 0; sh.b_n = sh.b_cont = 0; }
#endif
        __syncthreads();
        // the band (initial queue of both sweeps): collect, then order by raster position
        for (int ti = wid; ti < ntiles; ti += NW) {
            const int t = tiles[ti];
            const int ty = t / a.tw, tx = t - ty * a.tw;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int y = ty * TG + (lane >> 3) + 4 * h, x = tx * TG + (lane & 7);
                bool isb = false;
                unsigned p = 0;
                if (y < Hs && x < Ws) { p = (unsigned)y * (unsigned)Ws + (unsigned)x; isb = (V.st[p] & ST_BAND0) != 0; }
                const unsigned bm = __ballot_sync(0xffffffffu, isb);
                int base = 0;
                if (lane == 0 && bm) base = atomicAdd(&sh.nband, __popc(bm));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (isb) {
                    const int pos = base + __popc(bm & ((1u << lane) - 1));
                    sc.ka[pos] = p; sc.sa[pos] = p;
                }
                const unsigned pmn = __reduce_min_sync(0xffffffffu, isb ? p : 0xffffffffu), pmx = __reduce_max_sync(0xffffffffu, isb ? p : 0u);
                if (lane == 0 && bm) { atomicMin(&sh.kmin, pmn); atomicMax(&sh.kmax, pmx); }
            }
        }
        __syncthreads();
        const int nband = sh.nband;
        {
            const unsigned kmin = sh.kmin, span = nband ? sh.kmax - kmin : 0u;
            const int nbits = span ? 32 - __clz(span) : 0;
            if (!radix_sort<NW>(sc.ka, sc.sa, sc.kb, sc.band, nband, nbits, kmin, sh)) {
                for (int e = tid; e < nband; e += nt) sc.band[e] = sc.sa[e];
                __syncthreads();
            }
        }
        MSTAT(ck[1] = clock64());
        const int ntask = march_order<NW>(V, sh, sc, nband, Hs, Ws, V.keep_x0, V.keep_x1);
        MSTAT(ck[2] = clock64());
        // icvCalcFMM(..., negate = true): the popped pixels of the outer sweep (band and ring) get T = -T
        for (int ti = wid; ti < ntiles; ti += NW) {
            const int t = tiles[ti];
            const int ty = t / a.tw, tx = t - ty * a.tw;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int y = ty * TG + (lane >> 3) + 4 * h, x = tx * TG + (lane & 7);
                if (y < Hs && x < Ws) {
                    const size_t p = (size_t)y * Ws + x;
                    const unsigned char s = V.st[p];
                    if ((s & ST_BAND0) || (s & O_MASK) == O_INSIDE) V.tt[p] = -V.tt[p];
                }
            }
        }
        __syncthreads();
        MSTAT(ck[3] = clock64());
        march_colour<NW>(V, sh, sc, ntask, Hs, Ws);
#ifdef VSC_TELEA_STATS
        if (tid == 0 && a.stats) {
            ck[4] = clock64();
            unsigned long long* st = a.stats + (v & 1) * 32;
            const unsigned long long tot = (unsigned long long)(ck[4] - ck[0]);
            atomicAdd(&st[0], tot); atomicAdd(&st[1], 1ull);
            atomicAdd(&st[2], (unsigned long long)(ck[1] - ck[0])); atomicAdd(&st[3], (unsigned long long)(ck[2] - ck[1]));
            atomicAdd(&st[4], (unsigned long long)(ck[3] - ck[2])); atomicAdd(&st[5], (unsigned long long)(ck[4] - ck[3]));
            atomicAdd(&st[6], (unsigned long long)ntask);
            if (tot > st[8]) {      // the slowest cluster (racy, good enough for a profile)
                st[8] = tot; st[9] = ck[1] - ck[0]; st[10] = ck[2] - ck[1]; st[11] = ck[3] - ck[2]; st[12] = ck[4] - ck[3];
                st[13] = ntask; st[14] = nband; st[15] = sh.st_gens; st[16] = 0; st[17] = sh.st_sweeps;
                st[18] = sh.st_sort; st[19] = sh.st_claim; st[20] = sh.st_sweep; st[21] = sh.st_push;
                for (int k = 0; k < 7; k++) st[22 + k] = sh.b_ph[k];
                st[29] = sh.b_n; st[30] = sh.b_cont;
            }
        }
#endif
        if (tid == 0 && sh.flag) atomicMax(&V.fs->overflow, 0x7fffffff);    // broken bucket invariant: make the host fail loudly
        __syncthreads();
    }
}

}  // namespace vsc
