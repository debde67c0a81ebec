/*
 * march_model.c — CPU model of the GPU hole-filling schedule.  TEST INFRASTRUCTURE ONLY (see vsc_oracle.c).
 *
 * cv2.inpaint(..., INPAINT_TELEA) (/root/reference/helper/stereo_core.py:457) pops a sorted list
 * one pixel at a time.  The CUDA march (video-stereo-converter_b200/csrc/vsc_march.cuh) computes the same arrival
 * times T and the same computation order of the hole pixels in bulk-synchronous GENERATIONS:
 *
 *   bucket g   = queue entries with floor(T / 0.7) == g.  A pixel computed while bucket g is popped has
 *                T >= (popped T) + 1/sqrt(2) and T < (popped T) + 1 (+ rounding), i.e. it lands in bucket g+1 or g+2:
 *                three rotating lists replace the sorted list, and bucket g is complete before it is popped.
 *   pop order  = inside a bucket: by (T, push order).  Entries are appended in push order, so a STABLE sort by T
 *                alone gives the reference's pop order (FIFO among equal T).
 *   ownership  = a pixel is computed when the FIRST of its 4-neighbours is popped: the minimum over its popped
 *                neighbours of (pop rank * 4 + neighbour index q).
 *   task order = (owner's pop rank, q): the order in which the sequential algorithm computes (and pushes) the
 *                pixels; a prefix sum over the owners gives every task its global index J.
 *   distances  = task J sees a 4-neighbour as known iff it is outside the marched domain, was computed in an
 *                earlier generation, or is a task J' < J of this generation.  Only the T of those same-generation
 *                neighbours is not available up front; the (unique, because the dependencies follow J) fixed point
 *                is reached by re-evaluating all tasks until nothing changes ("sweeps").
 *
 *   one march   = the outer-ring sweep and the hole sweep are independent (a ring pixel and a hole pixel are never
 *                4-neighbours: the band separates them), start from the same band and use the same buckets, so the
 *                GPU runs them as ONE march over the union of both domains.  The order restricted to the hole pixels
 *                is the hole sweep's own order; ring times are negated afterwards (icvCalcFMM's negate).
 *
 * This file restates that schedule sequentially, so that tests can pin it against the one-pop-at-a-time oracle
 * (orc_telea_u8c3 / orc_telea_u8c3_two_pass): identical T for every pixel, identical order for every hole pixel.
 * It also reports the shape of the work (generations, bucket sizes, sweeps) for DESIGN.md.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

#define BUCKET_W_INV 1.4285714285714286f /* 1 / 0.7 */

static float solve2(float t1, float t2, int in1, int in2) { /* FastMarching_solve, see vsc_oracle.c:fmm_solve */
    double sol, a11 = t1, a22 = t2;
    double m12 = a11 < a22 ? a11 : a22;
    if (!in1) {
        if (!in2) {
            if (fabs(a11 - a22) >= 1.0) sol = 1 + m12;
            else sol = (a11 + a22 + sqrt((double)(2 - (a11 - a22) * (a11 - a22)))) * 0.5;
        } else sol = 1 + a11;
    } else if (!in2) sol = 1 + a22;
    else sol = 1 + m12;
    return (float)sol;
}

typedef struct { uint32_t *p; size_t n, cap; } list_t;
static void lpush(list_t *l, uint32_t v) {
    if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 1024; l->p = realloc(l->p, l->cap * sizeof(uint32_t)); }
    l->p[l->n++] = v;
}

typedef struct { float T; uint32_t pos; uint32_t p; } sent;
static int sent_cmp(const void *a, const void *b) {
    const sent *x = a, *y = b;
    if (x->T < y->T) return -1;
    if (x->T > y->T) return 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}

/* stats layout (int64): 0 generations, 1 tasks, 2 largest bucket, 3 sweeps in total, 4 largest sweep count of one
 * generation, 5 buckets that needed a sort (more than one distinct T), 6 largest number of distinct T in a bucket,
 * 7 task evaluations in total (sum over sweeps of the tasks evaluated) */
#define NSTAT 8

/* One fast-marching pass over the domain `dom` (1 = still to compute).  C = row pitch, the 1-pixel frame is
 * outside the domain and never popped.  band: initial queue in raster order (T = 0).  ord (optional): receives the
 * computation index of every domain pixel.  Returns the number of tasks. */
static size_t march_pass(const uint8_t *dom, const uint8_t *hole, float *t, int R, int C, const list_t *band, int32_t *ord, int64_t *stats) {
    const size_t N = (size_t)R * C;
    uint32_t *ow = malloc(N * sizeof(uint32_t)); /* 0xffffffff: not computed; < 2^31: task index J; else a claim */
    for (size_t k = 0; k < N; k++) ow[k] = 0xffffffffu;
    list_t B[3] = {{0}, {0}, {0}};
    for (size_t k = 0; k < band->n; k++) lpush(&B[0], band->p[k]);
    const int dq[4] = {-C, -1, C, 1}; /* up, left, down, right: the reference's neighbour order */
    size_t tbase = 0, nhole = 0;
    sent *srt = NULL; size_t srt_cap = 0;
    uint32_t *tl = NULL; size_t tl_cap = 0;
    float *tprev = NULL; size_t tprev_cap = 0;
    for (int g = 0;; g++) {
        list_t *cur = &B[g % 3];
        if (cur->n == 0) {
            if (B[(g + 1) % 3].n == 0 && B[(g + 2) % 3].n == 0) break;
            continue;
        }
        const size_t n = cur->n;
        stats[0]++;
        if ((int64_t)n > stats[2]) stats[2] = (int64_t)n;
        /* stable sort by T */
        if (n > srt_cap) { srt_cap = n * 2; srt = realloc(srt, srt_cap * sizeof(sent)); }
        for (size_t e = 0; e < n; e++) { srt[e].T = t[cur->p[e]]; srt[e].pos = (uint32_t)e; srt[e].p = cur->p[e]; }
        qsort(srt, n, sizeof(sent), sent_cmp);
        {
            int64_t distinct = 1;
            for (size_t e = 1; e < n; e++) if (srt[e].T != srt[e - 1].T) distinct++;
            if (distinct > 1) stats[5]++;
            if (distinct > stats[6]) stats[6] = distinct;
        }
        /* claims: first popped neighbour wins */
        for (size_t e = 0; e < n; e++)
            for (int q = 0; q < 4; q++) {
                const size_t nb = srt[e].p + dq[q];
                if (!dom[nb] || ow[nb] < 0x80000000u) continue;
                const uint32_t claim = 0x80000000u + (uint32_t)(e * 4 + q);
                if (claim < ow[nb]) ow[nb] = claim;
            }
        /* tasks in (owner rank, q) order */
        size_t ntask = 0;
        for (size_t e = 0; e < n; e++)
            for (int q = 0; q < 4; q++) {
                const size_t nb = srt[e].p + dq[q];
                if (!dom[nb] || ow[nb] != 0x80000000u + (uint32_t)(e * 4 + q)) continue;
                if (ntask == tl_cap) { tl_cap = tl_cap ? tl_cap * 2 : 1024; tl = realloc(tl, tl_cap * sizeof(uint32_t)); }
                tl[ntask] = (uint32_t)nb;
                ow[nb] = (uint32_t)(tbase + ntask);
                ntask++;
            }
        /* distances: sweeps until the fixed point.  Every sweep reads the previous sweep's values only (Jacobi), the
         * least favourable schedule a parallel machine can produce; in-place updates converge at least as fast */
        int64_t sweeps = 0;
        if (ntask > tprev_cap) { tprev_cap = ntask * 2; tprev = realloc(tprev, tprev_cap * sizeof(float)); }
        for (;;) {
            int changed = 0;
            sweeps++;
            for (size_t j = 0; j < ntask; j++) tprev[j] = t[tl[j]];
            for (size_t j = 0; j < ntask; j++) {
                const size_t p = tl[j];
                const uint32_t J = (uint32_t)(tbase + j);
                float tn[4]; int in_[4];
                for (int q = 0; q < 4; q++) {
                    const size_t nb = p + dq[q];
                    in_[q] = dom[nb] && ow[nb] >= J; /* not computed yet at the time task J runs */
                    tn[q] = (dom[nb] && ow[nb] >= tbase && ow[nb] < J) ? tprev[ow[nb] - tbase] : t[nb];
                }
                /* min4's pairing: (up,left) (down,left) (up,right) (down,right) */
                const float s0 = solve2(tn[0], tn[1], in_[0], in_[1]), s1 = solve2(tn[2], tn[1], in_[2], in_[1]);
                const float s2 = solve2(tn[0], tn[3], in_[0], in_[3]), s3 = solve2(tn[2], tn[3], in_[2], in_[3]);
                float a = s0 < s1 ? s0 : s1, b = s2 < s3 ? s2 : s3;
                const float d = a < b ? a : b;
                if (d != tprev[j]) { t[p] = d; changed = 1; }
            }
            stats[7] += (int64_t)ntask;
            if (!changed) break;
        }
        stats[3] += sweeps;
        if (sweeps > stats[4]) stats[4] = sweeps;
        /* push in task order */
        B[(g + 2) % 3].n = 0;
        for (size_t j = 0; j < ntask; j++) {
            const uint32_t p = tl[j];
            const int b = (int)floorf(t[p] * BUCKET_W_INV);
            if (b != g + 1 && b != g + 2) { stats[0] = -1000000 - g; goto out; } /* the bucket argument failed */
            lpush(&B[b % 3], p);
            if (ord && hole[p]) ord[p] = (int32_t)nhole++;   /* index among the hole pixels, in computation order */
        }
        tbase += ntask;
        cur->n = 0;
    }
out:
    stats[1] += (int64_t)tbase;
    free(ow); free(srt); free(tl); free(tprev);
    for (int k = 0; k < 3; k++) free(B[k].p);
    return tbase;
}

/* mask [H,W] nonzero = hole (already dilated).  t_out [(H+2)*(W+2)] as orc_telea_u8c3's t_out; ord_out
 * [(H+2)*(W+2)] as orc_telea_u8c3_two_pass's ord_out (-1 outside the mask); stats [NSTAT] */
ORC_API void orc_march_model(const uint8_t *mask, int H, int W, int radius, float *t_out, int32_t *ord_out, int64_t *stats) {
    const int R = H + 2, C = W + 2;
    const int range = radius < 1 ? 1 : (radius > 100 ? 100 : radius);
    const size_t N = (size_t)R * C;
    uint8_t *f = calloc(N, 1), *o = calloc(N, 1), *bnd = calloc(N, 1);
    float *t = malloc(N * sizeof(float));
    memset(stats, 0, NSTAT * sizeof(int64_t));
    for (size_t k = 0; k < N; k++) { t[k] = 1.0e6f; if (ord_out) ord_out[k] = -1; }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            if (mask[(size_t)y * W + x]) f[(size_t)(y + 1) * C + x + 1] = 1;
    list_t band = {0};
    for (int i = 1; i < R - 1; i++)
        for (int j = 1; j < C - 1; j++) {
            const size_t p = (size_t)i * C + j;
            if (f[p]) continue;
            if (f[p - C] || f[p + C] || f[p - 1] || f[p + 1]) { bnd[p] = 1; t[p] = 0.f; lpush(&band, (uint32_t)p); }
        }
    for (int i = 1; i < R - 1; i++)
        for (int j = 1; j < C - 1; j++) {
            const size_t p = (size_t)i * C + j;
            if (f[p] || bnd[p]) continue;
            int hit = 0;
            for (int di = -range; di <= range && !hit; di++) {
                const int ii = i + di; if (ii < 0 || ii >= R) continue;
                for (int dj = -range; dj <= range; dj++) {
                    const int jj = j + dj; if (jj < 0 || jj >= C) continue;
                    if (f[(size_t)ii * C + jj]) { hit = 1; break; }
                }
            }
            if (hit) o[p] = 1;
        }
    if (band.n) {
        uint8_t *dom = malloc(N);
        for (size_t k = 0; k < N; k++) dom[k] = (uint8_t)(o[k] | f[k]);
        if (ord_out) for (size_t k = 0; k < N; k++) if (f[k]) ord_out[k] = INT32_MAX;
        march_pass(dom, f, t, R, C, &band, ord_out, stats);
        /* icvCalcFMM(..., negate = true): every popped pixel of the outer sweep (band and ring) gets t = -t */
        for (size_t k = 0; k < N; k++) if (bnd[k] || (o[k] && t[k] != 1.0e6f)) t[k] = -t[k];
        free(dom);
    }
    if (t_out) memcpy(t_out, t, N * sizeof(float));
    free(f); free(o); free(bnd); free(t); free(band.p);
}
