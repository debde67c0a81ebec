/*
 * vsc_oracle.c — CPU restatement of the reference's SBS hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker / reported baseline.  The product path
 * (video-stereo-converter_b200/) never links, imports or calls anything in oracle/.
 *
 * What it restates (all citations into /root/reference):
 *   helper/stereo_core.py:225-311  StereoGenerator.process_frame (composition is in oracle.py)
 *   helper/stereo_core.py:71-88    normalize_depth            -> orc_normalize
 *   helper/stereo_core.py:348-366  _depth_upsampling / :262   -> orc_bilinear_up
 *   helper/stereo_core.py:368-385  _soft_depth_edges          -> orc_gauss_blur
 *   helper/stereo_core.py:91-107   apply_depth_gamma          -> orc_gamma
 *   helper/stereo_core.py:110-190  forward_warp_stereo        -> orc_warp (sort-free form)
 *   helper/stereo_core.py:387-412  _smooth_warping_artifacts  -> orc_bilateral_u8c3
 *   helper/stereo_core.py:436-457  _inpaint_missing_regions   -> orc_dilate3 + orc_telea_u8c3
 *   helper/stereo_core.py:414-434  _sharpen_image             -> orc_sharpen
 *   helper/stereo_core.py:298-299  F.interpolate(mode='area') -> orc_area_pool
 *   helper/stereo_core.py:253-254  cv2.resize(INTER_LANCZOS4) -> orc_lanczos4_h_*
 *   depth_map_generator.py:217-236 depth post-processing      -> orc_depth_post
 *
 * Third-party arithmetic that is NOT under /root/reference is restated from the libraries'
 * published algorithms: OpenCV 4.13.0 (unpinned in requirements.txt:2)
 * resize/bilateralFilter/dilate/inpaint, kornia gaussian_blur2d (unpinned, absent), PyTorch
 * 2.11 interpolate/pow/scatter_.  Each restatement is pinned in tests/ against the library
 * itself (cv2 / torch, both present in the image) and against golden vectors produced by the
 * unmodified reference (tests/golden/, generator: oracle/make_golden.py).
 *
 * Float determinism: this file is compiled with -ffp-contract=off; every fused multiply-add is
 * an explicit fmaf()/fma().  The CUDA kernels follow the same operation order so that the
 * product path is bit-identical to this oracle; the (small) distance between this oracle and
 * the reference's backend-defined float summation orders is what the golden tests quantify.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * cv2.resize(src,(dw,H),INTER_LANCZOS4), height unchanged (stereo_core.py:253-254, :65).
 * OpenCV imgproc/resize.cpp: interpolateLanczos4 + HResizeLanczos4 + VResizeLanczos4.
 * ---------------------------------------------------------------------------------------- */
static void lanczos4_coeffs(float x, float *c) {
    static const double s45 = 0.70710678118654752440084436210485;
    static const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45},
                                    {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    const double PI = 3.1415926535897932384626433832795;
    float sum = 0.f;
    double y0 = -(x + 3) * PI * 0.25, s0 = sin(y0), c0 = cos(y0);
    for (int i = 0; i < 8; i++) {
        float y0_ = (x + 3 - i);
        if (fabsf(y0_) >= 1e-6f) {
            double y = -y0_ * PI * 0.25;
            c[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
        } else {
            c[i] = 1e30f;
        }
        sum += c[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < 8; i++) c[i] *= sum;
}

/* per destination column: leftmost tap source column (sx-3) and the 8 float taps */
static void lanczos4_table(int W, int dW, int *sx0, float *taps) {
    double inv_scale = (double)dW / (double)W;
    double scale = 1.0 / inv_scale;
    for (int dx = 0; dx < dW; dx++) {
        float fx = (float)((dx + 0.5) * scale - 0.5);
        int sx = (int)floorf(fx);
        fx -= (float)sx;
        sx0[dx] = sx - 3;
        lanczos4_coeffs(fx, taps + 8 * dx);
    }
}

static inline short sat_short(float v) {
    long r = lrintf(v);
    return (short)(r < -32768 ? -32768 : (r > 32767 ? 32767 : r));
}

/* exported so the CUDA host side can be compared tap-for-tap in tests */
ORC_API void orc_lanczos4_table(int W, int dW, int *sx0, float *taps, short *itaps) {
    lanczos4_table(W, dW, sx0, taps);
    if (itaps)
        for (int i = 0; i < 8 * dW; i++) itaps[i] = sat_short(taps[i] * 2048.f);
}

/* the vertical pass at scale 1: identity tap (index 3) after normalisation, others ~1e-31 */
static void lanczos4_vtaps(float *beta, short *ibeta) {
    lanczos4_coeffs(0.f, beta);
    for (int i = 0; i < 8; i++) ibeta[i] = sat_short(beta[i] * 2048.f);
}

ORC_API void orc_lanczos4_h_u8(const uint8_t *src, int H, int W, int C, int dW, uint8_t *dst) {
    int *sx0 = (int *)malloc(sizeof(int) * dW);
    float *taps = (float *)malloc(sizeof(float) * 8 * dW);
    short *it = (short *)malloc(sizeof(short) * 8 * dW);
    lanczos4_table(W, dW, sx0, taps);
    for (int i = 0; i < 8 * dW; i++) it[i] = sat_short(taps[i] * 2048.f);
    float beta[8]; short ib[8];
    lanczos4_vtaps(beta, ib);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        const uint8_t *s = src + (size_t)y * W * C;
        uint8_t *d = dst + (size_t)y * dW * C;
        for (int dx = 0; dx < dW; dx++)
            for (int c = 0; c < C; c++) {
                int h = 0;
                for (int k = 0; k < 8; k++) {
                    int sx = clampi(sx0[dx] + k, 0, W - 1);
                    h += (int)s[sx * C + c] * (int)it[8 * dx + k];
                }
                /* vertical pass: rows y-3..y+4 (clamped) with ibeta; ibeta = {0,0,0,2048,0..} */
                int v = h * (int)ib[3];
                int r = (v + (1 << 21)) >> 22;
                d[dx * C + c] = (uint8_t)clampi(r, 0, 255);
            }
    }
    free(sx0); free(taps); free(it);
}

/* u16 and f32: float taps, accumulate in tap order, no FMA; vertical pass multiplies by beta[3] */
ORC_API void orc_lanczos4_h_u16(const uint16_t *src, int H, int W, int dW, uint16_t *dst) {
    int *sx0 = (int *)malloc(sizeof(int) * dW);
    float *taps = (float *)malloc(sizeof(float) * 8 * dW);
    lanczos4_table(W, dW, sx0, taps);
    float beta[8]; short ib[8];
    lanczos4_vtaps(beta, ib);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        const uint16_t *s = src + (size_t)y * W;
        uint16_t *d = dst + (size_t)y * dW;
        for (int dx = 0; dx < dW; dx++) {
            float v = 0.f;
            for (int k = 0; k < 8; k++) {
                int sx = clampi(sx0[dx] + k, 0, W - 1);
                v += (float)s[sx] * taps[8 * dx + k];
            }
            v = v * beta[3];
            long r = lrintf(v);
            d[dx] = (uint16_t)(r < 0 ? 0 : (r > 65535 ? 65535 : r));
        }
    }
    free(sx0); free(taps);
}

ORC_API void orc_lanczos4_h_f32(const float *src, int H, int W, int dW, float *dst) {
    int *sx0 = (int *)malloc(sizeof(int) * dW);
    float *taps = (float *)malloc(sizeof(float) * 8 * dW);
    lanczos4_table(W, dW, sx0, taps);
    float beta[8]; short ib[8];
    lanczos4_vtaps(beta, ib);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        const float *s = src + (size_t)y * W;
        float *d = dst + (size_t)y * dW;
        for (int dx = 0; dx < dW; dx++) {
            float v = 0.f;
            for (int k = 0; k < 8; k++) {
                int sx = clampi(sx0[dx] + k, 0, W - 1);
                v += s[sx] * taps[8 * dx + k];
            }
            d[dx] = v * beta[3];
        }
    }
    free(sx0); free(taps);
}

/* ------------------------------------------------------------------------------------------
 * normalize_depth (stereo_core.py:71-88): (d-min)/(max-min), zeros if range < 1e-6.
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_normalize(const float *d, size_t n, float *out, float *mn_out, float *mx_out) {
    float mn = d[0], mx = d[0];
    for (size_t i = 1; i < n; i++) { if (d[i] < mn) mn = d[i]; if (d[i] > mx) mx = d[i]; }
    if (mn_out) *mn_out = mn;
    if (mx_out) *mx_out = mx;
    float range = mx - mn;
    if (range < 1e-6f) { memset(out, 0, n * sizeof(float)); return; }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) out[i] = (d[i] - mn) / range;
}

/* ------------------------------------------------------------------------------------------
 * F.interpolate(..., mode='bilinear', align_corners=False) (stereo_core.py:262, :366).
 * PyTorch ATen UpSampleKernel: scale = in/out (f32); src = max(scale*(o+0.5)-0.5, 0);
 * i0 = min(floor(src), in-1), i1 = min(i0+1, in-1), l1 = clamp(src-i0,0,1), l0 = 1-l1;
 * out = l0y*(l0x*a00 + l1x*a01) + l1y*(l0x*a10 + l1x*a11).  The ATen CPU kernel is compiled with
 * FMA contraction; the contraction pattern below was identified empirically against torch
 * 2.11 (0 mismatching floats, see tests/test_oracle_vs_libs.py):
 *   src = fmaf(scale, o+0.5, -0.5);  row(r) = fmaf(l0x, r[x0], l1x*r[x1]);
 *   out = fmaf(l0y, row(r0), l1y*row(r1)).
 * ---------------------------------------------------------------------------------------- */
static void bilinear_axis(int in, int out, int *i0, int *i1, float *l0, float *l1) {
    float scale = (float)in / (float)out;
    for (int o = 0; o < out; o++) {
        float s = fmaf(scale, (float)o + 0.5f, -0.5f);
        if (s < 0.f) s = 0.f;
        int a = (int)floorf(s);
        if (a > in - 1) a = in - 1;
        float lam = s - (float)a;
        if (lam < 0.f) lam = 0.f;
        if (lam > 1.f) lam = 1.f;
        i0[o] = a;
        i1[o] = a + 1 < in ? a + 1 : in - 1;
        l1[o] = lam;
        l0[o] = 1.f - lam;
    }
}

ORC_API void orc_bilinear_axis(int in, int out, int *i0, int *i1, float *l0, float *l1) {
    bilinear_axis(in, out, i0, i1, l0, l1);
}

/* planar: src [C,H,W] -> dst [C,oH,oW] */
ORC_API void orc_bilinear_up(const float *src, int C, int H, int W, int oH, int oW, float *dst) {
    int *y0 = malloc(sizeof(int) * oH), *y1 = malloc(sizeof(int) * oH);
    int *x0 = malloc(sizeof(int) * oW), *x1 = malloc(sizeof(int) * oW);
    float *ly0 = malloc(sizeof(float) * oH), *ly1 = malloc(sizeof(float) * oH);
    float *lx0 = malloc(sizeof(float) * oW), *lx1 = malloc(sizeof(float) * oW);
    bilinear_axis(H, oH, y0, y1, ly0, ly1);
    bilinear_axis(W, oW, x0, x1, lx0, lx1);
    for (int c = 0; c < C; c++) {
        const float *s = src + (size_t)c * H * W;
        float *d = dst + (size_t)c * oH * oW;
#pragma omp parallel for schedule(static)
        for (int oy = 0; oy < oH; oy++) {
            const float *r0 = s + (size_t)y0[oy] * W, *r1 = s + (size_t)y1[oy] * W;
            float *o = d + (size_t)oy * oW;
            for (int ox = 0; ox < oW; ox++) {
                float top = fmaf(lx0[ox], r0[x0[ox]], lx1[ox] * r0[x1[ox]]);
                float bot = fmaf(lx0[ox], r1[x0[ox]], lx1[ox] * r1[x1[ox]]);
                o[ox] = fmaf(ly0[oy], top, ly1[oy] * bot);
            }
        }
    }
    free(y0); free(y1); free(x0); free(x1); free(ly0); free(ly1); free(lx0); free(lx1);
}

/* ------------------------------------------------------------------------------------------
 * kornia gaussian_blur2d (stereo_core.py:385, :432): normalised taps, reflect border,
 * horizontal pass then vertical pass.  Accumulation: acc = fmaf(g[k], x[k], acc), k ascending,
 * acc starting from 0 (the oracle's definition of the backend-defined conv2d order).
 * ---------------------------------------------------------------------------------------- */
static inline int reflect_idx(int i, int n) { /* 'reflect' (no edge repeat), valid for pad < n */
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

/* planar [C,H,W]; g has k taps */
ORC_API void orc_gauss_blur(const float *src, int C, int H, int W, int k, const float *g, float *dst) {
    int r = k / 2;
    float *tmp = (float *)malloc(sizeof(float) * (size_t)H * W);
    for (int c = 0; c < C; c++) {
        const float *s = src + (size_t)c * H * W;
        float *d = dst + (size_t)c * H * W;
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; y++) {
            const float *row = s + (size_t)y * W;
            float *o = tmp + (size_t)y * W;
            for (int x = 0; x < W; x++) {
                float acc = 0.f;
                for (int t = 0; t < k; t++) acc = fmaf(g[t], row[reflect_idx(x - r + t, W)], acc);
                o[x] = acc;
            }
        }
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; y++) {
            float *o = d + (size_t)y * W;
            for (int x = 0; x < W; x++) {
                float acc = 0.f;
                for (int t = 0; t < k; t++)
                    acc = fmaf(g[t], tmp[(size_t)reflect_idx(y - r + t, H) * W + x], acc);
                o[x] = acc;
            }
        }
    }
    free(tmp);
}

/* ------------------------------------------------------------------------------------------
 * apply_depth_gamma (stereo_core.py:91-107): pow(clamp(d,0.001,1), gamma).
 * Deterministic pow shared (by specification, not by code) with the CUDA path.  Table-driven, no division:
 *   x = m * 2^e, m in [1,2); i = top 7 mantissa bits, c_i = 1 + (i + 1/2)/128;
 *   r = fma(m, 1/c_i, -1)  (|r| < 2^-8);  ln(1+r) = r * Horner(1, -1/2, 1/3, -1/4, 1/5, -1/6, 1/7);
 *   log2(x) = fma(ln(1+r), log2(e), e + log2(c_i));
 *   t = g * log2(x); k = rint(64 t), j = k mod 64, n = (k - j)/64; f = (t - k/64) * ln2  (|f| < 2^-7.5);
 *   2^t = ldexp(2^(j/64) * Horner(1, 1, 1/2, 1/6, 1/24, 1/120), n); result rounded to f32.
 * The three tables (1/c_i, log2 c_i, 2^(j/64)) are themselves computed with a fixed series (below), so every
 * number is the result of IEEE-754 +,*,/,fma only => identical bits on any conforming machine, CPU or GPU.
 * The double result is accurate to ~1e-16, i.e. the float is correctly rounded except for ~1e-7 of inputs.
 * ---------------------------------------------------------------------------------------- */
static double tab_log2(double x) { /* series used for the tables only: atanh series of ln(m), m in [sqrt(1/2), sqrt(2)) */
    int e;
    double m = frexp(x, &e);
    if (m < 0.70710678118654752440) { m *= 2.0; e -= 1; }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 27.0;
    for (int d = 25; d >= 1; d -= 2) p = fma(p, z, 1.0 / (double)d);
    double ln_m = 2.0 * s * p;
    return fma(ln_m, 1.4426950408889634074, (double)e);
}
static double tab_exp2(double t) { /* tables only: Taylor series of exp(ln2 * f), |f| <= 1/2 */
    double n = nearbyint(t);
    double f = (t - n) * 0.69314718055994530942;
    static const double inv_fact[14] = {1.0, 1.0, 1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0,
                                        1.0 / 40320.0, 1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0,
                                        1.0 / 479001600.0, 1.0 / 6227020800.0};
    double p = inv_fact[13];
    for (int d = 12; d >= 0; d--) p = fma(p, f, inv_fact[d]);
    return ldexp(p, (int)n);
}
static double pw_invc[128], pw_log2c[128], pw_exp2t[64];
static int pw_ready;
static void pw_init(void) {
    if (pw_ready) return;
#pragma omp critical(orc_pw_init)
    if (!pw_ready) {
        for (int i = 0; i < 128; i++) {
            const double c = 1.0 + ((double)i + 0.5) / 128.0;
            pw_invc[i] = 1.0 / c;
            pw_log2c[i] = tab_log2(c);
        }
        for (int j = 0; j < 64; j++) pw_exp2t[j] = tab_exp2((double)j / 64.0);
        pw_ready = 1;
    }
}
ORC_API void orc_pow_tables(double *invc, double *log2c, double *exp2t) { /* for tests: the tables of the specification */
    pw_init();
    memcpy(invc, pw_invc, sizeof pw_invc); memcpy(log2c, pw_log2c, sizeof pw_log2c); memcpy(exp2t, pw_exp2t, sizeof pw_exp2t);
}
static double det_log2(double x) { /* x normal, positive */
    uint64_t b;
    memcpy(&b, &x, 8);
    const int e = (int)((b >> 52) & 0x7ff) - 1023, i = (int)((b >> 45) & 127);
    b = (b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m;
    memcpy(&m, &b, 8);
    const double r = fma(m, pw_invc[i], -1.0);
    double q = 1.0 / 7.0;
    q = fma(q, r, -1.0 / 6.0);
    q = fma(q, r, 1.0 / 5.0);
    q = fma(q, r, -1.0 / 4.0);
    q = fma(q, r, 1.0 / 3.0);
    q = fma(q, r, -1.0 / 2.0);
    q = fma(q, r, 1.0);
    const double ln1p = r * q;
    return fma(ln1p, 1.4426950408889634074, (double)e + pw_log2c[i]);
}
static double det_exp2(double t) {
    const double k = nearbyint(t * 64.0);
    const int ki = (int)k, j = ki & 63, n = (ki - j) / 64;
    const double f = (t - k * 0.015625) * 0.69314718055994530942;
    double p = 1.0 / 120.0;
    p = fma(p, f, 1.0 / 24.0);
    p = fma(p, f, 1.0 / 6.0);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    return ldexp(pw_exp2t[j] * p, n);
}

ORC_API float orc_powf(float x, float g) {
    /* ATen pow_tensor_scalar_optimized_kernel special cases (exactly rounded there and here) */
    if (g == 2.0f) return x * x;
    if (g == 3.0f) return (x * x) * x;
    if (x == 1.0f) return 1.0f;
    pw_init();
    return (float)det_exp2((double)g * det_log2((double)x));
}

ORC_API void orc_gamma(const float *d, size_t n, float gamma, float *out) {
    pw_init();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        float v = d[i];
        if (v < 0.001f) v = 0.001f;
        if (v > 1.0f) v = 1.0f;
        out[i] = orc_powf(v, gamma);
    }
}

/* ------------------------------------------------------------------------------------------
 * forward_warp_stereo (stereo_core.py:110-190), one direction (sign=+1 left view, -1 right).
 * Sort-free restatement: sources are scattered in ascending depth order, all
 * floor targets first, then all ceil targets with frac > 0.3.  Per target the surviving writer
 * of each pass is the source with the largest depth (ties cannot collide: equal depth => equal
 * disparity => distinct targets), and any ceil writer overrides the floor writer.
 *   image  [C,H,W] f32 planar, depth [H,W] f32, warped [C,H,W] f32 (0 where never written),
 *   mask   [H,W] u8 in {0,1} = (weight of the surviving writer > 0.1).
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_warp(const float *image, int C, const float *depth, int H, int W, float max_disp,
                      int sign, float *warped, uint8_t *mask) {
    memset(warped, 0, sizeof(float) * (size_t)C * H * W);
#pragma omp parallel
    {
        int *wf = (int *)malloc(sizeof(int) * W), *wc = (int *)malloc(sizeof(int) * W);
#pragma omp for schedule(static)
        for (int y = 0; y < H; y++) {
            const float *dr = depth + (size_t)y * W;
            for (int t = 0; t < W; t++) { wf[t] = -1; wc[t] = -1; }
            for (int x = 0; x < W; x++) {
                float disp = dr[x] * max_disp;
                float tx = sign > 0 ? (float)x + disp : (float)x + (-disp);
                float fl = floorf(tx);
                float frac = tx - fl;
                long t0 = (long)fl;
                if (t0 >= 0 && t0 < W) {
                    int cur = wf[t0];
                    if (cur < 0 || dr[x] > dr[cur]) wf[t0] = x;
                }
                long t1 = t0 + 1;
                if (t1 >= 0 && t1 < W && frac > 0.3f) {
                    int cur = wc[t1];
                    if (cur < 0 || dr[x] > dr[cur]) wc[t1] = x;
                }
            }
            for (int t = 0; t < W; t++) {
                int src = -1; float wgt = 0.f;
                if (wc[t] >= 0) {
                    src = wc[t];
                    float disp = dr[src] * max_disp;
                    float tx = sign > 0 ? (float)src + disp : (float)src + (-disp);
                    wgt = tx - floorf(tx);
                } else if (wf[t] >= 0) {
                    src = wf[t];
                    float disp = dr[src] * max_disp;
                    float tx = sign > 0 ? (float)src + disp : (float)src + (-disp);
                    wgt = 1.0f - (tx - floorf(tx));
                }
                mask[(size_t)y * W + t] = (uint8_t)(wgt > 0.1f);
                if (src >= 0)
                    for (int c = 0; c < C; c++)
                        warped[((size_t)c * H + y) * W + t] = image[((size_t)c * H + y) * W + src];
            }
        }
        free(wf); free(wc);
    }
}

/* ------------------------------------------------------------------------------------------
 * cv2.bilateralFilter(src8UC3, d, sigmaColor, sigmaSpace) (stereo_core.py:409-410).
 * OpenCV imgproc/bilateral_filter: radius=max(d/2,1), circular window, L1 colour distance LUT,
 * BORDER_REFLECT_101, float accumulation in tap order, cvRound(sum * (1/wsum)).
 * `use_fma` selects v_muladd-as-FMA accumulation (OpenCV's AVX2/FMA3 dispatch, the default here) vs
 * mul+add.  Pinned against cv2 4.13.0: with IPP disabled the FMA form differs on <= 2e-5 of the
 * values by 1 LSB (the non-FMA form on ~4e-5); cv2's default IPP build differs from both on ~1.4e-5.
 * ---------------------------------------------------------------------------------------- */
static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

ORC_API int orc_bilateral_tables(int d, double sigma_color, double sigma_space, float *color_w /*766*/,
                                 float *space_w, int *ofs_y, int *ofs_x) {
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    double gc = -0.5 / (sigma_color * sigma_color), gs = -0.5 / (sigma_space * sigma_space);
    int radius = d <= 0 ? (int)lrint(sigma_space * 1.5) : d / 2;
    if (radius < 1) radius = 1;
    for (int i = 0; i < 256 * 3; i++) color_w[i] = (float)exp((double)i * i * gc);
    int maxk = 0;
    for (int i = -radius; i <= radius; i++)
        for (int j = -radius; j <= radius; j++) {
            double r = sqrt((double)i * i + (double)j * j);
            if (r > radius) continue;
            space_w[maxk] = (float)exp(r * r * gs);
            ofs_y[maxk] = i; ofs_x[maxk] = j;
            maxk++;
        }
    return maxk;
}

ORC_API void orc_bilateral_u8c3(const uint8_t *src, int H, int W, int d, double sigma_color,
                                double sigma_space, int use_fma, uint8_t *dst) {
    float color_w[768], space_w[1024];
    int oy[1024], ox[1024];
    int maxk = orc_bilateral_tables(d, sigma_color, sigma_space, color_w, space_w, oy, ox);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            const uint8_t *p0 = src + ((size_t)y * W + x) * 3;
            int b0 = p0[0], g0 = p0[1], r0 = p0[2];
            float sb = 0.f, sg = 0.f, sr = 0.f, ws = 0.f;
            for (int k = 0; k < maxk; k++) {
                int yy = reflect101(y + oy[k], H), xx = reflect101(x + ox[k], W);
                const uint8_t *p = src + ((size_t)yy * W + xx) * 3;
                int b = p[0], g = p[1], r = p[2];
                float w = space_w[k] * color_w[abs(b - b0) + abs(g - g0) + abs(r - r0)];
                if (use_fma) {
                    sb = fmaf((float)b, w, sb); sg = fmaf((float)g, w, sg); sr = fmaf((float)r, w, sr);
                } else {
                    sb += (float)b * w; sg += (float)g * w; sr += (float)r * w;
                }
                ws += w;
            }
            ws = 1.f / ws;
            uint8_t *o = dst + ((size_t)y * W + x) * 3;
            o[0] = (uint8_t)clampi((int)lrintf(sb * ws), 0, 255);
            o[1] = (uint8_t)clampi((int)lrintf(sg * ws), 0, 255);
            o[2] = (uint8_t)clampi((int)lrintf(sr * ws), 0, 255);
        }
    }
}

/* cv2.dilate(mask, ones(3,3)) (stereo_core.py:455-456): 3x3 max, out-of-image ignored */
ORC_API void orc_dilate3(const uint8_t *m, int H, int W, uint8_t *out) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            uint8_t v = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int yy = y + dy, xx = x + dx;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    if (m[(size_t)yy * W + xx] > v) v = m[(size_t)yy * W + xx];
                }
            out[(size_t)y * W + x] = v;
        }
}

/* ------------------------------------------------------------------------------------------
 * cv2.inpaint(img8UC3, mask, radius, INPAINT_TELEA) (stereo_core.py:457).
 * OpenCV photo/inpaint.cpp (Telea 2004, fast marching).  SURVEY.md Appendix A.3.
 * The sorted linked list with FIFO ties is restated as a binary heap keyed (T, push sequence):
 * identical pop order.
 * ---------------------------------------------------------------------------------------- */
#define T_KNOWN 0
#define T_BAND 1
#define T_INSIDE 2
#define T_CHANGE 3

typedef struct { float T; uint32_t seq; int i, j; } hent;
typedef struct { hent *a; size_t n, cap; uint32_t seq; } pq_t;

static inline int hless(const hent *x, const hent *y) {
    return x->T < y->T || (x->T == y->T && x->seq < y->seq);
}
static void pq_push(pq_t *q, int i, int j, float T) {
    if (q->n == q->cap) { q->cap = q->cap ? q->cap * 2 : 1024; q->a = realloc(q->a, q->cap * sizeof(hent)); }
    hent e = {T, q->seq++, i, j};
    size_t k = q->n++;
    while (k > 0) {
        size_t p = (k - 1) / 2;
        if (!hless(&e, &q->a[p])) break;
        q->a[k] = q->a[p]; k = p;
    }
    q->a[k] = e;
}
static int pq_pop(pq_t *q, int *i, int *j) {
    if (q->n == 0) return 0;
    *i = q->a[0].i; *j = q->a[0].j;
    hent e = q->a[--q->n];
    size_t k = 0, n = q->n;
    for (;;) {
        size_t c = 2 * k + 1;
        if (c >= n) break;
        if (c + 1 < n && hless(&q->a[c + 1], &q->a[c])) c++;
        if (!hless(&q->a[c], &e)) break;
        q->a[k] = q->a[c]; k = c;
    }
    if (n) q->a[k] = e;
    return 1;
}

static inline float fmm_solve(int i1, int j1, int i2, int j2, const uint8_t *f, const float *t, int C) {
    double sol, a11 = t[(size_t)i1 * C + j1], a22 = t[(size_t)i2 * C + j2];
    double m12 = a11 < a22 ? a11 : a22;
    if (f[(size_t)i1 * C + j1] != T_INSIDE) {
        if (f[(size_t)i2 * C + j2] != T_INSIDE) {
            if (fabs(a11 - a22) >= 1.0) sol = 1 + m12;
            else sol = (a11 + a22 + sqrt((double)(2 - (a11 - a22) * (a11 - a22)))) * 0.5;
        } else sol = 1 + a11;
    } else if (f[(size_t)i2 * C + j2] != T_INSIDE) sol = 1 + a22;
    else sol = 1 + m12;
    return (float)sol;
}
static inline float min4f(float a, float b, float c, float d) {
    a = a < b ? a : b; c = c < d ? c : d; return a < c ? a : c;
}
static inline float fmm_min4(int i, int j, const uint8_t *f, const float *t, int C) {
    return min4f(fmm_solve(i - 1, j, i, j - 1, f, t, C), fmm_solve(i + 1, j, i, j - 1, f, t, C),
                 fmm_solve(i - 1, j, i, j + 1, f, t, C), fmm_solve(i + 1, j, i, j + 1, f, t, C));
}

/* img [H,W,3] u8 in/out; mask [H,W] nonzero = inpaint; t_out optional [(H+2)*(W+2)] */
ORC_API void orc_telea_u8c3(uint8_t *img, const uint8_t *mask, int H, int W, int radius, float *t_out) {
    const int R = H + 2, C = W + 2;
    int range = radius < 1 ? 1 : (radius > 100 ? 100 : radius);
    size_t N = (size_t)R * C;
    uint8_t *f = calloc(N, 1), *band = calloc(N, 1), *o = calloc(N, 1);
    float *t = malloc(N * sizeof(float));
    for (size_t k = 0; k < N; k++) t[k] = 1.0e6f;
    size_t nmask = 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            if (mask[(size_t)y * W + x]) { f[(size_t)(y + 1) * C + x + 1] = T_INSIDE; nmask++; }
    if (nmask == 0) goto done;
    /* band = dilate(f, 3x3 cross) - f, frame zeroed */
    for (int i = 1; i < R - 1; i++)
        for (int j = 1; j < C - 1; j++) {
            size_t p = (size_t)i * C + j;
            if (f[p]) continue;
            if (f[p - C] || f[p + C] || f[p - 1] || f[p + 1]) band[p] = T_INSIDE;
        }
    pq_t heap = {0}, outq = {0};
    for (int i = 1; i < R - 1; i++)
        for (int j = 1; j < C - 1; j++)
            if (band[(size_t)i * C + j]) {
                pq_push(&heap, i, j, 0.f); pq_push(&outq, i, j, 0.f);
                t[(size_t)i * C + j] = 0.f;
            }
    /* outer ring: o = dilate(f, (2r+1)^2 rect) - f - band, frame zeroed */
    for (int i = 1; i < R - 1; i++)
        for (int j = 1; j < C - 1; j++) {
            size_t p = (size_t)i * C + j;
            if (f[p] || band[p]) continue;
            int hit = 0;
            for (int di = -range; di <= range && !hit; di++) {
                int ii = i + di; if (ii < 0 || ii >= R) continue;
                for (int dj = -range; dj <= range; dj++) {
                    int jj = j + dj; if (jj < 0 || jj >= C) continue;
                    if (f[(size_t)ii * C + jj]) { hit = 1; break; }
                }
            }
            if (hit) o[p] = T_INSIDE;
        }
    {   /* icvCalcFMM(out, t, Out, negate=true) */
        int ii, jj;
        while (pq_pop(&outq, &ii, &jj)) {
            o[(size_t)ii * C + jj] = T_CHANGE;
            for (int q = 0; q < 4; q++) {
                int i = ii + (q == 0 ? -1 : (q == 2 ? 1 : 0)), j = jj + (q == 1 ? -1 : (q == 3 ? 1 : 0));
                if (i <= 0 || j <= 0 || i > R - 1 || j > C - 1) continue;
                if (i >= R - 1 || j >= C - 1) continue; /* frame is never INSIDE */
                if (o[(size_t)i * C + j] == T_INSIDE) {
                    float dist = fmm_min4(i, j, o, t, C);
                    t[(size_t)i * C + j] = dist;
                    o[(size_t)i * C + j] = T_BAND;
                    pq_push(&outq, i, j, dist);
                }
            }
        }
        for (size_t k = 0; k < N; k++)
            if (o[k] == T_CHANGE) { o[k] = T_KNOWN; t[k] = -t[k]; }
    }
    {   /* icvTeleaInpaintFMM(mask=f, t, out, range, Heap), 3 channels */
        int ii, jj;
        while (pq_pop(&heap, &ii, &jj)) {
            f[(size_t)ii * C + jj] = T_KNOWN;
            for (int q = 0; q < 4; q++) {
                int i = ii + (q == 0 ? -1 : (q == 2 ? 1 : 0)), j = jj + (q == 1 ? -1 : (q == 3 ? 1 : 0));
                if (i <= 0 || j <= 0 || i > R - 1 || j > C - 1) continue;
                if (i >= R - 1 || j >= C - 1) continue;
                size_t p = (size_t)i * C + j;
                if (f[p] != T_INSIDE) continue;
                float dist = fmm_min4(i, j, f, t, C);
                t[p] = dist;
                float gtx, gty;
                if (f[p + 1] != T_INSIDE) {
                    if (f[p - 1] != T_INSIDE) gtx = (t[p + 1] - t[p - 1]) * 0.5f;
                    else gtx = t[p + 1] - t[p];
                } else {
                    if (f[p - 1] != T_INSIDE) gtx = t[p] - t[p - 1];
                    else gtx = 0.f;
                }
                if (f[p + C] != T_INSIDE) {
                    if (f[p - C] != T_INSIDE) gty = (t[p + C] - t[p - C]) * 0.5f;
                    else gty = t[p + C] - t[p];
                } else {
                    if (f[p - C] != T_INSIDE) gty = t[p] - t[p - C];
                    else gty = 0.f;
                }
                float Jx[3] = {0, 0, 0}, Jy[3] = {0, 0, 0}, Ia[3] = {0, 0, 0};
                float s[3] = {1.0e-20f, 1.0e-20f, 1.0e-20f};
                for (int k = i - range; k <= i + range; k++) {
                    int km = k - 1 + (k == 1), kp = k - 1 - (k == R - 2);
                    for (int l = j - range; l <= j + range; l++) {
                        int lm = l - 1 + (l == 1), lp = l - 1 - (l == C - 2);
                        if (!(k > 0 && l > 0 && k < R - 1 && l < C - 1)) continue;
                        size_t pk = (size_t)k * C + l;
                        if (f[pk] == T_INSIDE) continue;
                        if ((l - j) * (l - j) + (k - i) * (k - i) > range * range) continue;
                        float ry = (float)(i - k), rx = (float)(j - l);
                        float vl = rx * rx + ry * ry;
                        float dst = (float)(1. / ((double)vl * sqrt((double)vl)));
                        float lev = (float)(1. / (1 + fabs((double)(t[pk] - t[p]))));
                        float dir = rx * gtx + ry * gty;
                        if (fabs((double)dir) <= 0.01) dir = 0.000001f;
                        float w = fabsf(dst * lev * dir);
                        int fr = f[pk + 1] != T_INSIDE, fl = f[pk - 1] != T_INSIDE;
                        int fd = f[pk + C] != T_INSIDE, fu = f[pk - C] != T_INSIDE;
                        for (int c = 0; c < 3; c++) {
#define PIX(yy, xx) ((int)img[((size_t)(yy) * W + (xx)) * 3 + c])
                            float gix, giy;
                            if (fr) {
                                if (fl) gix = (float)(PIX(km, lp + 1) - PIX(km, lm - 1)) * 2.0f;
                                else gix = (float)(PIX(km, lp + 1) - PIX(km, lm));
                            } else {
                                if (fl) gix = (float)(PIX(km, lp) - PIX(km, lm - 1));
                                else gix = 0.f;
                            }
                            if (fd) {
                                if (fu) giy = (float)(PIX(kp + 1, lm) - PIX(km - 1, lm)) * 2.0f;
                                else giy = (float)(PIX(kp + 1, lm) - PIX(km, lm));
                            } else {
                                if (fu) giy = (float)(PIX(kp, lm) - PIX(km - 1, lm));
                                else giy = 0.f;
                            }
                            Ia[c] += w * (float)PIX(k - 1, l - 1);
                            Jx[c] -= w * (gix * rx);
                            Jy[c] -= w * (giy * ry);
                            s[c] += w;
#undef PIX
                        }
                    }
                }
                for (int c = 0; c < 3; c++) {
                    float sat = (float)(Ia[c] / s[c] +
                                        (Jx[c] + Jy[c]) / (sqrt(Jx[c] * Jx[c] + Jy[c] * Jy[c]) + 1.0e-20f) + 0.5f);
                    long r = lrintf(sat);
                    img[((size_t)(i - 1) * W + (j - 1)) * 3 + c] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
                }
                f[p] = T_BAND;
                pq_push(&heap, i, j, dist);
            }
        }
    }
    free(heap.a); free(outq.a);
done:
    if (t_out) memcpy(t_out, t, N * sizeof(float));
    free(f); free(band); free(o); free(t);
}

/* ------------------------------------------------------------------------------------------
 * Two-pass restatement of the same algorithm (design check for the GPU march, DESIGN.md 6 "next"):
 * the arrival times T and the ORDER in which hole pixels are computed depend on the mask alone, so
 *   pass A runs the fast march without colours and records ord[p] (0,1,2,... in computation order);
 *   pass B visits the hole pixels in that order and applies the Telea formula, where "pixel q is still
 *   INSIDE at the time p is computed" is simply ord[q] >= ord[p] - no flags, no queue.
 * Must give exactly orc_telea_u8c3's result (tests/test_oracle_vs_libs.py).
 * ---------------------------------------------------------------------------------------- */
/* ord_out: optional [(H+2)*(W+2)] computation order of the hole pixels (-1 elsewhere), for analysis */
ORC_API void orc_telea_u8c3_two_pass(uint8_t *img, const uint8_t *mask, int H, int W, int radius, int32_t *ord_out) {
    const int R = H + 2, C = W + 2;
    int range = radius < 1 ? 1 : (radius > 100 ? 100 : radius);
    size_t N = (size_t)R * C;
    uint8_t *f = calloc(N, 1), *band = calloc(N, 1), *o = calloc(N, 1);
    float *t = malloc(N * sizeof(float));
    int32_t *ord = malloc(N * sizeof(int32_t));      /* -1: never a hole; INT32_MAX: hole not computed */
    uint32_t *seq = NULL;
    size_t nseq = 0, nmask = 0;
    for (size_t k = 0; k < N; k++) { t[k] = 1.0e6f; ord[k] = -1; }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            if (mask[(size_t)y * W + x]) { size_t p = (size_t)(y + 1) * C + x + 1; f[p] = T_INSIDE; ord[p] = INT32_MAX; nmask++; }
    if (nmask == 0) goto done2;
    seq = malloc(nmask * sizeof(uint32_t));
    for (int i = 1; i < R - 1; i++)
        for (int j = 1; j < C - 1; j++) {
            size_t p = (size_t)i * C + j;
            if (f[p]) continue;
            if (f[p - C] || f[p + C] || f[p - 1] || f[p + 1]) band[p] = T_INSIDE;
        }
    pq_t heap = {0}, outq = {0};
    for (int i = 1; i < R - 1; i++)
        for (int j = 1; j < C - 1; j++)
            if (band[(size_t)i * C + j]) { pq_push(&heap, i, j, 0.f); pq_push(&outq, i, j, 0.f); t[(size_t)i * C + j] = 0.f; }
    for (int i = 1; i < R - 1; i++)
        for (int j = 1; j < C - 1; j++) {
            size_t p = (size_t)i * C + j;
            if (f[p] || band[p]) continue;
            int hit = 0;
            for (int di = -range; di <= range && !hit; di++) {
                int ii = i + di; if (ii < 0 || ii >= R) continue;
                for (int dj = -range; dj <= range; dj++) {
                    int jj = j + dj; if (jj < 0 || jj >= C) continue;
                    if (f[(size_t)ii * C + jj]) { hit = 1; break; }
                }
            }
            if (hit) o[p] = T_INSIDE;
        }
    {   /* outer ring distances, as in orc_telea_u8c3 */
        int ii, jj;
        while (pq_pop(&outq, &ii, &jj)) {
            o[(size_t)ii * C + jj] = T_CHANGE;
            for (int q = 0; q < 4; q++) {
                int i = ii + (q == 0 ? -1 : (q == 2 ? 1 : 0)), j = jj + (q == 1 ? -1 : (q == 3 ? 1 : 0));
                if (i <= 0 || j <= 0 || i >= R - 1 || j >= C - 1) continue;
                if (o[(size_t)i * C + j] == T_INSIDE) {
                    float dist = fmm_min4(i, j, o, t, C);
                    t[(size_t)i * C + j] = dist; o[(size_t)i * C + j] = T_BAND;
                    pq_push(&outq, i, j, dist);
                }
            }
        }
        for (size_t k = 0; k < N; k++)
            if (o[k] == T_CHANGE) { o[k] = T_KNOWN; t[k] = -t[k]; }
    }
    {   /* pass A: the inpainting march WITHOUT colours - arrival times and computation order only */
        int ii, jj;
        while (pq_pop(&heap, &ii, &jj)) {
            f[(size_t)ii * C + jj] = T_KNOWN;
            for (int q = 0; q < 4; q++) {
                int i = ii + (q == 0 ? -1 : (q == 2 ? 1 : 0)), j = jj + (q == 1 ? -1 : (q == 3 ? 1 : 0));
                if (i <= 0 || j <= 0 || i >= R - 1 || j >= C - 1) continue;
                size_t p = (size_t)i * C + j;
                if (f[p] != T_INSIDE) continue;
                float dist = fmm_min4(i, j, f, t, C);
                t[p] = dist;
                ord[p] = (int32_t)nseq; seq[nseq++] = (uint32_t)p;
                f[p] = T_BAND;
                pq_push(&heap, i, j, dist);
            }
        }
    }
    /* pass B: colours in the recorded order; INSIDE(q) at step k  <=>  ord[q] >= k */
    for (size_t k = 0; k < nseq; k++) {
        const size_t p = seq[k];
        const int i = (int)(p / C), j = (int)(p % C);
        const int32_t kk = (int32_t)k;
#define INS(q) (ord[q] >= kk)
        float gtx, gty;
        if (!INS(p + 1)) { if (!INS(p - 1)) gtx = (t[p + 1] - t[p - 1]) * 0.5f; else gtx = t[p + 1] - t[p]; }
        else { if (!INS(p - 1)) gtx = t[p] - t[p - 1]; else gtx = 0.f; }
        if (!INS(p + C)) { if (!INS(p - C)) gty = (t[p + C] - t[p - C]) * 0.5f; else gty = t[p + C] - t[p]; }
        else { if (!INS(p - C)) gty = t[p] - t[p - C]; else gty = 0.f; }
        float Jx[3] = {0, 0, 0}, Jy[3] = {0, 0, 0}, Ia[3] = {0, 0, 0};
        float s[3] = {1.0e-20f, 1.0e-20f, 1.0e-20f};
        for (int kr = i - range; kr <= i + range; kr++) {
            int km = kr - 1 + (kr == 1), kp = kr - 1 - (kr == R - 2);
            for (int l = j - range; l <= j + range; l++) {
                int lm = l - 1 + (l == 1), lp = l - 1 - (l == C - 2);
                if (!(kr > 0 && l > 0 && kr < R - 1 && l < C - 1)) continue;
                size_t pk = (size_t)kr * C + l;
                if (INS(pk)) continue;
                if ((l - j) * (l - j) + (kr - i) * (kr - i) > range * range) continue;
                float ry = (float)(i - kr), rx = (float)(j - l);
                float vl = rx * rx + ry * ry;
                float dst = (float)(1. / ((double)vl * sqrt((double)vl)));
                float lev = (float)(1. / (1 + fabs((double)(t[pk] - t[p]))));
                float dir = rx * gtx + ry * gty;
                if (fabs((double)dir) <= 0.01) dir = 0.000001f;
                float w = fabsf(dst * lev * dir);
                int fr = !INS(pk + 1), fl = !INS(pk - 1), fd = !INS(pk + C), fu = !INS(pk - C);
                for (int c = 0; c < 3; c++) {
#define PIX(yy, xx) ((int)img[((size_t)(yy) * W + (xx)) * 3 + c])
                    float gix, giy;
                    if (fr) { if (fl) gix = (float)(PIX(km, lp + 1) - PIX(km, lm - 1)) * 2.0f; else gix = (float)(PIX(km, lp + 1) - PIX(km, lm)); }
                    else { if (fl) gix = (float)(PIX(km, lp) - PIX(km, lm - 1)); else gix = 0.f; }
                    if (fd) { if (fu) giy = (float)(PIX(kp + 1, lm) - PIX(km - 1, lm)) * 2.0f; else giy = (float)(PIX(kp + 1, lm) - PIX(km, lm)); }
                    else { if (fu) giy = (float)(PIX(kp, lm) - PIX(km - 1, lm)); else giy = 0.f; }
                    Ia[c] += w * (float)PIX(kr - 1, l - 1);
                    Jx[c] -= w * (gix * rx);
                    Jy[c] -= w * (giy * ry);
                    s[c] += w;
#undef PIX
                }
            }
        }
        for (int c = 0; c < 3; c++) {
            float sat = (float)(Ia[c] / s[c] + (Jx[c] + Jy[c]) / (sqrt(Jx[c] * Jx[c] + Jy[c] * Jy[c]) + 1.0e-20f) + 0.5f);
            long r = lrintf(sat);
            img[((size_t)(i - 1) * W + (j - 1)) * 3 + c] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
        }
#undef INS
    }
    free(heap.a); free(outq.a);
done2:
    if (ord_out) memcpy(ord_out, ord, N * sizeof(int32_t));
    free(f); free(band); free(o); free(t); free(ord); free(seq);
}

/* ------------------------------------------------------------------------------------------
 * _sharpen_image (stereo_core.py:414-434): blur = gauss(5x5, sigma 1, reflect);
 * out = clamp(img + s*(img - blur), 0, 255) with unfused sub, mul, add.  planar [C,H,W].
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_sharpen(const float *src, int C, int H, int W, const float *g5, float strength, float *dst) {
    float *blur = (float *)malloc(sizeof(float) * (size_t)C * H * W);
    orc_gauss_blur(src, C, H, W, 5, g5, blur);
    size_t n = (size_t)C * H * W;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        float d = src[i] - blur[i];
        float m = strength * d;
        float v = src[i] + m;
        dst[i] = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
    }
    free(blur);
}

/* ------------------------------------------------------------------------------------------
 * F.interpolate(mode='area') = adaptive_avg_pool2d (stereo_core.py:298-299): window
 * [floor(o*in/out), ceil((o+1)*in/out)), raster-order f32 sum, then / kh / kw.
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_area_pool(const float *src, int C, int H, int W, int oH, int oW, float *dst) {
    for (int c = 0; c < C; c++) {
        const float *s = src + (size_t)c * H * W;
        float *d = dst + (size_t)c * oH * oW;
#pragma omp parallel for schedule(static)
        for (int oy = 0; oy < oH; oy++) {
            int y0 = (int)(((long)oy * H) / oH), y1 = (int)((((long)oy + 1) * H + oH - 1) / oH);
            for (int ox = 0; ox < oW; ox++) {
                int x0 = (int)(((long)ox * W) / oW), x1 = (int)((((long)ox + 1) * W + oW - 1) / oW);
                float sum = 0.f;
                for (int y = y0; y < y1; y++)
                    for (int x = x0; x < x1; x++) sum += s[(size_t)y * W + x];
                d[(size_t)oy * oW + ox] = sum / (float)(y1 - y0) / (float)(x1 - x0);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Depth-map post-processing on the producer side (/root/reference/depth_map_generator.py:217-236, SURVEY.md 8(f)
 * rank 3): cv2.resize(depth f32, (W, H), INTER_LINEAR) -> min / max -> (d - min) / range -> * 255 | 65535 -> round
 * half to even -> u8 | u16.  Returns 0 (nothing written) when the resized map is flat, like the reference.
 * OpenCV imgproc/resize.cpp, 32F linear: scale = 1 / (dst / src) in double; fx = float((dx + 0.5) * scale - 0.5);
 * sx = floor(fx), fx -= sx; sx < 0 -> (0, 0); sx >= src - 1 -> (src - 1, 0); taps (1 - fx, fx) in float;
 * horizontal pass S[sx] * a0 + S[sx + 1] * a1, then vertical R0 * b0 + R1 * b1, every product and sum rounded to float
 * (OpenCV's own code, cv2.ipp.setUseIPP(False): bit-exact except the first / last output rows when they clamp, where
 * its single-row SIMD path fuses the multiply-add, <= 1 ulp; the IPP build that ships differs by up to 5e-6 relative).
 * ---------------------------------------------------------------------------------------- */
static void linear_axis(int ssize, int dsize, int *ofs, float *a0, float *a1) {
    const double scale = 1.0 / ((double)dsize / (double)ssize);
    for (int d = 0; d < dsize; d++) {
        float fx = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(fx);
        fx -= (float)s;
        if (s < 0) { fx = 0.f; s = 0; }
        if (s >= ssize - 1) { fx = 0.f; s = ssize - 1; }
        ofs[d] = s; a0[d] = 1.f - fx; a1[d] = fx;
    }
}
ORC_API void orc_resize_linear_f32(const float *src, int h, int w, int H, int W, float *dst) {
    int *ox = malloc(sizeof(int) * W), *oy = malloc(sizeof(int) * H);
    float *ax0 = malloc(sizeof(float) * W), *ax1 = malloc(sizeof(float) * W), *ay0 = malloc(sizeof(float) * H), *ay1 = malloc(sizeof(float) * H);
    linear_axis(w, W, ox, ax0, ax1);
    linear_axis(h, H, oy, ay0, ay1);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        const float *r0 = src + (size_t)oy[y] * w, *r1 = src + (size_t)(oy[y] + 1 < h ? oy[y] + 1 : h - 1) * w;
        for (int x = 0; x < W; x++) {
            const int x0 = ox[x], x1 = x0 + 1 < w ? x0 + 1 : w - 1;
            const float t0 = r0[x0] * ax0[x], t1 = r0[x1] * ax1[x], top = t0 + t1;
            const float u0 = r1[x0] * ax0[x], u1 = r1[x1] * ax1[x], bot = u0 + u1;
            const float v0 = top * ay0[y], v1 = bot * ay1[y];
            dst[(size_t)y * W + x] = v0 + v1;
        }
    }
    free(ox); free(oy); free(ax0); free(ax1); free(ay0); free(ay1);
}
ORC_API int orc_depth_post(const float *src, int h, int w, int H, int W, int bits, void *dst) {
    const size_t n = (size_t)H * W;
    float *r = malloc(sizeof(float) * n);
    orc_resize_linear_f32(src, h, w, H, W, r);
    float mn = r[0], mx = r[0];
    for (size_t i = 1; i < n; i++) { mn = r[i] < mn ? r[i] : mn; mx = r[i] > mx ? r[i] : mx; }
    const float range = mx - mn;
    if (!(range > 0.f)) { free(r); return 0; }
    const float q = bits == 16 ? 65535.f : 255.f;
    for (size_t i = 0; i < n; i++) {
        const float nv = (r[i] - mn) / range;
        const float v = rintf(nv * q);
        if (bits == 16) ((uint16_t *)dst)[i] = (uint16_t)v; else ((uint8_t *)dst)[i] = (uint8_t)v;
    }
    free(r);
    return 1;
}
