/*
 * ck_oracle.c — CPU ORACLE for the CamKifu stone-detection hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product path (camkifu_b200/) never links, imports or calls it.
 *
 * The reference (ArnaudPel/CamKifu) is pure Python; the arithmetic of its hot path lives in third-party native code
 * that is NOT under /root/reference:
 *     OpenCV  (README.md:11 "OpenCV 3"; ckmain.py:53 asserts 3.1.0; the image here has cv2 4.13.0)
 *     Keras-1.x on Theano (README.md:14)
 * This file restates the published algorithms of those calls, anchored on the reference's own call sites, and is
 * PINNED by (a) bit-exact comparison against cv2 4.13.0 in this image (tests/test_oracle_vs_cv2.py), (b) golden
 * vectors produced by running the unmodified reference modules from /root/reference (oracle/gen_golden.py ->
 * tests/golden/), (c) the reference's own known-answer tests for the label codec
 * (test/camkifu/stone/test_tmanager.py:18-27).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -fopenmp  (no FMA contraction: OpenCV's baseline build has none).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define CKO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------------------------
 * 3x3 inverse as cv::invert does it for a CV_64F 3x3 (DECOMP_LU special case: cofactors * 1/det).
 * Reference call site: cv2.warpPerspective(frame, transform, shape) stonesfinder.py:140 — OpenCV inverts `transform`
 * internally because WARP_INVERSE_MAP is not set.
 * ---------------------------------------------------------------------------------------------------------------- */
CKO_API int cko_invert3x3(const double *S, double *T)
{
#define s(i, j) S[(i) * 3 + (j)]
    double d = s(0, 0) * (s(1, 1) * s(2, 2) - s(1, 2) * s(2, 1)) - s(0, 1) * (s(1, 0) * s(2, 2) - s(1, 2) * s(2, 0)) +
               s(0, 2) * (s(1, 0) * s(2, 1) - s(1, 1) * s(2, 0));
    if (d == 0.) {
        memset(T, 0, 9 * sizeof(double));
        return 0;
    }
    d = 1. / d;
    double t[9];
    t[0] = (s(1, 1) * s(2, 2) - s(1, 2) * s(2, 1)) * d;
    t[1] = (s(0, 2) * s(2, 1) - s(0, 1) * s(2, 2)) * d;
    t[2] = (s(0, 1) * s(1, 2) - s(0, 2) * s(1, 1)) * d;
    t[3] = (s(1, 2) * s(2, 0) - s(1, 0) * s(2, 2)) * d;
    t[4] = (s(0, 0) * s(2, 2) - s(0, 2) * s(2, 0)) * d;
    t[5] = (s(0, 2) * s(1, 0) - s(0, 0) * s(1, 2)) * d;
    t[6] = (s(1, 0) * s(2, 1) - s(1, 1) * s(2, 0)) * d;
    t[7] = (s(0, 1) * s(2, 0) - s(0, 0) * s(2, 1)) * d;
    t[8] = (s(0, 0) * s(1, 1) - s(0, 1) * s(1, 0)) * d;
    memcpy(T, t, sizeof t);
#undef s
    return 1;
}

/* ------------------------------------------------------------------------------------------------------------------
 * cv2.warpPerspective(src 8UC3, M, (dw, dh))  flags = INTER_LINEAR, BORDER_CONSTANT(0)      [stonesfinder.py:140]
 *
 * OpenCV's fixed-point remap: source coordinates in float64 per 64-wide destination block
 *   X0 = Mi0*bx + Mi1*y + Mi2 ;  W = W0 + Mi6*x1 ; W = W ? 32/W : 0 ; X = saturate_int(rint((X0 + Mi0*x1)*W))
 * then 5 fractional bits select a 10-bit bilinear weight table ((32-fx)(32-fy) ... normalised to sum 1024), taps
 * outside the image read the constant border 0, and the blend is (sum + 512) >> 10.
 * ---------------------------------------------------------------------------------------------------------------- */
static inline int sat_int_from_double(double v)
{
    if (v < (double)INT32_MIN) v = (double)INT32_MIN;
    if (v > (double)INT32_MAX) v = (double)INT32_MAX;
    return (int)lrint(v); /* round half to even under the default rounding mode, as cvRound */
}

static inline int sat_s16(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

CKO_API void cko_warp_perspective_u8c3(const uint8_t *src, int sh, int sw, size_t sstep, const double *M /*src->dst*/,
                                       uint8_t *dst, int dh, int dw)
{
    double Mi[9];
    cko_invert3x3(M, Mi);
    const int BW = 64;
    for (int y = 0; y < dh; y++) {
        for (int bx = 0; bx < dw; bx += BW) {
            double X0 = Mi[0] * bx + Mi[1] * y + Mi[2];
            double Y0 = Mi[3] * bx + Mi[4] * y + Mi[5];
            double W0 = Mi[6] * bx + Mi[7] * y + Mi[8];
            int bw = dw - bx < BW ? dw - bx : BW;
            for (int x1 = 0; x1 < bw; x1++) {
                double W = W0 + Mi[6] * x1;
                W = W ? 32. / W : 0;
                int X = sat_int_from_double((X0 + Mi[0] * x1) * W);
                int Y = sat_int_from_double((Y0 + Mi[3] * x1) * W);
                int sx = sat_s16(X >> 5), sy = sat_s16(Y >> 5);
                int fx = X & 31, fy = Y & 31;
                /* integer weight table as cv::initInterTab2D(INTER_LINEAR, fixpt): products of the 1/32 taps scaled
                   to 1<<10; for the linear kernel these are exact and need no rounding fix-up */
                int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy), w10 = (32 - fx) * fy, w11 = fx * fy;
                uint8_t *d = dst + ((size_t)y * dw + bx + x1) * 3;
                for (int c = 0; c < 3; c++) {
                    int t00 = 0, t01 = 0, t10 = 0, t11 = 0;
                    if ((unsigned)sy < (unsigned)sh) {
                        if ((unsigned)sx < (unsigned)sw) t00 = src[(size_t)sy * sstep + sx * 3 + c];
                        if ((unsigned)(sx + 1) < (unsigned)sw) t01 = src[(size_t)sy * sstep + (sx + 1) * 3 + c];
                    }
                    if ((unsigned)(sy + 1) < (unsigned)sh) {
                        if ((unsigned)sx < (unsigned)sw) t10 = src[(size_t)(sy + 1) * sstep + sx * 3 + c];
                        if ((unsigned)(sx + 1) < (unsigned)sw) t11 = src[(size_t)(sy + 1) * sstep + (sx + 1) * 3 + c];
                    }
                    d[c] = (uint8_t)((t00 * w00 + t01 * w01 + t10 * w10 + t11 * w11 + 512) >> 10);
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * cv2.accumulateWeighted(src_u8, acc_f32, alpha)                                       [sf_clustering.py:33-36]
 * With OpenCV's default optimisations (IPP HAL) this is acc = fma(alpha_f32, (float)src - acc, acc).
 * first == 1 reproduces `self.accu = gframe.astype(np.float32)` (sf_clustering.py:34).
 * ---------------------------------------------------------------------------------------------------------------- */
CKO_API void cko_accumulate_weighted(const uint8_t *src, float *acc, size_t n, float alpha, int first)
{
    if (first) {
        for (size_t i = 0; i < n; i++) acc[i] = (float)src[i];
        return;
    }
    for (size_t i = 0; i < n; i++) acc[i] = fmaf(alpha, (float)src[i] - acc[i], acc[i]);
}

/* ------------------------------------------------------------------------------------------------------------------
 * cv2.kmeans(pixels[N,3] f32, K=3, None, (TERM_CRITERIA_EPS, 15, 3), attempts=3, KMEANS_PP_CENTERS)
 *                                                                                       [sf_clustering.py:103-104]
 * cv::RNG (multiply-with-carry), k-means++ seeding with 3 trials, Lloyd iterations with sequential float32 centre
 * sums, stop when the largest squared centre shift <= eps^2 (criteria has EPS only => maxCount = 100), labels are
 * NOT re-assigned on the last iteration, best of `attempts` by compactness (strict <).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct { uint64_t state; } cko_rng;

static inline uint32_t rng_next(cko_rng *r)
{
    r->state = (uint64_t)(uint32_t)r->state * 4164903690ULL + (uint32_t)(r->state >> 32);
    return (uint32_t)r->state;
}
static inline double rng_double(cko_rng *r)
{
    uint32_t t = rng_next(r);
    return (double)(((uint64_t)t << 32) | rng_next(r)) * 5.4210108624275221700372640043497e-20;
}

CKO_API uint64_t cko_rng_seed_state(uint32_t seed) { return seed ? (uint64_t)seed : 0xffffffffULL; } /* cv::RNG(s) */

static inline float l2sqr3(const float *a, const float *b)
{
    float t0 = a[0] - b[0], t1 = a[1] - b[1], t2 = a[2] - b[2];
    float d = t0 * t0;
    d += t1 * t1;
    d += t2 * t2;
    return d;
}

static void centers_pp(const float *data, int N, float *out_centers, int K, cko_rng *rng, int trials, float *buf)
{
    float *dist = buf, *tdist = buf + N, *tdist2 = tdist + N;
    int centers[16];
    double sum0 = 0;
    centers[0] = (int)(rng_next(rng) % (uint32_t)N);
    for (int i = 0; i < N; i++) {
        dist[i] = l2sqr3(data + 3 * (size_t)i, data + 3 * (size_t)centers[0]);
        sum0 += dist[i];
    }
    for (int k = 1; k < K; k++) {
        double bestSum = DBL_MAX;
        int bestCenter = -1;
        for (int j = 0; j < trials; j++) {
            double p = rng_double(rng) * sum0;
            int ci = 0;
            for (; ci < N - 1; ci++) {
                p -= dist[ci];
                if (p <= 0) break;
            }
            double s = 0;
            for (int i = 0; i < N; i++) {
                float d = l2sqr3(data + 3 * (size_t)i, data + 3 * (size_t)ci);
                tdist2[i] = d < dist[i] ? d : dist[i];
                s += tdist2[i];
            }
            if (s < bestSum) {
                bestSum = s;
                bestCenter = ci;
                float *t = tdist; tdist = tdist2; tdist2 = t;
            }
        }
        centers[k] = bestCenter; /* OpenCV raises StsNoConv when < 0 (NaN / huge input); not reachable for u8-range data */
        sum0 = bestSum;
        { float *t = dist; dist = tdist; tdist = t; }
    }
    for (int k = 0; k < K; k++)
        for (int j = 0; j < 3; j++) out_centers[k * 3 + j] = data[3 * (size_t)centers[k] + j];
}

/* rng_state: in/out (cv::theRNG() is process-global and carries across calls; the caller owns it here).
 * iters_out (optional, [attempts]): number of loop iterations each attempt ran.
 * returns best compactness; labels[N], centers[K*3] of the best attempt. */
CKO_API double cko_kmeans3(const float *data, int N, int K, double eps, int max_count_or_0, int attempts,
                           uint64_t *rng_state, int32_t *best_labels, float *best_centers, int *iters_out)
{
    cko_rng rng = { *rng_state };
    double eps2 = (eps < 0 ? 0 : eps) * (eps < 0 ? 0 : eps);
    int maxCount = 100;
    if (max_count_or_0 > 0) maxCount = max_count_or_0 < 2 ? 2 : (max_count_or_0 > 100 ? 100 : max_count_or_0);
    if (K == 1) { attempts = 1; maxCount = 2; }
    float *buf = (float *)malloc(sizeof(float) * 3 * (size_t)N);
    int32_t *labels = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
    double *dists = (double *)malloc(sizeof(double) * (size_t)N);
    float centers[16 * 3], old_centers[16 * 3], temp[3];
    int counters[16];
    double best_compactness = DBL_MAX;
    memset(centers, 0, sizeof centers);
    memset(old_centers, 0, sizeof old_centers);
    for (int a = 0; a < attempts; a++) {
        double compactness = 0;
        for (int iter = 0;;) {
            double max_center_shift = iter == 0 ? DBL_MAX : 0.0;
            { float t[48]; memcpy(t, centers, sizeof t); memcpy(centers, old_centers, sizeof t); memcpy(old_centers, t, sizeof t); }
            if (iter == 0) {
                centers_pp(data, N, centers, K, &rng, 3, buf);
            } else {
                memset(centers, 0, sizeof(float) * 3 * K);
                for (int k = 0; k < K; k++) counters[k] = 0;
                for (int i = 0; i < N; i++) {
                    const float *sample = data + 3 * (size_t)i;
                    int k = labels[i];
                    float *c = centers + 3 * k;
                    c[0] += sample[0]; c[1] += sample[1]; c[2] += sample[2];
                    counters[k]++;
                }
                for (int k = 0; k < K; k++) {
                    if (counters[k] != 0) continue;
                    /* empty cluster: split the farthest point off the biggest cluster */
                    int max_k = 0;
                    for (int k1 = 1; k1 < K; k1++) if (counters[max_k] < counters[k1]) max_k = k1;
                    double max_dist = 0;
                    int farthest_i = -1;
                    float *base_center = centers + 3 * max_k;
                    float scale = 1.f / counters[max_k];
                    for (int j = 0; j < 3; j++) temp[j] = base_center[j] * scale;
                    for (int i = 0; i < N; i++) {
                        if (labels[i] != max_k) continue;
                        double dist = l2sqr3(data + 3 * (size_t)i, temp);
                        if (max_dist <= dist) { max_dist = dist; farthest_i = i; }
                    }
                    counters[max_k]--;
                    counters[k]++;
                    labels[farthest_i] = k;
                    const float *sample = data + 3 * (size_t)farthest_i;
                    float *cur = centers + 3 * k;
                    for (int j = 0; j < 3; j++) { base_center[j] -= sample[j]; cur[j] += sample[j]; }
                }
                for (int k = 0; k < K; k++) {
                    float *c = centers + 3 * k;
                    float scale = 1.f / counters[k];
                    for (int j = 0; j < 3; j++) c[j] *= scale;
                    if (iter > 0) {
                        double dist = 0;
                        const float *oc = old_centers + 3 * k;
                        for (int j = 0; j < 3; j++) { double t = c[j] - oc[j]; dist += t * t; }
                        if (dist > max_center_shift) max_center_shift = dist;
                    }
                }
            }
            ++iter;
            int last = (iter == (maxCount > 2 ? maxCount : 2)) || max_center_shift <= eps2;
            if (last) {
                for (int i = 0; i < N; i++) dists[i] = l2sqr3(data + 3 * (size_t)i, centers + 3 * labels[i]);
                compactness = 0;
                for (int i = 0; i < N; i++) compactness += dists[i];
                if (iters_out) iters_out[a] = iter;
                break;
            }
            for (int i = 0; i < N; i++) {
                const float *sample = data + 3 * (size_t)i;
                int k_best = 0;
                double min_dist = DBL_MAX;
                for (int k = 0; k < K; k++) {
                    double d = l2sqr3(sample, centers + 3 * k);
                    if (min_dist > d) { min_dist = d; k_best = k; }
                }
                labels[i] = k_best;
            }
        }
        if (compactness < best_compactness) {
            best_compactness = compactness;
            memcpy(best_centers, centers, sizeof(float) * 3 * K);
            memcpy(best_labels, labels, sizeof(int32_t) * (size_t)N);
        }
    }
    free(buf); free(labels); free(dists);
    *rng_state = rng.state;
    return best_compactness;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Geometry: PosGrid.__init__ (stonesfinder.py:964-981), getrect(r, c, cursor=1.0) (:412-450), getmask (:452-493).
 * The float -> int16 / int conversions truncate toward zero exactly as numpy's assignment / Python's int().
 * ---------------------------------------------------------------------------------------------------------------- */
static void posgrid(int gsize, int size, int16_t *mtx /*[g][g][2]*/)
{
    double start = (double)size / gsize / 2, end = size - start;
    double hull[4][2] = { { start, start }, { end, start }, { end, end }, { start, end } };
    for (int i = 0; i < gsize; i++) {
        double xup = (hull[0][0] * (gsize - 1 - i) + hull[1][0] * i) / (gsize - 1);
        double xdown = (hull[3][0] * (gsize - 1 - i) + hull[2][0] * i) / (gsize - 1);
        for (int j = 0; j < gsize; j++) {
            mtx[(i * gsize + j) * 2 + 0] = (int16_t)((xup * (gsize - 1 - j) + xdown * j) / (gsize - 1));
            double yleft = (hull[0][1] * (gsize - 1 - j) + hull[3][1] * j) / (gsize - 1);
            double yright = (hull[1][1] * (gsize - 1 - j) + hull[2][1] * j) / (gsize - 1);
            mtx[(i * gsize + j) * 2 + 1] = (int16_t)((yleft * (gsize - 1 - i) + yright * i) / (gsize - 1));
        }
    }
}

/* rects[(r*g+c)*4 + {0,1,2,3}] = x0, y0, x1, y1 (x = image row, y = image column, as in the reference) */
CKO_API void cko_zone_rects(int gsize, int32_t *rects)
{
    int size = 20 * gsize; /* cvconf.py:10 */
    int16_t *m = (int16_t *)malloc(sizeof(int16_t) * 2 * gsize * gsize);
    posgrid(gsize, size, m);
#define P(r, c, k) m[(((r) + gsize) % gsize * gsize + ((c) + gsize) % gsize) * 2 + (k)] /* mtx[r-1] wraps at r=0 */
    for (int r = 0; r < gsize; r++)
        for (int c = 0; c < gsize; c++) {
            int p0 = P(r, c, 0), p1 = P(r, c, 1);
            int rb = r - 1, cb = c - 1;
            int ra = r + 1 < gsize - 1 ? r + 1 : gsize - 1, ca = c + 1 < gsize - 1 ? c + 1 : gsize - 1;
            /* int16 arithmetic in numpy for the assignments below; values are far from overflow */
            int pb0 = P(rb, cb, 0), pb1 = P(rb, cb, 1), pa0 = P(ra, ca, 0), pa1 = P(ra, ca, 1);
            if (r == 0) pb0 = -p0; else if (r == gsize - 1) pa0 = 2 * size - p0 - 2;
            if (c == 0) pb1 = -p1; else if (c == gsize - 1) pa1 = 2 * size - p1 - 2;
            double w = 1.0 / 2;
            int x0 = (int)(w * pb0 + (1 - w) * p0), y0 = (int)(w * pb1 + (1 - w) * p1);
            int x1 = (int)((1 - w) * p0 + w * pa0), y1 = (int)((1 - w) * p1 + w * pa1);
            int32_t *o = rects + (r * gsize + c) * 4;
            o[0] = x0 < 0 ? 0 : x0; o[1] = y0 < 0 ? 0 : y0;
            o[2] = x1 > size ? size : x1; o[3] = y1 > size ? size : y1;
        }
#undef P
    free(m);
}

/* mask[size*size] 0/1; returns zone_area = mask sum over zone (0,0). Pixels not covered by any zone stay 0 here
 * (the reference leaves them uninitialised: np.empty, stonesfinder.py:468 — they are never read through a zone). */
CKO_API int cko_zone_mask(int gsize, uint8_t *mask)
{
    int size = 20 * gsize;
    int32_t *rects = (int32_t *)malloc(sizeof(int32_t) * 4 * gsize * gsize);
    cko_zone_rects(gsize, rects);
    memset(mask, 0, (size_t)size * size);
    for (int r = 0; r < gsize; r++)
        for (int c = 0; c < gsize; c++) {
            int32_t *q = rects + (r * gsize + c) * 4;
            int h = q[2] - q[0], w = q[3] - q[1];
            double a = h / 2.0, b = w / 2.0, rad = a < b ? a : b;
            for (int i = 0; i < h; i++)
                for (int j = 0; j < w; j++) {
                    double y = -a + i, x = -b + j;
                    mask[(size_t)(q[0] + i) * size + q[1] + j] = (x * x + y * y <= rad * rad);
                }
        }
    int area = 0;
    for (int i = rects[0]; i < rects[2]; i++)
        for (int j = rects[1]; j < rects[3]; j++) area += mask[(size_t)i * size + j];
    free(rects);
    return area;
}

/* ------------------------------------------------------------------------------------------------------------------
 * SfClustering.cluster_colors part 2 + interpret_ratios + check_density           [sf_clustering.py:105-178]
 * labels: h*w int32 k-means labels (0..2) of the sub-image [x0:x1, y0:y1]; centers 3x3 f32.
 * ratios_out: g*g*3 uint8; stones_out: g*g uint8 codes (0=E,1=B,2=W). returns trusted (check_density).
 * ---------------------------------------------------------------------------------------------------------------- */
static int center_grey(const float *c)
{
    /* int(sum(x) / 3): Python sum() over numpy float32 scalars starts from int 0 and stays float32; /3 -> float32
       (NEP-50 weak Python int), then truncation toward zero */
    float s = 0.0f;
    s = s + c[0]; s = s + c[1]; s = s + c[2];
    float q = s / 3.0f;
    return (int)q;
}

CKO_API int cko_zone_classify(const int32_t *labels, const float *centers, int gsize, int rs, int re, int cs, int ce,
                              uint8_t *ratios_out, uint8_t *stones_out)
{
    int size = 20 * gsize;
    int32_t *rects = (int32_t *)malloc(sizeof(int32_t) * 4 * gsize * gsize);
    uint8_t *mask = (uint8_t *)malloc((size_t)size * size);
    cko_zone_rects(gsize, rects);
    cko_zone_mask(gsize, mask);
    int x0 = rects[(rs * gsize + cs) * 4 + 0], y0 = rects[(rs * gsize + cs) * 4 + 1];
    int y1 = rects[((re - 1) * gsize + (ce - 1)) * 4 + 3];
    int w = y1 - y0;
    int cv[3] = { center_grey(centers), center_grey(centers + 3), center_grey(centers + 6) };
    /* centers_val.index(sorted(centers_val)[1]) — first index holding the median value */
    int sorted[3] = { cv[0], cv[1], cv[2] };
    for (int i = 0; i < 3; i++) for (int j = i + 1; j < 3; j++) if (sorted[j] < sorted[i]) { int t = sorted[i]; sorted[i] = sorted[j]; sorted[j] = t; }
    int mid = 0;
    while (cv[mid] != sorted[1]) mid++;
    memset(ratios_out, 0, (size_t)gsize * gsize * 3);
    for (int i = 0; i < gsize * gsize; i++) ratios_out[i * 3 + mid] = 1;
    for (int x = rs; x < re; x++)
        for (int y = cs; y < ce; y++) {
            int32_t *q = rects + (x * gsize + y) * 4;
            int cnt[3] = { 0, 0, 0 }, total = 0;
            for (int i = q[0]; i < q[2]; i++)
                for (int j = q[1]; j < q[3]; j++) {
                    total++;
                    if (mask[(size_t)i * size + j]) cnt[labels[(size_t)(i - x0) * w + (j - y0)]]++;
                }
            /* ratios[x][y][label-1] = 100 * counts[i] / sum(counts): float64 true division, truncating cast to uint8;
               only labels that are present are written */
            for (int k = 0; k < 3; k++)
                if (cnt[k]) ratios_out[(x * gsize + y) * 3 + k] = (uint8_t)(int)(100.0 * cnt[k] / (double)total);
        }
    /* interpret_ratios: grey == min -> B, elif grey == max -> W, else E; argmax (first max wins) */
    int mn = sorted[0], mx = sorted[2];
    uint8_t col[3];
    for (int k = 0; k < 3; k++) col[k] = cv[k] == mn ? 1 : (cv[k] == mx ? 2 : 0);
    memset(stones_out, 0, (size_t)gsize * gsize);
    for (int i = rs; i < re; i++)
        for (int j = cs; j < ce; j++) {
            const uint8_t *r3 = ratios_out + (i * gsize + j) * 3;
            int best = 0;
            if (r3[1] > r3[best]) best = 1;
            if (r3[2] > r3[best]) best = 2;
            stones_out[i * gsize + j] = col[best];
        }
    int hist[3] = { 0, 0, 0 };
    for (int i = 0; i < gsize * gsize; i++) hist[stones_out[i]]++;
    free(rects); free(mask);
    return hist[0] >= 2 && hist[1] >= 2 && hist[2] >= 2;
}

/* ------------------------------------------------------------------------------------------------------------------
 * SfNeural CNN (NNManager.create_net, nn_manager.py:277-298), forward only, float32:
 *   Conv 5x5x32 valid + ReLU -> Conv 5x5x32 + ReLU -> MaxPool 2 -> Conv 3x3x90 + ReLU -> Conv 3x3x90 + ReLU ->
 *   MaxPool 2 -> Flatten (H,W,C) -> Dense 160 + ReLU -> Dense 81 + softmax.   Dropout is identity at inference.
 * Input is the raw uint8 patch cast to float32 with NO scaling (nn_cache.py:49-50, nn_manager.py:327).
 * Weights in Keras channels-last layout: conv (kh, kw, cin, cout), dense (in, out). Convolutions are computed as
 * cross-correlations over these arrays (a Theano true-convolution model would need its kernels flipped at load).
 * `acc` selects the accumulator type: 0 = float32 (the parity oracle), 1 = float64 (ground truth for error budgets).
 * ---------------------------------------------------------------------------------------------------------------- */
#define CKO_CNN_NPARAM 658665

typedef struct {
    const float *w1, *b1, *w2, *b2, *w3, *b3, *w4, *b4, *w5, *b5, *w6, *b6;
} cnn_w;

static cnn_w cnn_split(const float *p)
{
    cnn_w w;
    w.w1 = p; p += 5 * 5 * 3 * 32;   w.b1 = p; p += 32;
    w.w2 = p; p += 5 * 5 * 32 * 32;  w.b2 = p; p += 32;
    w.w3 = p; p += 3 * 3 * 32 * 90;  w.b3 = p; p += 90;
    w.w4 = p; p += 3 * 3 * 90 * 90;  w.b4 = p; p += 90;
    w.w5 = p; p += 3240 * 160;       w.b5 = p; p += 160;
    w.w6 = p; p += 160 * 81;         w.b6 = p; p += 81;
    return w;
}

#define DEF_CONV(NAME, ACC)                                                                                          \
    static void NAME(const float *in, int ih, int iw, int cin, const float *wt, const float *bias, int kh, int kw,   \
                     int cout, float *out)                                                                           \
    {                                                                                                                \
        int oh = ih - kh + 1, ow = iw - kw + 1;                                                                      \
        ACC acc[96];                                                                                                 \
        for (int y = 0; y < oh; y++)                                                                                 \
            for (int x = 0; x < ow; x++) {                                                                           \
                for (int o = 0; o < cout; o++) acc[o] = 0;                                                           \
                for (int dy = 0; dy < kh; dy++)                                                                      \
                    for (int dx = 0; dx < kw; dx++) {                                                                \
                        const float *ip = in + ((size_t)(y + dy) * iw + (x + dx)) * cin;                             \
                        const float *wp = wt + (size_t)(dy * kw + dx) * cin * cout;                                  \
                        for (int c = 0; c < cin; c++) {                                                              \
                            ACC v = ip[c];                                                                           \
                            const float *wr = wp + (size_t)c * cout;                                                 \
                            for (int o = 0; o < cout; o++) acc[o] += v * (ACC)wr[o];                                 \
                        }                                                                                            \
                    }                                                                                                \
                float *op = out + ((size_t)y * ow + x) * cout;                                                       \
                for (int o = 0; o < cout; o++) {                                                                     \
                    ACC v = acc[o] + (ACC)bias[o];                                                                   \
                    op[o] = v > 0 ? (float)v : 0.f;                                                                  \
                }                                                                                                    \
            }                                                                                                        \
    }
DEF_CONV(conv_relu_f32, float)
DEF_CONV(conv_relu_f64, double)

static void maxpool2(const float *in, int ih, int iw, int c, float *out)
{
    int oh = ih / 2, ow = iw / 2;
    for (int y = 0; y < oh; y++)
        for (int x = 0; x < ow; x++)
            for (int k = 0; k < c; k++) {
                float a = in[((size_t)(2 * y) * iw + 2 * x) * c + k], b = in[((size_t)(2 * y) * iw + 2 * x + 1) * c + k];
                float d = in[((size_t)(2 * y + 1) * iw + 2 * x) * c + k], e = in[((size_t)(2 * y + 1) * iw + 2 * x + 1) * c + k];
                float m = a > b ? a : b, n = d > e ? d : e;
                out[((size_t)y * ow + x) * c + k] = m > n ? m : n;
            }
}

#define DEF_DENSE(NAME, ACC)                                                                                         \
    static void NAME(const float *in, int n_in, const float *wt, const float *bias, int n_out, float *out, int relu, \
                     double *out64)                                                                                  \
    {                                                                                                                \
        ACC acc[160];                                                                                                \
        for (int o = 0; o < n_out; o++) acc[o] = 0;                                                                  \
        for (int i = 0; i < n_in; i++) {                                                                             \
            ACC v = in[i];                                                                                           \
            const float *wr = wt + (size_t)i * n_out;                                                                \
            for (int o = 0; o < n_out; o++) acc[o] += v * (ACC)wr[o];                                                \
        }                                                                                                            \
        for (int o = 0; o < n_out; o++) {                                                                            \
            ACC v = acc[o] + (ACC)bias[o];                                                                           \
            if (out64) out64[o] = (double)v;                                                                         \
            out[o] = relu ? (v > 0 ? (float)v : 0.f) : (float)v;                                                     \
        }                                                                                                            \
    }
DEF_DENSE(dense_f32, float)
DEF_DENSE(dense_f64, double)

/* x: n patches of 40x40x3 uint8 (HWC); y: n x 81 softmax; logits (optional): n x 81 pre-softmax.
 * acts (optional, debugging aid): per patch the post-ReLU outputs of conv1 (36*36*32), conv2 pooled (16*16*32),
 * conv3 (14*14*90), conv4 pooled (6*6*90), fc1 (160), concatenated (CKO_ACTS_PER_PATCH floats). */
#define CKO_ACTS_PER_PATCH (36 * 36 * 32 + 16 * 16 * 32 + 14 * 14 * 90 + 6 * 6 * 90 + 160)
CKO_API int cko_acts_per_patch(void) { return CKO_ACTS_PER_PATCH; }
CKO_API int cko_cnn_nparam(void) { return CKO_CNN_NPARAM; }

CKO_API void cko_cnn_forward(const uint8_t *x, int n, const float *params, int acc, float *y, float *logits,
                             float *acts)
{
    cnn_w w = cnn_split(params);
#pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < n; p++) {
        float *in = (float *)malloc(sizeof(float) * 40 * 40 * 3);
        float *a1 = (float *)malloc(sizeof(float) * 36 * 36 * 32);
        float *a2 = (float *)malloc(sizeof(float) * 32 * 32 * 32);
        float *p2 = (float *)malloc(sizeof(float) * 16 * 16 * 32);
        float *a3 = (float *)malloc(sizeof(float) * 14 * 14 * 90);
        float *a4 = (float *)malloc(sizeof(float) * 12 * 12 * 90);
        float *p4 = (float *)malloc(sizeof(float) * 6 * 6 * 90);
        float f5[160], f6[81];
        double l64[81];
        for (int i = 0; i < 40 * 40 * 3; i++) in[i] = (float)x[(size_t)p * 4800 + i];
        if (acc == 0) {
            conv_relu_f32(in, 40, 40, 3, w.w1, w.b1, 5, 5, 32, a1);
            conv_relu_f32(a1, 36, 36, 32, w.w2, w.b2, 5, 5, 32, a2);
            maxpool2(a2, 32, 32, 32, p2);
            conv_relu_f32(p2, 16, 16, 32, w.w3, w.b3, 3, 3, 90, a3);
            conv_relu_f32(a3, 14, 14, 90, w.w4, w.b4, 3, 3, 90, a4);
            maxpool2(a4, 12, 12, 90, p4);
            dense_f32(p4, 3240, w.w5, w.b5, 160, f5, 1, NULL);
            dense_f32(f5, 160, w.w6, w.b6, 81, f6, 0, l64);
        } else {
            conv_relu_f64(in, 40, 40, 3, w.w1, w.b1, 5, 5, 32, a1);
            conv_relu_f64(a1, 36, 36, 32, w.w2, w.b2, 5, 5, 32, a2);
            maxpool2(a2, 32, 32, 32, p2);
            conv_relu_f64(p2, 16, 16, 32, w.w3, w.b3, 3, 3, 90, a3);
            conv_relu_f64(a3, 14, 14, 90, w.w4, w.b4, 3, 3, 90, a4);
            maxpool2(a4, 12, 12, 90, p4);
            dense_f64(p4, 3240, w.w5, w.b5, 160, f5, 1, NULL);
            dense_f64(f5, 160, w.w6, w.b6, 81, f6, 0, l64);
        }
        if (logits) memcpy(logits + (size_t)p * 81, f6, sizeof f6);
        /* softmax: exp(z - max) / sum, in the accumulator precision */
        if (acc == 0) {
            float m = f6[0];
            for (int o = 1; o < 81; o++) if (f6[o] > m) m = f6[o];
            float e[81], s = 0.f;
            for (int o = 0; o < 81; o++) { e[o] = expf(f6[o] - m); s += e[o]; }
            for (int o = 0; o < 81; o++) y[(size_t)p * 81 + o] = e[o] / s;
        } else {
            double m = l64[0];
            for (int o = 1; o < 81; o++) if (l64[o] > m) m = l64[o];
            double e[81], s = 0;
            for (int o = 0; o < 81; o++) { e[o] = exp(l64[o] - m); s += e[o]; }
            for (int o = 0; o < 81; o++) y[(size_t)p * 81 + o] = (float)(e[o] / s);
        }
        if (acts) {
            float *q = acts + (size_t)p * CKO_ACTS_PER_PATCH;
            memcpy(q, a1, sizeof(float) * 36 * 36 * 32); q += 36 * 36 * 32;
            memcpy(q, p2, sizeof(float) * 16 * 16 * 32); q += 16 * 16 * 32;
            memcpy(q, a3, sizeof(float) * 14 * 14 * 90); q += 14 * 14 * 90;
            memcpy(q, p4, sizeof(float) * 6 * 6 * 90);   q += 6 * 6 * 90;
            memcpy(q, f5, sizeof(float) * 160);
        }
        free(in); free(a1); free(a2); free(p2); free(a3); free(a4); free(p4);
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Patch geometry (NNManager._subregion / getrect / _get_rect_nn / generate_xs, nn_manager.py:92-126,216-225,256-275)
 * and the decode of nn_cache.py:25-41 + the 0.6 confidence rule of sf_neural.py:57-70 (19x19 only).
 * ---------------------------------------------------------------------------------------------------------------- */
static void nn_subregion(int i, int j, int *rs, int *re, int *cs, int *ce)
{
    const int gsize = 19, step = 2;
    *rs = i * step; *re = (i + 1) * step;
    if (gsize - *rs < step) { *rs = gsize - step; *re = gsize; }
    *cs = j * step; *ce = (j + 1) * step;
    if (gsize - *cs < step) { *cs = gsize - step; *ce = gsize; }
}

CKO_API void cko_nn_patch_origin(int i, int j, int *x0, int *y0)
{
    int rs, re, cs, ce;
    nn_subregion(i, j, &rs, &re, &cs, &ce);
    /* getrect(r, c): x0 = int(r*380/19); x1 of getrect(re-1, ce-1) = int(re*380/19); width forced to 40 from x1 */
    int x1 = (int)(re * 380.0 / 19), y1 = (int)(ce * 380.0 / 19);
    int xa = (int)(rs * 380.0 / 19), ya = (int)(cs * 380.0 / 19);
    if (x1 - xa != 40) xa = x1 - 40;
    if (y1 - ya != 40) ya = y1 - 40;
    *x0 = xa; *y0 = ya;
}

CKO_API void cko_nn_gather(const uint8_t *goban /*380x380x3*/, uint8_t *xs /*100x40x40x3*/)
{
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 10; j++) {
            int x0, y0;
            cko_nn_patch_origin(i, j, &x0, &y0);
            for (int r = 0; r < 40; r++)
                memcpy(xs + ((size_t)(i * 10 + j) * 40 + r) * 120, goban + ((size_t)(x0 + r) * 380 + y0) * 3, 120);
        }
}

/* compute_stones(label): base-3 digits, little-endian (nn_manager.py:246-254) */
CKO_API void cko_nn_compute_stones(int label, uint8_t *four)
{
    int k = label;
    static const int p3[4] = { 1, 3, 9, 27 };
    for (int i = 3; i >= 0; i--) { four[i] = (uint8_t)(k / p3[i]); k %= p3[i]; }
}

/* y: 100x81 softmax. stones_out[361] codes, conf_out[361]; keep[361] = (stone != E && conf > 0.6) */
CKO_API void cko_nn_decode(const float *y, uint8_t *stones_out, float *conf_out, uint8_t *keep)
{
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 10; j++) {
            const float *v = y + (size_t)(i * 10 + j) * 81;
            int best = 0;
            float s = 0.f; /* Python sum(): sequential float32 adds starting from int 0 */
            for (int o = 0; o < 81; o++) { if (v[o] > v[best]) best = o; s = s + v[o]; }
            float conf = v[best] / s;
            uint8_t four[4];
            cko_nn_compute_stones(best, four);
            int rs, re, cs, ce;
            nn_subregion(i, j, &rs, &re, &cs, &ce);
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++) {
                    stones_out[(rs + a) * 19 + cs + b] = four[a * 2 + b];
                    conf_out[(rs + a) * 19 + cs + b] = conf;
                }
        }
    for (int k = 0; k < 361; k++) keep[k] = stones_out[k] != 0 && conf_out[k] > 0.6f;
}

/* ------------------------------------------------------------------------------------------------------------------
 * cv2.createBackgroundSubtractorMOG2(detectShadows=False).apply(img 8UC3, learningRate=lr)     [stonesfinder.py:113-115,
 * 171-176: StonesFinder.__init__ creates the model, _learn_bg applies every canonical frame with lr = 0.01 during the
 * first bg_init_frames frames and 0.005 afterwards]
 *
 * Zivkovic's adaptive Gaussian mixture (OpenCV video/bgfg_gaussmix2.cpp, MOG2Invoker), defaults: history 500,
 * nmixtures 5, varThreshold (Tb) 16, varThresholdGen (Tg) 9, backgroundRatio (TB) 0.9, varInit 15, varMin 4,
 * varMax 75, complexity reduction CT 0.05. Per pixel, modes sorted by weight (descending):
 *   frame counter n (1-based): alpha = lr if n > 1 and lr >= 0, else 1 / min(2 n, history)  (so the first frame
 *   always uses 0.5); prune = -alpha * CT; all arithmetic float32 without FMA contraction.
 * State: weight[5], variance[5], mean[5][3], nmodes, laid out here per pixel as 25 floats + 1 byte.
 * Pinned bit-exactly against cv2 4.13.0 (tests/test_oracle_vs_cv2.py::test_mog2_bit_exact).
 * ---------------------------------------------------------------------------------------------------------------- */
#define MOG2_NMIX 5
CKO_API void cko_mog2_apply(const uint8_t *img /*[npix][3]*/, int npix, float *state /*[npix][25]*/, uint8_t *nmodes_io,
                            int frame_no /*1-based*/, double learning_rate, uint8_t *mask)
{
    const int history = 500;
    const float Tb = 16.f, Tg = 9.f, TB = 0.9f, varInit = 15.f, varMin = 4.f, varMax = 75.f, CT = 0.05f;
    const double lrd = (learning_rate >= 0 && frame_no > 1) ? learning_rate
                                                            : 1. / (2 * frame_no < history ? 2 * frame_no : history);
    const float alphaT = (float)lrd;
    const float alpha1 = 1.f - alphaT;
    const float prune = (float)(-lrd * CT);
    for (int p = 0; p < npix; p++) {
        float *w = state + (size_t)p * 25, *var = w + 5, *mean = w + 10;
        const float data[3] = {(float)img[p * 3], (float)img[p * 3 + 1], (float)img[p * 3 + 2]};
        int background = 0, fits = 0, nmodes = nmodes_io[p];
        float totalWeight = 0.f;
        for (int mode = 0; mode < nmodes; mode++) {   /* the bound shrinks as modes are pruned, as in OpenCV */
            float weight = alpha1 * w[mode] + prune;
            int swap_count = 0;
            if (!fits) {
                const float v = var[mode];
                float d[3];
                d[0] = mean[mode * 3] - data[0];
                d[1] = mean[mode * 3 + 1] - data[1];
                d[2] = mean[mode * 3 + 2] - data[2];
                const float dist2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                if (totalWeight < TB && dist2 < Tb * v) background = 1;
                if (dist2 < Tg * v) {
                    fits = 1;
                    weight += alphaT;
                    const float k = alphaT / weight;
                    for (int c = 0; c < 3; c++) mean[mode * 3 + c] -= k * d[c];
                    float varnew = v + k * (dist2 - v);
                    varnew = varnew > varMin ? varnew : varMin;
                    varnew = varnew < varMax ? varnew : varMax;
                    var[mode] = varnew;
                    for (int i = mode; i > 0; i--) {
                        if (weight < w[i - 1]) break;
                        swap_count++;
                        float t;
                        t = w[i]; w[i] = w[i - 1]; w[i - 1] = t;
                        t = var[i]; var[i] = var[i - 1]; var[i - 1] = t;
                        for (int c = 0; c < 3; c++) {
                            t = mean[i * 3 + c]; mean[i * 3 + c] = mean[(i - 1) * 3 + c]; mean[(i - 1) * 3 + c] = t;
                        }
                    }
                }
            }
            if (weight < -prune) {
                weight = 0.f;
                nmodes--;
            }
            w[mode - swap_count] = weight;
            totalWeight += weight;
        }
        float invWeight = 0.f;
        if (fabsf(totalWeight) > FLT_EPSILON) invWeight = 1.f / totalWeight;
        for (int mode = 0; mode < nmodes; mode++) w[mode] *= invWeight;
        if (!fits && alphaT > 0.f) {
            const int mode = nmodes == MOG2_NMIX ? MOG2_NMIX - 1 : nmodes++;
            if (nmodes == 1)
                w[mode] = 1.f;
            else {
                w[mode] = alphaT;
                for (int i = 0; i < nmodes - 1; i++) w[i] *= alpha1;
            }
            for (int c = 0; c < 3; c++) mean[mode * 3 + c] = data[c];
            var[mode] = varInit;
            for (int i = nmodes - 1; i > 0; i--) {
                if (alphaT < w[i - 1]) break;
                float t;
                t = w[i]; w[i] = w[i - 1]; w[i - 1] = t;
                t = var[i]; var[i] = var[i - 1]; var[i - 1] = t;
                for (int c = 0; c < 3; c++) {
                    t = mean[i * 3 + c]; mean[i * 3 + c] = mean[(i - 1) * 3 + c]; mean[(i - 1) * 3 + c] = t;
                }
            }
        }
        nmodes_io[p] = (uint8_t)nmodes;
        mask[p] = background ? 0 : 255;
    }
}
