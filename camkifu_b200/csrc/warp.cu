// K1 — batched perspective warp to the canonical goban image, bit-identical to OpenCV's fixed-point remap, and the
// running-average update that follows it in SfClustering.
//
// Replaces cv2.warpPerspective(frame, transform, canonical_shape) (src/camkifu/stone/stonesfinder.py:140) and
// cv2.accumulateWeighted(gframe, self.accu, 0.2) (src/camkifu/stone/sf_clustering.py:33-36).
//
// OpenCV (INTER_LINEAR, BORDER_CONSTANT 0, 8UC3) maps every destination pixel through the inverse homography in
// float64, per 64-pixel destination block:  X0 = Mi0*bx + Mi1*y + Mi2,  W = W0 + Mi6*x1,  W = W ? 32/W : 0,
// X = rint((X0 + Mi0*x1) * W)  — so source coordinates are quantised to 1/32 pixel — and blends the four taps with
// 10-bit integer weights, (sum + 512) >> 10. The float64 expressions below use explicit round-to-nearest intrinsics in
// OpenCV's operation order so that nvcc cannot contract them into FMAs; everything after the rint is integer.
//
// Mapping: one thread produces 4 consecutive destination pixels of one row (12 output bytes = three aligned 32-bit
// stores; a warp writes 384 contiguous bytes). Source taps are 2 rows x 6 contiguous bytes per pixel, fetched as
// aligned 32-bit words through the read-only path and realigned with funnel shifts. The source footprint of a warp is
// a short run of two image rows, so neighbouring threads share 32-byte sectors; the kernel is a gather bounded by
// sectors touched in HBM/L2.
#include "ckb_common.cuh"

#define CKB_WARP_CHUNK 32  // frames per launch: their inverse homographies travel by value in the kernel parameters

struct WarpMats {
    double m[CKB_WARP_CHUNK][9];
};

__device__ __forceinline__ int coord_to_fixed(double num, double w)
{
    double v = __dmul_rn(num, w);
    v = (v < 2147483647.0) ? v : 2147483647.0;    // std::min((double)INT_MAX, v)
    v = (-2147483648.0 < v) ? v : -2147483648.0;  // std::max((double)INT_MIN, v)
    return __double2int_rn(v);                    // cvRound: nearest, ties to even
}

__device__ __forceinline__ int sat_s16(int v) { return max(-32768, min(32767, v)); }

// 6 consecutive bytes at an arbitrary address: p[0..3] -> lo, p[4..5] -> low half of hi
__device__ __forceinline__ void load6(const uint8_t *p, uint32_t &lo, uint32_t &hi)
{
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3);
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1);
    const uint32_t w2 = (sh == 3) ? __ldg(q + 2) : 0u;
    lo = __funnelshift_r(w0, w1, sh * 8);
    hi = __funnelshift_r(w1, w2, sh * 8);
}

// taps of one source row for pixel columns sx and sx+1 (BGR each); zero outside the image
__device__ __forceinline__ void load_row_taps(const uint8_t *frame, size_t row_pitch, int H, int W, int sy, int sx,
                                              uint32_t &lo, uint32_t &hi)
{
    lo = 0;
    hi = 0;
    if ((unsigned)sy >= (unsigned)H) return;
    const uint8_t *row = frame + (size_t)sy * row_pitch;
    // load6 reads the aligned words around the 6 bytes it needs (up to 3 bytes before and after them): not where that
    // window could leave the frame (its first bytes, the end of its last row)
    if (sx >= 0 && sx + 1 < W && (sy > 0 || sx > 0) && (sy + 1 < H || sx + 3 <= W)) {
        load6(row + (size_t)sx * 3, lo, hi);
        hi &= 0xffffu;
    } else {
        if ((unsigned)sx < (unsigned)W) {
            const uint8_t *p = row + (size_t)sx * 3;
            lo = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
        }
        if ((unsigned)(sx + 1) < (unsigned)W) {
            const uint8_t *p = row + (size_t)(sx + 1) * 3;
            lo |= (uint32_t)__ldg(p) << 24;
            hi = (uint32_t)__ldg(p + 1) | ((uint32_t)__ldg(p + 2) << 8);
        }
    }
}

__global__ void __launch_bounds__(128) ckb_warp_kernel(const uint8_t *__restrict__ frames, int H, int W,
                                                       size_t row_pitch, size_t frame_pitch,
                                                       const __grid_constant__ WarpMats mats, int per_frame,
                                                       uint8_t *__restrict__ goban, int S)
{
    const int quads_per_row = S >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= quads_per_row * S) return;
    const int f = blockIdx.y;
    const int y = idx / quads_per_row;
    const int x = (idx - y * quads_per_row) << 2;
    const double *Mi = mats.m[per_frame ? f : 0];
    const uint8_t *frame = frames + (size_t)f * frame_pitch;

    const int bx = x & ~63;  // OpenCV evaluates the row terms at the origin of each 64-wide block
    const int x1 = x - bx;
    const double dy = (double)y, dbx = (double)bx;
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(Mi[0], dbx), __dmul_rn(Mi[1], dy)), Mi[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(Mi[3], dbx), __dmul_rn(Mi[4], dy)), Mi[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(Mi[6], dbx), __dmul_rn(Mi[7], dy)), Mi[8]);

    uint32_t out[3] = {0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double dx1 = (double)(x1 + k);
        double Wd = __dadd_rn(W0, __dmul_rn(Mi[6], dx1));
        Wd = (Wd != 0.0) ? __ddiv_rn(32.0, Wd) : 0.0;
        const int X = coord_to_fixed(__dadd_rn(X0, __dmul_rn(Mi[0], dx1)), Wd);
        const int Y = coord_to_fixed(__dadd_rn(Y0, __dmul_rn(Mi[3], dx1)), Wd);
        const int sx = sat_s16(X >> 5), sy = sat_s16(Y >> 5);
        const int fx = X & 31, fy = Y & 31;
        const int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy), w10 = (32 - fx) * fy, w11 = fx * fy;
        uint32_t a_lo, a_hi, b_lo, b_hi;
        load_row_taps(frame, row_pitch, H, W, sy, sx, a_lo, a_hi);
        load_row_taps(frame, row_pitch, H, W, sy + 1, sx, b_lo, b_hi);
        // bytes: lo = [B0 G0 R0 B1], hi = [G1 R1]
        const int t00b = a_lo & 0xff, t00g = (a_lo >> 8) & 0xff, t00r = (a_lo >> 16) & 0xff;
        const int t01b = a_lo >> 24, t01g = a_hi & 0xff, t01r = (a_hi >> 8) & 0xff;
        const int t10b = b_lo & 0xff, t10g = (b_lo >> 8) & 0xff, t10r = (b_lo >> 16) & 0xff;
        const int t11b = b_lo >> 24, t11g = b_hi & 0xff, t11r = (b_hi >> 8) & 0xff;
        const uint32_t vb = (uint32_t)(t00b * w00 + t01b * w01 + t10b * w10 + t11b * w11 + 512) >> 10;
        const uint32_t vg = (uint32_t)(t00g * w00 + t01g * w01 + t10g * w10 + t11g * w11 + 512) >> 10;
        const uint32_t vr = (uint32_t)(t00r * w00 + t01r * w01 + t10r * w10 + t11r * w11 + 512) >> 10;
        // byte position of this pixel's B inside the 12-byte group: 3k
        const uint32_t px = vb | (vg << 8) | (vr << 16);
        const int bpos = 3 * k;
        out[bpos >> 2] |= px << ((bpos & 3) * 8);
        if ((bpos & 3) > 1) out[(bpos >> 2) + 1] |= px >> (32 - (bpos & 3) * 8);
    }
    uint32_t *dst = (uint32_t *)(goban + ((size_t)f * S * S + (size_t)y * S + x) * 3);
    dst[0] = out[0];
    dst[1] = out[1];
    dst[2] = out[2];
}

int ckb_launch_warp(ckb_ctx *ctx, const uint8_t *d_frames, int n, int H, int W, size_t row_pitch, size_t frame_pitch,
                    const double *h_minv, int n_mtx, uint8_t *d_goban, cudaStream_t st)
{
    const int S = ctx->S;
    const int threads = 128;
    const int work = (S >> 2) * S;
    for (int f0 = 0; f0 < n; f0 += CKB_WARP_CHUNK) {
        const int nf = n - f0 < CKB_WARP_CHUNK ? n - f0 : CKB_WARP_CHUNK;
        WarpMats mats;
        if (n_mtx == 1)
            memcpy(mats.m[0], h_minv, 9 * sizeof(double));
        else
            memcpy(mats.m[0], h_minv + (size_t)f0 * 9, (size_t)nf * 9 * sizeof(double));
        dim3 grid((work + threads - 1) / threads, nf);
        ckb_warp_kernel<<<grid, threads, 0, st>>>(d_frames + (size_t)f0 * frame_pitch, H, W, row_pitch, frame_pitch,
                                                  mats, n_mtx != 1, d_goban + (size_t)f0 * S * S * 3, S);
        CKB_LAUNCH_CHECK(ctx, "ckb_warp_kernel");
    }
    return CKB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Running average. One thread owns 4 consecutive float32 elements of the S*S*3 state and walks the n frames in order
// (the recurrence is sequential per element, parallel across elements): accu = fma(alpha, src - accu, accu), which is
// what OpenCV's optimised accumulateWeighted evaluates (bit-exact, tests/test_oracle_vs_cv2.py).
__global__ void __launch_bounds__(256) ckb_accumulate_kernel(const uint8_t *__restrict__ goban, int n, int n4,
                                                             float *__restrict__ accu, float alpha, int first,
                                                             float *__restrict__ snap, int snap_every, int snap_phase)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const size_t img_words = (size_t)n4;
    float4 a;
    int f = 0;
    int snaps = 0;
    if (first) {
        const uint32_t s = __ldg((const uint32_t *)goban + i);
        a = make_float4((float)(s & 0xff), (float)((s >> 8) & 0xff), (float)((s >> 16) & 0xff), (float)(s >> 24));
        if ((snap_phase % snap_every) == 0) {
            if (snap) ((float4 *)snap)[(size_t)snaps * img_words + i] = a;
            snaps++;
        }
        f = 1;
    } else {
        a = ((const float4 *)accu)[i];
    }
    for (; f < n; f++) {
        const uint32_t s = __ldg((const uint32_t *)goban + (size_t)f * img_words + i);
        a.x = __fmaf_rn(alpha, __fsub_rn((float)(s & 0xff), a.x), a.x);
        a.y = __fmaf_rn(alpha, __fsub_rn((float)((s >> 8) & 0xff), a.y), a.y);
        a.z = __fmaf_rn(alpha, __fsub_rn((float)((s >> 16) & 0xff), a.z), a.z);
        a.w = __fmaf_rn(alpha, __fsub_rn((float)(s >> 24), a.w), a.w);
        if (((f + snap_phase) % snap_every) == 0) {
            if (snap) ((float4 *)snap)[(size_t)snaps * img_words + i] = a;
            snaps++;
        }
    }
    ((float4 *)accu)[i] = a;
}

int ckb_launch_accumulate(ckb_ctx *ctx, const uint8_t *d_goban, int n, float *d_accu, float alpha, int first,
                          float *d_snap, int snap_every, int snap_phase, cudaStream_t st)
{
    const int n4 = ctx->S * ctx->S * 3 / 4;  // S*S*3 = 1200 * gsize^2 is a multiple of 4
    const int threads = 256;
    if (snap_every < 1) snap_every = 1;
    ckb_accumulate_kernel<<<(n4 + threads - 1) / threads, threads, 0, st>>>(d_goban, n, n4, d_accu, alpha, first,
                                                                           d_snap, snap_every, snap_phase);
    CKB_LAUNCH_CHECK(ctx, "ckb_accumulate_kernel");
    return CKB_OK;
}
