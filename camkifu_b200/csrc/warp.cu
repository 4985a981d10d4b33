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
#include <stdlib.h>

#include "ckb_common.cuh"
#include "tc_ptx.cuh"

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

// ---- variant with TMA-staged source tiles (north-star item 1; measured against the direct kernel) --------
// CTA = 64 x 8 destination pixels (one OpenCV 64-pixel block x 8 rows; 16 threads x 4 pixels per row). Every thread first
// computes the fixed-point source coordinates of its 4 pixels (same arithmetic as above), the CTA reduces them to the exact
// bounding box of all taps, one thread issues a 1-D bulk copy (cp.async.bulk, the TMA engine: UBLKCP) per source row of the
// box into shared memory — each row from its own 16-byte aligned start — and the taps are then read from shared memory.
// Tiles whose taps leave the image, or whose box does not fit, take the direct path (whole CTA, uniform).
#define WS_ROWS 8
#define WS_SMEM_BYTES (40 * 1024)
#define WS_MAX_SRC_ROWS 96

__device__ __forceinline__ void load6_smem(uint32_t addr, uint32_t &lo, uint32_t &hi)
{
    const uint32_t q = addr & ~3u, sh = addr & 3u;
    uint32_t w0, w1, w2 = 0u;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(q));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(q + 4));
    if (sh == 3) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w2) : "r"(q + 8));
    lo = __funnelshift_r(w0, w1, sh * 8);
    hi = __funnelshift_r(w1, w2, sh * 8) & 0xffffu;
}

__global__ void __launch_bounds__(128) ckb_warp_kernel_staged(const uint8_t *__restrict__ frames, int H, int W,
                                                              size_t row_pitch, size_t frame_pitch, size_t total_bytes,
                                                              const __grid_constant__ WarpMats mats, int per_frame,
                                                              uint8_t *__restrict__ goban, int S)
{
    extern __shared__ __align__(128) uint8_t ws_tile[];
    __shared__ int s_box[4];                 // sx min, sx max, sy min, sy max over the CTA's taps
    __shared__ int s_ok;
    __shared__ uint32_t s_rowoff[WS_MAX_SRC_ROWS];   // per staged row: byte offset of column sxmin inside the tile row
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    const int blocks_x = (S + 63) >> 6;
    const int bxi = blockIdx.x % blocks_x, byi = blockIdx.x / blocks_x;
    const int f = blockIdx.y;
    const int y = byi * WS_ROWS + (tid >> 4);
    const int x = (bxi << 6) + ((tid & 15) << 2);
    const bool active = y < S && x < S;      // S is a multiple of 4: an active thread owns 4 pixels
    const double *Mi = mats.m[per_frame ? f : 0];
    const uint8_t *frame = frames + (size_t)f * frame_pitch;
    if (tid == 0) {
        s_box[0] = INT_MAX; s_box[1] = INT_MIN; s_box[2] = INT_MAX; s_box[3] = INT_MIN;
        s_ok = 1;
        mbar_init(smem_u32(&s_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int X[4], Y[4];
    int sxmin = INT_MAX, sxmax = INT_MIN, symin = INT_MAX, symax = INT_MIN;
    bool inside = true;
    if (active) {
        const int bx = x & ~63, x1 = x - bx;
        const double dy = (double)y, dbx = (double)bx;
        const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(Mi[0], dbx), __dmul_rn(Mi[1], dy)), Mi[2]);
        const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(Mi[3], dbx), __dmul_rn(Mi[4], dy)), Mi[5]);
        const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(Mi[6], dbx), __dmul_rn(Mi[7], dy)), Mi[8]);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const double dx1 = (double)(x1 + k);
            double Wd = __dadd_rn(W0, __dmul_rn(Mi[6], dx1));
            Wd = (Wd != 0.0) ? __ddiv_rn(32.0, Wd) : 0.0;
            X[k] = coord_to_fixed(__dadd_rn(X0, __dmul_rn(Mi[0], dx1)), Wd);
            Y[k] = coord_to_fixed(__dadd_rn(Y0, __dmul_rn(Mi[3], dx1)), Wd);
            const int sx = sat_s16(X[k] >> 5), sy = sat_s16(Y[k] >> 5);
            inside = inside && sx >= 0 && sx + 1 < W && sy >= 0 && sy + 1 < H;
            sxmin = min(sxmin, sx); sxmax = max(sxmax, sx);
            symin = min(symin, sy); symax = max(symax, sy);
        }
    }
    sxmin = __reduce_min_sync(0xffffffffu, sxmin); sxmax = __reduce_max_sync(0xffffffffu, sxmax);
    symin = __reduce_min_sync(0xffffffffu, symin); symax = __reduce_max_sync(0xffffffffu, symax);
    const bool all_in = __all_sync(0xffffffffu, inside);
    if ((tid & 31) == 0) {
        atomicMin(&s_box[0], sxmin); atomicMax(&s_box[1], sxmax);
        atomicMin(&s_box[2], symin); atomicMax(&s_box[3], symax);
        if (!all_in) s_ok = 0;
    }
    __syncthreads();
    // box of source pixels [bx0, bx1] x [by0, by1] (taps reach one pixel right / down)
    const int bx0 = s_box[0], bx1 = s_box[1] + 1, by0 = s_box[2], by1 = s_box[3] + 1;
    const int nrows = by1 - by0 + 1;
    const uint32_t span = (uint32_t)(bx1 - bx0 + 1) * 3u;               // bytes of a row of the box
    const uint32_t pitch = (span + 15u + 15u + 15u) & ~15u;             // + worst-case misalignment, rounded to 16
    bool staged = s_ok && nrows >= 1 && nrows <= WS_MAX_SRC_ROWS && (size_t)nrows * pitch <= WS_SMEM_BYTES;
    if (staged) {
        // the last row's rounded-up copy must stay inside the frames buffer
        const size_t last = (size_t)f * frame_pitch + (size_t)by1 * row_pitch + (size_t)bx0 * 3;
        if (((last & ~(size_t)15) + pitch) > total_bytes) staged = false;
    }
    if (staged) {
        const uint32_t bar = smem_u32(&s_bar);
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t)nrows * pitch);
            for (int r = 0; r < nrows; r++) {
                const uint8_t *src = frame + (size_t)(by0 + r) * row_pitch + (size_t)bx0 * 3;
                const uintptr_t a = (uintptr_t)src;
                s_rowoff[r] = (uint32_t)r * pitch + (uint32_t)(a & 15);
                bulk_g2s(smem_u32(ws_tile) + (uint32_t)r * pitch, (const void *)(a & ~(uintptr_t)15), pitch, bar);
            }
        }
        __syncthreads();            // s_rowoff visible
        mbar_wait(bar, 0);
    }
    if (!active) return;

    uint32_t out[3] = {0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int sx = sat_s16(X[k] >> 5), sy = sat_s16(Y[k] >> 5);
        const int fx = X[k] & 31, fy = Y[k] & 31;
        const int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy), w10 = (32 - fx) * fy, w11 = fx * fy;
        uint32_t a_lo, a_hi, b_lo, b_hi;
        if (staged) {
            const uint32_t base = smem_u32(ws_tile) + (uint32_t)(sx - bx0) * 3u;
            load6_smem(base + s_rowoff[sy - by0], a_lo, a_hi);
            load6_smem(base + s_rowoff[sy + 1 - by0], b_lo, b_hi);
        } else {
            load_row_taps(frame, row_pitch, H, W, sy, sx, a_lo, a_hi);
            load_row_taps(frame, row_pitch, H, W, sy + 1, sx, b_lo, b_hi);
        }
        const int t00b = a_lo & 0xff, t00g = (a_lo >> 8) & 0xff, t00r = (a_lo >> 16) & 0xff;
        const int t01b = a_lo >> 24, t01g = a_hi & 0xff, t01r = (a_hi >> 8) & 0xff;
        const int t10b = b_lo & 0xff, t10g = (b_lo >> 8) & 0xff, t10r = (b_lo >> 16) & 0xff;
        const int t11b = b_lo >> 24, t11g = b_hi & 0xff, t11r = (b_hi >> 8) & 0xff;
        const uint32_t vb = (uint32_t)(t00b * w00 + t01b * w01 + t10b * w10 + t11b * w11 + 512) >> 10;
        const uint32_t vg = (uint32_t)(t00g * w00 + t01g * w01 + t10g * w10 + t11g * w11 + 512) >> 10;
        const uint32_t vr = (uint32_t)(t00r * w00 + t01r * w01 + t10r * w10 + t11r * w11 + 512) >> 10;
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

static int ckb_warp_use_staged()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("CKB_WARP_STAGED"); v = (e && e[0] == '1') ? 1 : 0; }
    return v;
}

int ckb_launch_warp(ckb_ctx *ctx, const uint8_t *d_frames, int n, int H, int W, size_t row_pitch, size_t frame_pitch,
                    const double *h_minv, int n_mtx, uint8_t *d_goban, cudaStream_t st)
{
    const int S = ctx->S;
    const int threads = 128;
    const int work = (S >> 2) * S;
    // one homography for the whole batch (a video segment): a single launch; per-frame matrices travel 32 at a time
    const int chunk = n_mtx == 1 ? 65535 : CKB_WARP_CHUNK;
    for (int f0 = 0; f0 < n; f0 += chunk) {
        const int nf = n - f0 < chunk ? n - f0 : chunk;
        WarpMats mats;
        if (n_mtx == 1)
            memcpy(mats.m[0], h_minv, 9 * sizeof(double));
        else
            memcpy(mats.m[0], h_minv + (size_t)f0 * 9, (size_t)nf * 9 * sizeof(double));
        if (ckb_warp_use_staged()) {
            static bool attr_done = false;
            if (!attr_done) {
                CKB_CUDA(ctx, cudaFuncSetAttribute(ckb_warp_kernel_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM_BYTES));
                attr_done = true;
            }
            dim3 sgrid(((S + 63) / 64) * ((S + WS_ROWS - 1) / WS_ROWS), nf);
            const size_t total = (size_t)(nf - 1) * frame_pitch + (size_t)H * row_pitch;
            ckb_warp_kernel_staged<<<sgrid, 128, WS_SMEM_BYTES, st>>>(d_frames + (size_t)f0 * frame_pitch, H, W, row_pitch,
                                                                      frame_pitch, total, mats, n_mtx != 1,
                                                                      d_goban + (size_t)f0 * S * S * 3, S);
            CKB_LAUNCH_CHECK(ctx, "ckb_warp_kernel");
            continue;
        }
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
