// C ABI of camkifu_b200 (see include/camkifu_b200.h): context management and argument validation; the kernels live in
// warp.cu (K1), kmeans.cu (K3), zones.cu (K2), cnn_*.cu (K4).
#include <new>

#include "ckb_common.cuh"

int ckb_kmeans_init_tables(ckb_ctx *ctx);
void ckb_cnn_free(ckb_ctx *ctx);

extern "C" int ckb_version(void) { return CKB_VERSION; }

extern "C" const char *ckb_last_error(const ckb_ctx *ctx) { return ctx ? ctx->err : "null context"; }

extern "C" uint64_t ckb_launch_count(const ckb_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" uint64_t ckb_rng_seed(uint32_t seed) { return seed ? (uint64_t)seed : 0xffffffffULL; }  // cv::RNG(seed)

extern "C" uint64_t ckb_rng_advance(uint64_t state, uint64_t n_draws)
{
    // cv::RNG::next(): multiply-with-carry, state = (uint32)state * 4164903690 + (state >> 32)
    for (uint64_t i = 0; i < n_draws; i++) state = (uint64_t)(uint32_t)state * 4164903690ULL + (state >> 32);
    return state;
}

extern "C" int ckb_invert_homography(const double *m, double *t)
{
    // cv::invert, 3x3 CV_64F: cofactors times the reciprocal determinant (what warpPerspective applies to `transform`).
    // This translation unit is compiled with host FMA contraction disabled (see build.py).
    if (!m || !t) return CKB_E_INVALID;
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0.) {
        for (int k = 0; k < 9; k++) t[k] = 0.;
        return CKB_E_INVALID;
    }
    det = 1. / det;
    double r[9];
    r[0] = (e * i - f * h) * det;
    r[1] = (c * h - b * i) * det;
    r[2] = (b * f - c * e) * det;
    r[3] = (f * g - d * i) * det;
    r[4] = (a * i - c * g) * det;
    r[5] = (c * d - a * f) * det;
    r[6] = (d * h - e * g) * det;
    r[7] = (b * g - a * h) * det;
    r[8] = (a * e - b * d) * det;
    for (int k = 0; k < 9; k++) t[k] = r[k];
    return CKB_OK;
}

extern "C" int ckb_create(ckb_ctx **out, int device, int gsize)
{
    if (!out) return CKB_E_INVALID;
    *out = nullptr;
    if (gsize != 9 && gsize != 13 && gsize != 19) return CKB_E_INVALID;
    ckb_ctx *ctx = new (std::nothrow) ckb_ctx();
    if (!ctx) return CKB_E_NOMEM;
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->gsize = gsize;
    ctx->S = 20 * gsize;
    *out = ctx;  // returned even on failure so that the caller can read ckb_last_error(); destroy it either way
    CKB_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    CKB_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) CKB_FAIL(ctx, CKB_E_STATE, "camkifu_b200 is built for sm_100a only; device %d is sm_%d%d", device,
                                   prop.major, prop.minor);
    ctx->num_sms = prop.multiProcessorCount;
    ckb_host_zone_rects(gsize, ctx->h_rects);
    const size_t S2 = (size_t)ctx->S * ctx->S;
    uint8_t *mask = new (std::nothrow) uint8_t[S2];
    if (!mask) CKB_FAIL(ctx, CKB_E_NOMEM, "host allocation failed");
    ckb_host_zone_mask(gsize, ctx->h_rects, mask);
    cudaError_t e = cudaMalloc(&ctx->d_rects, sizeof(int32_t) * 4 * gsize * gsize);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_mask, S2);
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_rects, ctx->h_rects, sizeof(int32_t) * 4 * gsize * gsize, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_mask, mask, S2, cudaMemcpyHostToDevice);
    delete[] mask;
    if (e != cudaSuccess) CKB_FAIL(ctx, CKB_E_CUDA, "table upload failed: %s", cudaGetErrorString(e));
    return ckb_kmeans_init_tables(ctx);
}

extern "C" int ckb_destroy(ckb_ctx *ctx)
{
    if (!ctx) return CKB_E_INVALID;
    cudaSetDevice(ctx->device);
    ckb_cnn_free(ctx);
    if (ctx->d_rects) cudaFree(ctx->d_rects);
    if (ctx->d_mask) cudaFree(ctx->d_mask);
    delete ctx;
    return CKB_OK;
}

extern "C" int ckb_warp(ckb_ctx *ctx, const uint8_t *d_frames, int n, int H, int W, size_t row_pitch,
                        size_t frame_pitch, const double *h_mtx, int n_mtx, uint8_t *d_goban, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    if (!d_frames || !d_goban || !h_mtx || n < 0 || H < 1 || W < 1) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_warp: bad argument");
    if (n_mtx != 1 && n_mtx != n) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_warp: n_mtx must be 1 or n");
    if (row_pitch < (size_t)W * 3 || (n > 1 && frame_pitch < row_pitch * (size_t)H))
        CKB_FAIL(ctx, CKB_E_INVALID, "ckb_warp: pitches smaller than the image");
    if (H > 32767 || W > 32767) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_warp: frames larger than 32767 are not supported (as OpenCV)");
    if (n == 0) return CKB_OK;
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    // invert on the host exactly as OpenCV does; a singular matrix yields the zero matrix (every tap at (0,0))
    double *minv = new (std::nothrow) double[(size_t)n_mtx * 9];
    if (!minv) CKB_FAIL(ctx, CKB_E_NOMEM, "host allocation failed");
    for (int k = 0; k < n_mtx; k++) ckb_invert_homography(h_mtx + (size_t)k * 9, minv + (size_t)k * 9);
    int rc = ckb_launch_warp(ctx, d_frames, n, H, W, row_pitch, frame_pitch, minv, n_mtx, d_goban, (cudaStream_t)stream);
    delete[] minv;
    return rc;
}

extern "C" int ckb_accumulate(ckb_ctx *ctx, const uint8_t *d_goban, int n, float *d_accu, float alpha, int first,
                              float *d_snapshots, int snap_every, int snap_phase, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    if (!d_goban || !d_accu || n < 0 || snap_phase < 0) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_accumulate: bad argument");
    if (n == 0) return CKB_OK;
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    return ckb_launch_accumulate(ctx, d_goban, n, d_accu, alpha, first, d_snapshots, snap_every, snap_phase,
                                 (cudaStream_t)stream);
}
