// C ABI of camkifu_b200 (see include/camkifu_b200.h): context management and argument validation; the kernels live in
// warp.cu (K1), kmeans.cu (K3), zones.cu (K2), cnn_*.cu (K4).
#include <math.h>

#include <new>

#include "ckb_common.cuh"

int ckb_kmeans_init_tables(ckb_ctx *ctx);
void ckb_cnn_free(ckb_ctx *ctx);
void ckb_jpeg_free(ckb_ctx *ctx);

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

// ---- per-kernel timing -------------------------------------------------------------------------------------------
void ckb_prof_mark(ckb_ctx *ctx, const char *name)
{
    if (ctx->prof_n >= ctx->prof_cap) return;  // full: later launches go untimed
    ckb_prof_entry &e = ctx->prof[ctx->prof_n];
    if (!e.ev && cudaEventCreate(&e.ev) != cudaSuccess) return;
    e.name = name;
    if (cudaEventRecord(e.ev, ctx->cur_stream) == cudaSuccess) ctx->prof_n++;
}

extern "C" int ckb_profile_begin(ckb_ctx *ctx, int capacity)
{
    if (!ctx || capacity < 2) return CKB_E_INVALID;
    if (capacity > ctx->prof_cap) {
        ckb_prof_entry *p = new (std::nothrow) ckb_prof_entry[capacity];
        if (!p) CKB_FAIL(ctx, CKB_E_NOMEM, "host allocation failed");
        memset(p, 0, sizeof(ckb_prof_entry) * capacity);
        for (int i = 0; i < ctx->prof_cap; i++) p[i] = ctx->prof[i];
        delete[] ctx->prof;
        ctx->prof = p;
        ctx->prof_cap = capacity;
    }
    ctx->prof_n = 0;
    ctx->prof_on = 1;
    return CKB_OK;
}

extern "C" int ckb_profile_end(ckb_ctx *ctx, int max_entries, char *names, float *ms, int *n_out)
{
    if (!ctx || !names || !ms || !n_out) return CKB_E_INVALID;
    ctx->prof_on = 0;
    *n_out = 0;
    if (ctx->prof_n == 0) return CKB_OK;
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    CKB_CUDA(ctx, cudaEventSynchronize(ctx->prof[ctx->prof_n - 1].ev));
    int k = 0;
    for (int i = 1; i < ctx->prof_n && k < max_entries; i++) {
        if (!ctx->prof[i].name) continue;
        float t = 0.f;
        CKB_CUDA(ctx, cudaEventElapsedTime(&t, ctx->prof[i - 1].ev, ctx->prof[i].ev));
        strncpy(names + (size_t)k * 32, ctx->prof[i].name, 31);
        names[(size_t)k * 32 + 31] = 0;
        ms[k++] = t;
    }
    *n_out = k;
    return CKB_OK;
}

// ---- host -> device staging of the part of each frame the warp can touch --------------------------------------------
extern "C" int ckb_frame_roi(const double *m9, int H, int W, int S, int *roi4)
{
    // The canonical square [0, S-1]^2 maps (through the inverse homography) inside the convex hull of its four corner
    // images; bilinear taps reach one pixel further right / down. roi = {y0, y1, x0, x1}, half-open, clipped.
    if (!m9 || !roi4 || H < 1 || W < 1 || S < 1) return CKB_E_INVALID;
    double inv[9];
    if (ckb_invert_homography(m9, inv) != CKB_OK) { roi4[0] = 0; roi4[1] = H; roi4[2] = 0; roi4[3] = W; return CKB_OK; }
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    bool ok = true;
    double wsign = 0;
    for (int c = 0; c < 4; c++) {
        const double x = (c & 1) ? S - 1 : 0, y = (c & 2) ? S - 1 : 0;
        const double w = inv[6] * x + inv[7] * y + inv[8];
        if (w == 0 || (wsign != 0 && (w > 0) != (wsign > 0))) ok = false;   // horizon crosses the board: no bound
        wsign = w;
        const double sx = (inv[0] * x + inv[1] * y + inv[2]) / w, sy = (inv[3] * x + inv[4] * y + inv[5]) / w;
        xmin = sx < xmin ? sx : xmin; xmax = sx > xmax ? sx : xmax;
        ymin = sy < ymin ? sy : ymin; ymax = sy > ymax ? sy : ymax;
    }
    if (!ok || !(xmin == xmin) || !(ymin == ymin)) { roi4[0] = 0; roi4[1] = H; roi4[2] = 0; roi4[3] = W; return CKB_OK; }
    auto clampi = [](double v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : (int)v); };
    roi4[0] = clampi(floor(ymin) - 1, 0, H);
    roi4[1] = clampi(ceil(ymax) + 3, 0, H);
    roi4[2] = clampi(floor(xmin) - 1, 0, W);
    roi4[3] = clampi(ceil(xmax) + 3, 0, W);
    if (roi4[1] <= roi4[0] || roi4[3] <= roi4[2]) { roi4[0] = roi4[1] = roi4[2] = roi4[3] = 0; }
    return CKB_OK;
}

extern "C" int ckb_upload_frames(ckb_ctx *ctx, const uint8_t *h_frames, int n, int H, int W, size_t h_row_pitch,
                                 size_t h_frame_pitch, const int *roi4, uint8_t *d_frames, size_t d_row_pitch,
                                 size_t d_frame_pitch, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    if (n == 0) return CKB_OK;   // an empty batch is a no-op, whatever the pointers
    if (!h_frames || !d_frames || n < 0 || H < 1 || W < 1) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_upload_frames: bad argument");
    int y0 = 0, y1 = H, x0 = 0, x1 = W;
    if (roi4) { y0 = roi4[0]; y1 = roi4[1]; x0 = roi4[2]; x1 = roi4[3]; }
    if (y0 < 0 || x0 < 0 || y1 > H || x1 > W) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_upload_frames: roi outside the frame");
    // same pitch rules as ckb_warp, for both sides: rows hold W pixels, frames hold H rows
    if (h_row_pitch < (size_t)W * 3 || d_row_pitch < (size_t)W * 3)
        CKB_FAIL(ctx, CKB_E_INVALID, "ckb_upload_frames: row pitch smaller than a row (%d pixels)", W);
    if (n > 1 && (h_frame_pitch < h_row_pitch * (size_t)H || d_frame_pitch < d_row_pitch * (size_t)H))
        CKB_FAIL(ctx, CKB_E_INVALID, "ckb_upload_frames: frame pitch smaller than a frame (%d rows)", H);
    if (y1 <= y0 || x1 <= x0 || n == 0) return CKB_OK;
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t wbytes = (size_t)(x1 - x0) * 3;
    if (n > 1 && h_frame_pitch % h_row_pitch == 0 && d_frame_pitch % d_row_pitch == 0) {
        // one pitched 3-D copy for the whole sub-batch: width = ROI row bytes, height = ROI rows, depth = frames
        cudaMemcpy3DParms p;
        memset(&p, 0, sizeof p);
        p.srcPtr = make_cudaPitchedPtr((void *)h_frames, h_row_pitch, (size_t)W * 3, h_frame_pitch / h_row_pitch);
        p.dstPtr = make_cudaPitchedPtr((void *)d_frames, d_row_pitch, (size_t)W * 3, d_frame_pitch / d_row_pitch);
        p.srcPos = make_cudaPos((size_t)x0 * 3, (size_t)y0, 0);
        p.dstPos = make_cudaPos((size_t)x0 * 3, (size_t)y0, 0);
        p.extent = make_cudaExtent(wbytes, (size_t)(y1 - y0), (size_t)n);
        p.kind = cudaMemcpyHostToDevice;
        CKB_CUDA(ctx, cudaMemcpy3DAsync(&p, (cudaStream_t)stream));
        return CKB_OK;
    }
    for (int i = 0; i < n; i++) {
        const uint8_t *src = h_frames + (size_t)i * h_frame_pitch + (size_t)y0 * h_row_pitch + (size_t)x0 * 3;
        uint8_t *dst = d_frames + (size_t)i * d_frame_pitch + (size_t)y0 * d_row_pitch + (size_t)x0 * 3;
        if (wbytes == h_row_pitch && wbytes == d_row_pitch)
            CKB_CUDA(ctx, cudaMemcpyAsync(dst, src, wbytes * (size_t)(y1 - y0), cudaMemcpyHostToDevice, (cudaStream_t)stream));
        else
            CKB_CUDA(ctx, cudaMemcpy2DAsync(dst, d_row_pitch, src, h_row_pitch, wbytes, (size_t)(y1 - y0),
                                            cudaMemcpyHostToDevice, (cudaStream_t)stream));
    }
    return CKB_OK;
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
    ckb_jpeg_free(ctx);
    if (ctx->d_rects) cudaFree(ctx->d_rects);
    if (ctx->d_mask) cudaFree(ctx->d_mask);
    for (int i = 0; i < ctx->prof_cap; i++)
        if (ctx->prof[i].ev) cudaEventDestroy(ctx->prof[i].ev);
    delete[] ctx->prof;
    delete ctx;
    return CKB_OK;
}

extern "C" int ckb_warp(ckb_ctx *ctx, const uint8_t *d_frames, int n, int H, int W, size_t row_pitch,
                        size_t frame_pitch, const double *h_mtx, int n_mtx, uint8_t *d_goban, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    if (n == 0) return CKB_OK;   // an empty batch is a no-op, whatever the pointers
    if (!d_frames || !d_goban || !h_mtx || n < 0 || H < 1 || W < 1) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_warp: bad argument");
    if (n_mtx != 1 && n_mtx != n) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_warp: n_mtx must be 1 or n");
    if (row_pitch < (size_t)W * 3 || (n > 1 && frame_pitch < row_pitch * (size_t)H))
        CKB_FAIL(ctx, CKB_E_INVALID, "ckb_warp: pitches smaller than the image");
    if (H > 32767 || W > 32767) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_warp: frames larger than 32767 are not supported (as OpenCV)");
    if (n == 0) return CKB_OK;
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    CKB_ENTER(ctx, stream);
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
    if (n == 0) return CKB_OK;   // an empty batch is a no-op, whatever the pointers
    if (!d_goban || !d_accu || n < 0 || snap_phase < 0) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_accumulate: bad argument");
    if (n == 0) return CKB_OK;
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    CKB_ENTER(ctx, stream);
    return ckb_launch_accumulate(ctx, d_goban, n, d_accu, alpha, first, d_snapshots, snap_every, snap_phase,
                                 (cudaStream_t)stream);
}
