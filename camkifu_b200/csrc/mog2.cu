// Background model of every finder: cv2.createBackgroundSubtractorMOG2(detectShadows=False).apply(goban_img, lr)
// (StonesFinder.__init__ / _learn_bg, src/camkifu/stone/stonesfinder.py:113-115,171-176) and the per-zone foreground
// counts its consumers reduce the mask to (SfNeural.is_agitated, sf_neural.py:178-180).
//
// Zivkovic's adaptive Gaussian mixture as OpenCV implements it (video/bgfg_gaussmix2.cpp, MOG2Invoker): up to 5 modes per
// pixel kept sorted by weight; the defaults of createBackgroundSubtractorMOG2 (history 500, varThreshold 16,
// varThresholdGen 9, backgroundRatio 0.9, varInit 15, varMin 4, varMax 75, CT 0.05). Bit-exact model: float32 without
// FMA contraction, same operation order, same dynamic loop bounds (oracle: cko_mog2_apply, pinned against cv2).
//
// One thread owns one pixel: its 25-float state is read once, n frames are applied in order in registers (the model
// is a per-pixel recurrence: sequential in time, parallel across the 144 400 pixels), and written back once — HBM
// traffic per launch = 2 x state (29 MB) + n x (image 433 KB + mask 144 KB). State is planar ([25][pixels] floats
// followed by [pixels] mode counts) so that every load and store is coalesced.
#include "ckb_common.cuh"

#define MOG2_MAX_FRAMES 64

struct Mog2Rates {
    float alpha[MOG2_MAX_FRAMES];   // alphaT of each frame
    float prune[MOG2_MAX_FRAMES];   // float(-learningRate * CT)
};

struct Mog2Px {
    float w[5], var[5], m0[5], m1[5], m2[5];
    int nmodes;
};

__device__ __forceinline__ void mog2_swap(Mog2Px &s, const int i, const int j)
{
    float t;
    t = s.w[i]; s.w[i] = s.w[j]; s.w[j] = t;
    t = s.var[i]; s.var[i] = s.var[j]; s.var[j] = t;
    t = s.m0[i]; s.m0[i] = s.m0[j]; s.m0[j] = t;
    t = s.m1[i]; s.m1[i] = s.m1[j]; s.m1[j] = t;
    t = s.m2[i]; s.m2[i] = s.m2[j]; s.m2[j] = t;
}

// one frame of one pixel; returns the mask value. Every array index is a compile-time constant after unrolling.
__device__ __forceinline__ uint8_t mog2_update(Mog2Px &s, const float d0, const float d1, const float d2, const float alphaT,
                                               const float prune)
{
    const float Tb = 16.f, Tg = 9.f, TB = 0.9f, varInit = 15.f, varMin = 4.f, varMax = 75.f;
    const float alpha1 = __fsub_rn(1.f, alphaT);
    bool background = false, fits = false;
    float total = 0.f;
    int nmodes = s.nmodes;
#pragma unroll
    for (int mode = 0; mode < 5; mode++) {
        if (mode < nmodes) {                                  // the bound shrinks as modes are pruned, as in OpenCV
            float weight = __fadd_rn(__fmul_rn(alpha1, s.w[mode]), prune);
            int pos = mode;
            if (!fits) {
                const float v = s.var[mode];
                const float e0 = __fsub_rn(s.m0[mode], d0), e1 = __fsub_rn(s.m1[mode], d1), e2 = __fsub_rn(s.m2[mode], d2);
                const float dist2 = __fadd_rn(__fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1)), __fmul_rn(e2, e2));
                if (total < TB && dist2 < __fmul_rn(Tb, v)) background = true;
                if (dist2 < __fmul_rn(Tg, v)) {
                    fits = true;
                    weight = __fadd_rn(weight, alphaT);
                    const float k = __fdiv_rn(alphaT, weight);
                    s.m0[mode] = __fsub_rn(s.m0[mode], __fmul_rn(k, e0));
                    s.m1[mode] = __fsub_rn(s.m1[mode], __fmul_rn(k, e1));
                    s.m2[mode] = __fsub_rn(s.m2[mode], __fmul_rn(k, e2));
                    float varnew = __fadd_rn(v, __fmul_rn(k, __fsub_rn(dist2, v)));
                    varnew = varnew > varMin ? varnew : varMin;
                    varnew = varnew < varMax ? varnew : varMax;
                    s.var[mode] = varnew;
                    bool sorting = true;                      // bubble the matched mode up past lighter ones
#pragma unroll
                    for (int i = mode; i > 0; i--) {
                        if (sorting) {
                            if (weight < s.w[i - 1]) sorting = false;
                            else { mog2_swap(s, i, i - 1); pos = i - 1; }
                        }
                    }
                }
            }
            if (weight < -prune) {
                weight = 0.f;
                nmodes--;
            }
#pragma unroll
            for (int j = 0; j <= mode; j++)
                if (j == pos) s.w[j] = weight;
            total = __fadd_rn(total, weight);
        }
    }
    float inv = 0.f;
    if (fabsf(total) > 1.1920929e-7f) inv = __fdiv_rn(1.f, total);
#pragma unroll
    for (int mode = 0; mode < 5; mode++)
        if (mode < nmodes) s.w[mode] = __fmul_rn(s.w[mode], inv);
    if (!fits && alphaT > 0.f) {
        int mode;
        if (nmodes == 5) mode = 4; else mode = nmodes++;
#pragma unroll
        for (int j = 0; j < 5; j++) {
            if (j == mode) {
                s.w[j] = nmodes == 1 ? 1.f : alphaT;
                s.m0[j] = d0; s.m1[j] = d1; s.m2[j] = d2;
                s.var[j] = varInit;
            } else if (nmodes != 1 && j < nmodes - 1) {
                s.w[j] = __fmul_rn(s.w[j], alpha1);
            }
        }
        bool sorting = true;
#pragma unroll
        for (int i = 4; i > 0; i--) {
            if (i <= nmodes - 1 && sorting) {
                if (alphaT < s.w[i - 1]) sorting = false;
                else mog2_swap(s, i, i - 1);
            }
        }
    }
    s.nmodes = nmodes;
    return background ? 0 : 255;
}

#ifndef MOG2_MINB
#define MOG2_MINB 8      // resident CTAs per SM the register allocation is sized for (measured: see DESIGN.md)
#endif
__global__ void __launch_bounds__(128, MOG2_MINB) ckb_mog2_kernel(const uint8_t *__restrict__ img, int n, int npix, float *__restrict__ state,
                                                       const __grid_constant__ Mog2Rates rates, uint8_t *__restrict__ mask)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    uint8_t *nm = (uint8_t *)(state + (size_t)25 * npix);
    Mog2Px s;
    s.nmodes = nm[p];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        s.w[k] = state[(size_t)k * npix + p];
        s.var[k] = state[(size_t)(5 + k) * npix + p];
        s.m0[k] = state[(size_t)(10 + 3 * k) * npix + p];
        s.m1[k] = state[(size_t)(11 + 3 * k) * npix + p];
        s.m2[k] = state[(size_t)(12 + 3 * k) * npix + p];
    }
    for (int f = 0; f < n; f++) {
        const uint8_t *px = img + ((size_t)f * npix + p) * 3;
        const float d0 = (float)__ldg(px), d1 = (float)__ldg(px + 1), d2 = (float)__ldg(px + 2);
        mask[(size_t)f * npix + p] = mog2_update(s, d0, d1, d2, rates.alpha[f], rates.prune[f]);
    }
    nm[p] = (uint8_t)s.nmodes;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        state[(size_t)k * npix + p] = s.w[k];
        state[(size_t)(5 + k) * npix + p] = s.var[k];
        state[(size_t)(10 + 3 * k) * npix + p] = s.m0[k];
        state[(size_t)(11 + 3 * k) * npix + p] = s.m1[k];
        state[(size_t)(12 + 3 * k) * npix + p] = s.m2[k];
    }
}

// per zone: number of foreground pixels of the zone rectangle = np.sum(fg[a0:a1, b0:b1]) / 255 (sf_neural.py:178-180).
// One warp per (frame, zone), warp-shuffle reduction.
__global__ void __launch_bounds__(128) ckb_zone_fg_kernel(const uint8_t *__restrict__ mask, int n, int S, int nz,
                                                          const int32_t *__restrict__ rects, int32_t *__restrict__ counts)
{
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (gw >= n * nz) return;
    const int f = gw / nz, z = gw - f * nz;
    const int x0 = rects[z * 4], y0 = rects[z * 4 + 1], x1 = rects[z * 4 + 2], y1 = rects[z * 4 + 3];   // rows x0..x1, cols y0..y1
    const int w = y1 - y0, area = (x1 - x0) * w;
    const uint8_t *m = mask + (size_t)f * S * S;
    int c = 0;
    for (int i = lane; i < area; i += 32) {
        const int r = i / w, q = i - r * w;
        c += m[(size_t)(x0 + r) * S + y0 + q] != 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) counts[gw] = c;
}

extern "C" size_t ckb_mog2_state_bytes(const ckb_ctx *ctx)
{
    if (!ctx) return 0;
    const size_t npix = (size_t)ctx->S * ctx->S;
    return (npix * 25 * sizeof(float) + npix + 255) / 256 * 256;
}

extern "C" int ckb_mog2_reset(ckb_ctx *ctx, void *d_state, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    if (!d_state) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_mog2_reset: bad argument");
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    CKB_CUDA(ctx, cudaMemsetAsync(d_state, 0, ckb_mog2_state_bytes(ctx), (cudaStream_t)stream));
    return CKB_OK;
}

extern "C" int ckb_mog2_apply(ckb_ctx *ctx, const uint8_t *d_goban, int n, void *d_state, long long frames_before,
                              const double *h_learning_rates, uint8_t *d_fgmask, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    if (n == 0) return CKB_OK;
    if (!d_goban || !d_state || !d_fgmask || !h_learning_rates || n < 0 || frames_before < 0)
        CKB_FAIL(ctx, CKB_E_INVALID, "ckb_mog2_apply: bad argument");
    if (((uintptr_t)d_state & 3) != 0) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_mog2_apply: state must be 4-byte aligned");
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    CKB_ENTER(ctx, stream);
    const int npix = ctx->S * ctx->S;
    const int history = 500;
    const float CT = 0.05f;
    for (int f0 = 0; f0 < n; f0 += MOG2_MAX_FRAMES) {
        const int m = n - f0 < MOG2_MAX_FRAMES ? n - f0 : MOG2_MAX_FRAMES;
        Mog2Rates r;
        for (int i = 0; i < m; i++) {
            const long long frame_no = frames_before + f0 + i + 1;   // OpenCV's nframes after its increment
            const double lr = h_learning_rates[f0 + i];
            if (lr >= 1.) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_mog2_apply: learning rates >= 1 re-initialise the model in OpenCV; call ckb_mog2_reset");
            const double lrd = (lr >= 0 && frame_no > 1) ? lr : 1. / (double)(2 * frame_no < history ? 2 * frame_no : history);
            r.alpha[i] = (float)lrd;
            r.prune[i] = (float)(-lrd * CT);
        }
        ckb_mog2_kernel<<<(npix + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_goban + (size_t)f0 * npix * 3, m, npix, (float *)d_state, r,
                                                                             d_fgmask + (size_t)f0 * npix);
        CKB_LAUNCH_CHECK(ctx, "ckb_mog2_kernel");
    }
    return CKB_OK;
}

extern "C" int ckb_zone_fg_counts(ckb_ctx *ctx, const uint8_t *d_fgmask, int n, int32_t *d_counts, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    if (n == 0) return CKB_OK;
    if (!d_fgmask || !d_counts || n < 0) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_zone_fg_counts: bad argument");
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    CKB_ENTER(ctx, stream);
    const int nz = ctx->gsize * ctx->gsize;
    const long long warps = (long long)n * nz;
    ckb_zone_fg_kernel<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_fgmask, n, ctx->S, nz, ctx->d_rects, d_counts);
    CKB_LAUNCH_CHECK(ctx, "ckb_zone_fg_kernel");
    return CKB_OK;
}
