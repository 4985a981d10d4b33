// Per-zone statistics of SfMeta and its low-density delegate SfContours.
//
//   ckb_zone_means    SfContours.find_stones' zone table (src/camkifu/stone/sf_contours.py:87-102) and _norm_channels
//                     (:113-126): given the image and the mask of filled convex hulls (the contour extraction itself —
//                     Canny, findContours, convexHull, drawContours — stays on the CPU), per zone: visible = more than 40 %
//                     of the zone lies under the mask; the int16 mean B, G, R of the visible pixels (image * mask) if so,
//                     else of the masked-out pixels (image * (1 - mask)).
//   ckb_history_vote  Region.commit (src/camkifu/stone/sf_meta.py:305-340): per intersection, the colours recorded in the
//                     last `histo` detection results -> the move to submit, if any.
//
// One warp per zone, warp-shuffle reductions (REDUX); byte traffic = the region's pixels + mask once.
#include "ckb_common.cuh"

__global__ void __launch_bounds__(256) ckb_zone_means_kernel(const uint8_t *__restrict__ imgs, const uint8_t *__restrict__ masks,
                                                             int S, int gsize, int rs, int re, int cs, int ce,
                                                             const int32_t *__restrict__ rects, int16_t *__restrict__ zones)
{
    const int frame = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int zw_n = ce - cs, nz = (re - rs) * zw_n;
    const int z = blockIdx.x * (blockDim.x >> 5) + warp;
    if (z >= nz) return;
    const int r = rs + z / zw_n, c = cs + z % zw_n;
    const int32_t *q = rects + (r * gsize + c) * 4;
    const int a0 = q[0], b0 = q[1], a1 = q[2], b1 = q[3];
    const int zw = b1 - b0, area = (a1 - a0) * zw;
    const uint8_t *img = imgs + (size_t)frame * S * S * 3;
    const uint8_t *mask = masks + (size_t)frame * S * S;
    int vis = 0, sv[3] = {0, 0, 0}, sa[3] = {0, 0, 0};
    for (int t = lane; t < area; t += 32) {
        const int i = a0 + t / zw, j = b0 + t % zw;
        const size_t o = (size_t)i * S + j;
        const int m = __ldg(mask + o) != 0;
        vis += m;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int v = __ldg(img + o * 3 + k);
            sa[k] += v;
            sv[k] += m ? v : 0;
        }
    }
    vis = __reduce_add_sync(0xffffffffu, vis);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        sa[k] = __reduce_add_sync(0xffffffffu, sa[k]);
        sv[k] = __reduce_add_sync(0xffffffffu, sv[k]);
    }
    if (lane == 0) {
        // `if 0.4 * area < visible_area` in float64; then np.sum(channel) / norm in float64, truncated into the int16 slot
        const bool visible = __dmul_rn(0.4, (double)area) < (double)vis;
        int16_t *o = zones + ((size_t)frame * nz + z) * 4;
        o[0] = visible ? 1 : 0;
        const double norm = visible ? (double)vis : (double)(area - vis);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double s = visible ? (double)sv[k] : (double)(sa[k] - sv[k]);
            o[k + 1] = (int16_t)__ddiv_rn(s, norm);
        }
    }
}

__global__ void __launch_bounds__(256) ckb_history_vote_kernel(const uint8_t *__restrict__ hist, const uint8_t *__restrict__ empty,
                                                               int n_items, int histo, uint8_t *__restrict__ moves)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    int cnt[3] = {0, 0, 0};
    for (int h = 0; h < histo; h++) {
        const int v = hist[(size_t)i * histo + h];
        cnt[v < 3 ? v : 0]++;
    }
    const int e = cnt[CKB_E], b = cnt[CKB_B], w = cnt[CKB_W];
    const int distinct = (e > 0) + (b > 0) + (w > 0);
    uint8_t mv = 0;
    if (empty[i]) {
        // exactly one of B / W seen, next to E: commit if E fills less than 40 % of the history (counts[k] / cb.size < 0.4)
        if (distinct == 2 && e > 0) {
            if (__ddiv_rn((double)e, (double)histo) < 0.4) mv = b > 0 ? CKB_B : CKB_W;
        } else if (distinct == 1 && e == 0) {
            mv = b > 0 ? CKB_B : CKB_W;
        }
    }
    moves[i] = mv;
}

extern "C" int ckb_zone_means(ckb_ctx *ctx, const uint8_t *d_imgs, const uint8_t *d_masks, int n, int rs, int re, int cs, int ce,
                              int16_t *d_zones, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    if (n == 0) return CKB_OK;
    const int g = ctx->gsize;
    if (!d_imgs || !d_masks || !d_zones || n < 0) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_zone_means: bad argument");
    if (rs < 0 || cs < 0 || re > g || ce > g || rs >= re || cs >= ce)
        CKB_FAIL(ctx, CKB_E_INVALID, "ckb_zone_means: region [%d,%d)x[%d,%d) outside the %dx%d goban", rs, re, cs, ce, g, g);
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    CKB_ENTER(ctx, stream);
    const int nz = (re - rs) * (ce - cs);
    ckb_zone_means_kernel<<<dim3((nz + 7) / 8, n), 256, 0, (cudaStream_t)stream>>>(d_imgs, d_masks, ctx->S, g, rs, re, cs, ce,
                                                                                ctx->d_rects, d_zones);
    CKB_LAUNCH_CHECK(ctx, "ckb_zone_means_kernel");
    return CKB_OK;
}

extern "C" int ckb_history_vote(ckb_ctx *ctx, const uint8_t *d_history, const uint8_t *d_is_empty, int n_items, int histo,
                                uint8_t *d_moves, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    if (n_items == 0) return CKB_OK;
    if (!d_history || !d_is_empty || !d_moves || n_items < 0 || histo < 1)
        CKB_FAIL(ctx, CKB_E_INVALID, "ckb_history_vote: bad argument");
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    CKB_ENTER(ctx, stream);
    ckb_history_vote_kernel<<<(n_items + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_history, d_is_empty, n_items, histo, d_moves);
    CKB_LAUNCH_CHECK(ctx, "ckb_history_vote_kernel");
    return CKB_OK;
}
