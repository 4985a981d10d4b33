// Constant geometry tables of the canonical goban image (host side).
//
// Reference: PosGrid.__init__ (src/camkifu/stone/stonesfinder.py:964-981) places intersection (r, c) at pixel
// (10 + 20 r, 10 + 20 c) of the 20*gsize canonical image; getrect(r, c, cursor=1.0) (:412-450) then yields the zone
// [20 r, 20 r + 20) x [20 c, 20 c + 20), except that the last row / column ends at 20*gsize - 1 (the "- 2" of :440,444
// under float->int truncation); getmask (:452-493) keeps, per zone of h x w pixels, the disk
// (x - w/2)^2 + (y - h/2)^2 <= min(h/2, w/2)^2 sampled at integer offsets from -h/2, -w/2.
#include "ckb_common.cuh"

void ckb_host_zone_rects(int gsize, int32_t *rects)
{
    const int S = 20 * gsize;
    for (int r = 0; r < gsize; r++)
        for (int c = 0; c < gsize; c++) {
            int32_t *q = rects + (r * gsize + c) * 4;
            q[0] = 20 * r;
            q[1] = 20 * c;
            q[2] = (r == gsize - 1) ? S - 1 : 20 * r + 20;
            q[3] = (c == gsize - 1) ? S - 1 : 20 * c + 20;
        }
}

void ckb_host_zone_mask(int gsize, const int32_t *rects, uint8_t *mask)
{
    const int S = 20 * gsize;
    memset(mask, 0, (size_t)S * S);
    for (int z = 0; z < gsize * gsize; z++) {
        const int32_t *q = rects + z * 4;
        const int h = q[2] - q[0], w = q[3] - q[1];
        const double a = 0.5 * h, b = 0.5 * w, rad = a < b ? a : b;
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++) {
                const double dy = i - a, dx = j - b;
                mask[(size_t)(q[0] + i) * S + q[1] + j] = dx * dx + dy * dy <= rad * rad ? 1 : 0;
            }
    }
}

extern "C" int ckb_zone_rects(int gsize, int32_t *rects_out)
{
    if (gsize < 2 || gsize > CKB_MAX_G || !rects_out) return CKB_E_INVALID;
    ckb_host_zone_rects(gsize, rects_out);
    return CKB_OK;
}

extern "C" int ckb_zone_mask(int gsize, uint8_t *mask_out)
{
    if (gsize < 2 || gsize > CKB_MAX_G || !mask_out) return CKB_E_INVALID;
    int32_t rects[CKB_MAX_ZONES * 4];
    ckb_host_zone_rects(gsize, rects);
    ckb_host_zone_mask(gsize, rects, mask_out);
    return CKB_OK;
}
