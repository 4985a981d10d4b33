// K3 for uint8 images — one THREAD-BLOCK CLUSTER per frame, region pixels resident in distributed shared memory, the
// three k-means attempts advanced in lock step.
//
// Replaces cv2.kmeans(pixels, 3, None, (TERM_CRITERIA_EPS, 15, 3), 3, KMEANS_PP_CENTERS)   (sf_clustering.py:103-104)
// for the canonical uint8 image, bit for bit (same contract as ckb_kmeans_attempt in kmeans.cu, which stays the path for
// float32 images and the fallback for the rare attempts this kernel declines).
//
// Why a second kernel. One CTA per (frame, attempt) left 3 n work units for 148 SMs (64 frames: 1.3 waves) and re-read the
// region from L2 in every one of its ~8 passes. Here a cluster of C CTAs (C = 8 for a full board) owns one frame: CTA r
// keeps pixels [r L, (r+1) L) of the region in its shared memory (packed uchar4, loaded once, straight from the image: no
// pack kernel, no scratch), every pass reads shared memory only, and the CTAs exchange their partial sums through
// distributed shared memory (remote st.shared::cluster + barrier.cluster). The three attempts of cv2.kmeans share the
// pixels and every pass / exchange / barrier: each chunk of pixels is loaded once per pass and classified against the
// centres of the three attempts, so the barriers, whose latency (not the arithmetic) bounds a cluster, are paid once per
// frame instead of once per attempt. n C CTAs of 512 threads, two per SM.
//
// Arithmetic. For uint8 pixels every quantity of k-means++ is an integer: distances by dp4a, sums exact in any order.
// Lloyd iterations: the label of a pixel is the first minimum of three float32 distances (8 roundings each). They are
// only evaluated where they can matter: a fixed-point filter F_ab(x) ~ 128 (d_a(x) - d_b(x)) / 2 (weights 128 (c_b - c_a)
// rounded to integers and split into two bytes: two dp4a + one shift-add per pair of centres) is within 3.06 of the
// float32 value, so |F_ab| > 3.5 * 128 decides the comparison; a warp in which some pixel is left undecided recomputes
// that group of pixels with the exact float32 chain. Centre sums are OpenCV's sequential float32 sums in pixel order:
// exact integers up to 2^24 (taken as packed 16-bit warp reductions per 128-pixel chunk, third cluster = chunk total
// minus the other two), the chunk in which a sum crosses 2^24 is walked serially by one warp, and in [2^24, 2^25) float32
// addition of an integer x to the even integer 2u is u += (x >> 1) + (x odd ? (u + (x >> 1)) & 1 : 0): a two-state
// automaton on the parity of u. The chunks after the crossing are shared out over all warps of all CTAs of the cluster
// (pixels read through distributed shared memory); each lane runs its 4 pixels for both entry parities, the lanes' tables
// are composed in order with shuffles, and so on up to the cluster, so the hand-over is a table lookup. An attempt with
// an empty cluster or a sum within reach of 2^25 is handed to ckb_kmeans_attempt (KmAttempt.iters = KM_ITERS_FALLBACK).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "kmeans_common.cuh"

namespace cg = cooperative_groups;

#define KC_CH 128                 // pixels per chunk: one warp iteration, 4 consecutive pixels per lane
#define KC_MAXC 8                 // portable cluster size limit
#define KC_T 448                  // filter threshold: 3.5 in units of 1/128 of (d_a - d_b) / 2
#define KC_2_24 (1 << 24)
#ifndef KC_EXP
#define KC_EXP 0
#endif
#define KC_GUARD 150000           // > number of pixels: bound on |float32 partial sum - exact partial sum| below 2^25

struct KcFilter {
    uint32_t lo[3], hi[3];        // pairs (0,1), (0,2), (1,2): weights 128 (c_b - c_a), low byte (unsigned) / high byte (signed)
    int k[3];                     // 64 (|c_a|^2 - |c_b|^2)
};

struct __align__(16) KcShared {
    long long xt[3][KC_MAXC][3][3];   // k-means++ [pass A, B, C][rank][attempt][candidate]: totals of the distance sums
    int cand_idx[2][3][3];            // sampled candidates [round][attempt][trial], written by the owning CTA into every CTA
    int xl[2][KC_MAXC][3][12];        // Lloyd [iteration parity][rank][attempt]: slice sums (9 chains) and counts (3)
    int2 tl[2][KC_MAXC][3][9];        // Lloyd tail: per-rank automaton tables (add for entry parity 0 / 1)
    int walk_u[2][3][9];              // half of the float32 sum after the crossing chunk (written by the rank that walks it)
    double xcomp[KC_MAXC][3];         // compactness partials (rank 0's copy is the one that is read)
    double u[3][6];                   // per attempt: the six uniform draws of k-means++
    KcFilter filt[3];
    float cen[3][9], oldc[3][9];
    uint32_t cenw[3][3];              // k-means++ centres as packed pixels
    uint32_t candw[3][3];
    int T[3][12], P[3][12];           // totals over the cluster / prefix before this rank
    int tailmask[3];
    int state[3];                     // 0 = iterating, 1 = converged, 2 = handed to the one-CTA kernel
    int iters[3];
    int best[3];
    int rc[3][9], xc[3][9], xstart[3][9];   // per tail chain: rank / local chunk of the 2^24 crossing, exact sum before it
    int2 wt[3][9][16];                // per-warp automaton tables
    float walkbuf[16][KC_CH];         // per warp: the crossing chunk's masked values in pixel order
};

__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c)
{
    uint32_t d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"((uint32_t)c));
    return (int)d;
}

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---- distributed shared memory: explicit shared::cluster accesses (a generic pointer from map_shared_rank compiles to
// generic LD / ST, a slower path into another CTA's shared memory)
__device__ __forceinline__ uint32_t dsmem_addr(const void *p, int rank)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void dsmem_st32(const void *p, int rank, uint32_t v)
{
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(dsmem_addr(p, rank)), "r"(v) : "memory");
}
__device__ __forceinline__ void dsmem_st64(const void *p, int rank, unsigned long long v)
{
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(dsmem_addr(p, rank)), "l"(v) : "memory");
}
__device__ __forceinline__ uint4 dsmem_ld128(const void *p, int rank)
{
    uint4 v;
    asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(dsmem_addr(p, rank))
                 : "memory");
    return v;
}
__device__ __forceinline__ uint2 dsmem_ld64(const void *p, int rank)
{
    uint2 v;
    asm volatile("ld.shared::cluster.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(dsmem_addr(p, rank)) : "memory");
    return v;
}

__device__ __forceinline__ float3 unpack_px(uint32_t p)
{
    return make_float3((float)(p & 0xffu), (float)((p >> 8) & 0xffu), (float)((p >> 16) & 0xffu));
}

__device__ __forceinline__ uint32_t region_px(const uint8_t *img, int S, const Region &rg, int i)
{
    const int row = i / rg.w, col = i - row * rg.w;
    const uint8_t *s = img + ((size_t)(rg.x0 + row) * S + rg.y0 + col) * 3;
    return (uint32_t)__ldg(s) | ((uint32_t)__ldg(s + 1) << 8) | ((uint32_t)__ldg(s + 2) << 16);
}

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ long long warp_incl_scan_ll(long long v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// One 128-pixel chunk of a Lloyd pass for one attempt: labels of this lane's 4 pixels against the centres `oc` (filter, or
// the exact float32 chain for the whole warp if any pixel is undecided), packed sums of clusters 0 and 1 (x | y << 16,
// z | count << 16) reduced over the warp, and the label written back into the pixel word's spare byte (2 bits per
// attempt: the tail and the compactness pass read it back). FULL: every pixel of the chunk is inside the region (all
// chunks but the region's last), so no validity masks. Classification and accumulation are one loop, so that no
// predicate outlives its pixel.
template <bool FULL>
__device__ __forceinline__ uint4 lloyd_chunk(uint4 *slot, unsigned valid, const KcFilter &F, const float *oc, int lsh)
{
    const uint4 v = *slot;
    const uint32_t p[4] = {v.x, v.y, v.z, v.w};
    const uint32_t keep = ~(3u << lsh), l1 = 1u << lsh, l2 = 2u << lsh;
    uint32_t a0 = 0u, b0 = 0u, a1 = 0u, b1 = 0u, pl[4];
    bool unsure = false;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int f01 = dp4a_uu(p[q], F.lo[0], F.k[0]) + dp4a_us(p[q], F.hi[0], 0) * 256;
        const int f02 = dp4a_uu(p[q], F.lo[1], F.k[1]) + dp4a_us(p[q], F.hi[1], 0) * 256;
        const int f12 = dp4a_uu(p[q], F.lo[2], F.k[2]) + dp4a_us(p[q], F.hi[2], 0) * 256;
        const bool a = f01 < -KC_T && f02 < -KC_T;
        const bool b = f01 > KC_T && f12 < -KC_T;
        const bool c = f02 > KC_T && f12 > KC_T;
        const bool vq = FULL || ((valid >> q) & 1u);
        const uint32_t pa = __byte_perm(p[q], 0u, 0x4140);       // x | y << 16
        const uint32_t pb = __byte_perm(p[q], 0x100u, 0x4542);   // z | 1 << 16
        if (a && vq) { a0 += pa; b0 += pb; }
        if (b && vq) { a1 += pa; b1 += pb; }
        pl[q] = (p[q] & keep) | (a ? 0u : (b ? l1 : l2));
        unsure |= vq && !(a || b || c);
    }
    if (__any_sync(0xffffffffu, unsure)) {
        a0 = b0 = a1 = b1 = 0u;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int lab = argmin3(unpack_px(p[q]), oc);
            const bool vq = FULL || ((valid >> q) & 1u);
            const uint32_t pa = __byte_perm(p[q], 0u, 0x4140);
            const uint32_t pb = __byte_perm(p[q], 0x100u, 0x4542);
            if (vq && lab == 0) { a0 += pa; b0 += pb; }
            if (vq && lab == 1) { a1 += pa; b1 += pb; }
            pl[q] = (p[q] & keep) | ((uint32_t)lab << lsh);
        }
    }
    *slot = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    a0 = __reduce_add_sync(0xffffffffu, a0);
    b0 = __reduce_add_sync(0xffffffffu, b0);
    a1 = __reduce_add_sync(0xffffffffu, a1);
    b1 = __reduce_add_sync(0xffffffffu, b1);
    return make_uint4(a0, b0, a1, b1);
}

// chunk sum of chain c = 3 k + j (cluster k, channel j) from the packed per-chunk reductions
__device__ __forceinline__ int chain_chunk_sum(const uint4 &cs, const uint2 &ct, int c)
{
    const int k = c / 3, j = c - 3 * k;
    const uint32_t a0 = j == 2 ? cs.y : cs.x, a1 = j == 2 ? cs.w : cs.z, at = j == 2 ? ct.y : ct.x;
    const int s0 = j == 1 ? (int)(a0 >> 16) : (int)(a0 & 0xffffu);
    const int s1 = j == 1 ? (int)(a1 >> 16) : (int)(a1 & 0xffffu);
    const int st = j == 1 ? (int)(at >> 16) : (int)(at & 0xffffu);
    return k == 0 ? s0 : (k == 1 ? s1 : st - s0 - s1);
}

// k-means++ pass of one attempt over this CTA's chunks: per chunk, for each of ncand candidates, the sum over the pixels of
// min(d(x, cand), d(x, nearest chosen centre)) -> seg_a[(1 + t) * cpc + chunk]   (seg_a = this attempt's four arrays)
template <int NW>
__device__ __forceinline__ void pp_pass_attempt(const KcShared &sh, int a, const uint32_t *pix, const int *cxx, int *seg_a,
                                                int cpc, int nch, int ch_lo, int N, int ncen, int ncand, int warp, int lane)
{
    uint32_t cw[3], bw[2] = {0u, 0u};
    int cc[3], bc[2] = {0, 0};
#pragma unroll
    for (int t = 0; t < 3; t++) {
        cw[t] = sh.candw[a][t < ncand ? t : 0];
        cc[t] = dp4a_uu(cw[t], cw[t], 0);
    }
#pragma unroll
    for (int b = 0; b < 2; b++) {
        if (b < ncen) {
            bw[b] = sh.cenw[a][b];
            bc[b] = dp4a_uu(bw[b], bw[b], 0);
        }
    }
    for (int lc = warp; lc < nch; lc += NW) {
        const uint4 v = ((const uint4 *)pix)[lc * 32 + lane];
        const uint32_t p[4] = {v.x, v.y, v.z, v.w};
        // pixels of the region this lane holds in the chunk: 4, except in the region's last chunk
        const int rem = (ch_lo + lc + 1) * KC_CH <= N ? 4 : N - ((ch_lo + lc) * KC_CH + lane * 4);
        int acc[3] = {0, 0, 0};
        // distances relative to |x|^2, which is added per chunk (cxx): d' = |c|^2 - 2 x.c
        if (rem >= 4) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int base = 0x7fffffff;
                if (ncen > 0) base = bc[0] - 2 * dp4a_uu(p[q], bw[0], 0);
                if (ncen > 1) base = min(base, bc[1] - 2 * dp4a_uu(p[q], bw[1], 0));
#pragma unroll
                for (int t = 0; t < 3; t++)
                    if (t < ncand) acc[t] += min(cc[t] - 2 * dp4a_uu(p[q], cw[t], 0), base);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int base = 0x7fffffff;
                if (ncen > 0) base = bc[0] - 2 * dp4a_uu(p[q], bw[0], 0);
                if (ncen > 1) base = min(base, bc[1] - 2 * dp4a_uu(p[q], bw[1], 0));
#pragma unroll
                for (int t = 0; t < 3; t++)
                    if (t < ncand && q < rem) acc[t] += min(cc[t] - 2 * dp4a_uu(p[q], cw[t], 0), base);
            }
        }
#pragma unroll
        for (int t = 0; t < 3; t++) {
            if (t < ncand) {
                const int s = __reduce_add_sync(0xffffffffu, acc[t]);
                if (lane == 0) seg_a[(1 + t) * cpc + lc] = s + cxx[lc];
            }
        }
    }
}

// the pass for the three attempts, then the per-CTA totals -> every CTA's xt[phase][rank][a][t]; one cluster barrier
template <int NW>
__device__ __forceinline__ void pp_pass_cluster(cg::cluster_group &cluster, KcShared &sh, const uint32_t *pix, const int *cxx,
                                                int *seg, int cpc, int nch, int ch_lo, int N, int ncen, int ncand, int phase,
                                                int rank, int C, int warp, int swarp, int lane)
{
#pragma unroll 1
    for (int a = 0; a < 3; a++)
        pp_pass_attempt<NW>(sh, a, pix, cxx, seg + a * 4 * cpc, cpc, nch, ch_lo, N, ncen, ncand, warp, lane);
    __syncthreads();
    if (swarp < 3 * ncand) {
        const int a = swarp / ncand, t = swarp - a * ncand;
        long long tot = 0;
        for (int lc = lane; lc < nch; lc += 32) tot += seg[(a * 4 + 1 + t) * cpc + lc];
        tot = warp_sum_ll(tot);
        if (lane < C) dsmem_st64(&sh.xt[phase][rank][a][t], lane, (unsigned long long)tot);
    }
    cluster.sync();
}

// k-means++ sampling (generateCentersPP): first index ci in [0, N-1) with p - sum_{i<=ci} dist[i] <= 0, else N-1, for the
// three trials of the three attempts (nine warps); dist = distance to the nearest of the ncen chosen centres, whose
// per-chunk sums are seg_a[0 .. cpc) here and whose per-rank totals are sh.xt[phase][r][a][slot_a]. The CTA that owns the
// index publishes it to every CTA.
__device__ __forceinline__ void pp_sample_cluster(cg::cluster_group &cluster, KcShared &sh, const uint32_t *pix, const int *seg,
                                                  int cpc, int nch, int ch_lo, int N, int ncen, int phase, int C, int rank,
                                                  int round, int swarp, int lane)
{
    if (swarp < 9) {
        const int a = swarp / 3, trial = swarp - 3 * a;
        const int *seg_a = seg + a * 4 * cpc;
        const int slot = phase == 0 ? 0 : sh.best[a];
        long long off = 0, total = 0, mytot = 0;
        for (int r = 0; r < C; r++) {
            const long long t = sh.xt[phase][r][a][slot];
            if (r < rank) off += t;
            if (r == rank) mytot = t;
            total += t;
        }
        const double p = __dmul_rn(sh.u[a][round * 3 + trial], (double)total);
        const bool before = rank > 0 && (double)off >= p;                 // an earlier CTA owns it
        const bool mine = !before && (double)(off + mytot) >= p;
        const bool nobody = rank == C - 1 && (double)total < p;           // cannot happen (u <= 1); kept for safety
        int ci = -1;
        if (nobody) ci = N - 1;
        if (mine && nch > 0) {
            // chunk: every lane owns a run of consecutive chunks
            const int per = (nch + 31) >> 5, c_lo = lane * per, c_hi = min(nch, c_lo + per);
            long long loc = 0;
            for (int lc = c_lo; lc < c_hi; lc++) loc += seg_a[lc];
            const long long incl = warp_incl_scan_ll(loc, lane);
            long long run = off + incl - loc;
            int found = -1;
            long long before_chunk = 0;
            for (int lc = c_lo; lc < c_hi; lc++) {
                const long long nb = run + seg_a[lc];
                if ((double)nb >= p) { found = lc; before_chunk = run; break; }
                run = nb;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, found >= 0);
            // bal != 0: (off + mytot) >= p and the chunk sums add up to mytot
            const int src = bal ? __ffs(bal) - 1 : 0;
            const int lc = max(0, __shfl_sync(0xffffffffu, found, src));
            const long long base_sum = __shfl_sync(0xffffffffu, before_chunk, src);
            // pixel inside the chunk: 4 consecutive pixels per lane
            const uint4 v = ((const uint4 *)pix)[lc * 32 + lane];
            const uint32_t px[4] = {v.x, v.y, v.z, v.w};
            const int i0 = (ch_lo + lc) * KC_CH + lane * 4;
            const uint32_t c0 = sh.cenw[a][0], c1 = sh.cenw[a][1];
            int d[4];
            int lsum = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int b = 0;
                if (i0 + q < N) {
                    const int xx = dp4a_uu(px[q], px[q], 0);
                    b = xx + dp4a_uu(c0, c0, 0) - 2 * dp4a_uu(px[q], c0, 0);
                    if (ncen > 1) b = min(b, xx + dp4a_uu(c1, c1, 0) - 2 * dp4a_uu(px[q], c1, 0));
                }
                d[q] = b;
                lsum += b;
            }
            int incl_l = lsum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl_l, o);
                if (lane >= o) incl_l += t;
            }
            long long r2 = base_sum + incl_l - lsum;
            int hit = -1;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                r2 += d[q];
                if (hit < 0 && i0 + q < N && (double)r2 >= p) hit = i0 + q;
            }
            const unsigned hb = __ballot_sync(0xffffffffu, hit >= 0);
            ci = hb ? __shfl_sync(0xffffffffu, hit, __ffs(hb) - 1) : min(N - 1, (ch_lo + lc) * KC_CH + KC_CH - 1);
            ci = min(ci, N - 1);
        }
        if (ci >= 0 && lane < C) dsmem_st32(&sh.cand_idx[round][a][trial], lane, (uint32_t)ci);
    }
    cluster.sync();
}

template <int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) ckb_kmeans_cluster_u8(const uint8_t *__restrict__ imgs, int S,
                                                                      const __grid_constant__ RegionSet regs, int cpc,
                                                                      const uint64_t *__restrict__ rng_states,
                                                                      KmAttempt *__restrict__ results)
{
    constexpr int NW = NT / 32;
    const size_t img_bytes = (size_t)S * S * 3;     // the vector loads never read past the frame they belong to
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    // one cluster per (frame, region): rng_states[unit], results[3 unit + attempt]
    const int unit = blockIdx.x / C, frame = unit / regs.n;
    const Region rg = regs.r[unit - frame * regs.n];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // the serial steps (totals, sampling, centre update) run on the highest-numbered warps: the warp scheduler favours
    // them, and they are the critical path of every CTA that waits at the next barrier
    const int swarp = NW - 1 - warp;
    const int N = rg.N;
    const int nchunk = (N + KC_CH - 1) / KC_CH;
    const int ch_lo = rank * cpc;
    const int nch = max(0, min(cpc, nchunk - ch_lo));

#ifdef KC_TIMING
    long long tk[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = clock64();
    const long long t_begin = t_prev;
#define KC_TICK(slot) do { const long long t_now = clock64(); tk[slot] += t_now - t_prev; t_prev = t_now; } while (0)
#else
#define KC_TICK(slot) do { } while (0)
#endif
    __shared__ KcShared sh;
    extern __shared__ __align__(16) unsigned char kc_dyn[];
    uint32_t *pix = (uint32_t *)kc_dyn;                  // [cpc * 128] packed pixels (byte 3 = 0), zero beyond the region
    uint4 *csum = (uint4 *)(pix + (size_t)cpc * KC_CH);   // [3][cpc] Lloyd: packed sums of clusters 0 and 1 (x | y << 16, z | count << 16)
    uint2 *ctot = (uint2 *)(csum + 3 * cpc);             // [cpc] chunk totals, same packing
    int *seg = (int *)(ctot + cpc);                      // [3][4][cpc] k-means++ chunk sums: current dist, three candidates
    int *cxx = seg + 12 * cpc;                           // [cpc] sum of |x|^2

    const uint8_t *img = imgs + (size_t)frame * S * S * 3;

    // ---- load this CTA's slice of the region (fused "pack") and take the chunk totals. A lane owns 4 consecutive region
    // pixels = 12 consecutive bytes of one image row (unless they straddle the region's right edge): three or four aligned
    // 32-bit loads realigned with funnel shifts, one 128-bit store to shared memory.
    const bool vec_ok = (((uintptr_t)imgs) & 3) == 0;
    if (tid < 3) {
        // cv::RNG draws of attempt `tid`: 1 integer + 6 doubles = 13 draws per attempt
        uint64_t st = rng_states[unit];
        for (int k = 0; k < 13 * tid; k++) rng_next(st);
        const int c0 = (int)(rng_next(st) % (uint32_t)N);
        for (int k = 0; k < 6; k++) sh.u[tid][k] = rng_double(st);
        const uint32_t w0 = region_px(img, S, rg, c0);
        sh.candw[tid][0] = w0;
        sh.candw[tid][1] = w0;
        sh.candw[tid][2] = w0;
        sh.state[tid] = 0;
        sh.iters[tid] = 1;
        sh.best[tid] = 0;
    }
    for (int lc = warp; lc < nch; lc += NW) {
        const int i0 = (ch_lo + lc) * KC_CH + lane * 4;
        const int rem = min(4, max(0, N - i0));
        uint32_t p[4] = {0u, 0u, 0u, 0u};
        const int row = i0 / rg.w, col = i0 - row * rg.w;
        const size_t off = ((size_t)(rg.x0 + row) * S + rg.y0 + col) * 3;      // byte offset inside this frame
        if (rem == 4 && col + 3 < rg.w && vec_ok && off + 16 <= img_bytes) {
            const uint32_t *wp = (const uint32_t *)(img + (off & ~(size_t)3));
            const unsigned sh8 = (unsigned)(off & 3) * 8u;
            const uint32_t a0 = __ldg(wp), a1 = __ldg(wp + 1), a2 = __ldg(wp + 2), a3 = sh8 ? __ldg(wp + 3) : 0u;
            const uint32_t w0 = __funnelshift_r(a0, a1, sh8), w1 = __funnelshift_r(a1, a2, sh8),
                           w2 = __funnelshift_r(a2, a3, sh8);                  // the 12 bytes b0 .. b11
            p[0] = w0 & 0x00ffffffu;
            p[1] = __byte_perm(w0, w1, 0x0543) & 0x00ffffffu;                  // b3 b4 b5
            p[2] = __byte_perm(w1, w2, 0x0432) & 0x00ffffffu;                  // b6 b7 b8
            p[3] = w2 >> 8;                                                    // b9 b10 b11
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (q < rem) p[q] = region_px(img, S, rg, i0 + q);
        }
        ((uint4 *)pix)[lc * 32 + lane] = make_uint4(p[0], p[1], p[2], p[3]);
        uint32_t ta = 0u, tb = (uint32_t)rem << 16;
        int xx = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {   // pixels beyond the region are zero words
            ta += __byte_perm(p[q], 0u, 0x4140);
            tb += __byte_perm(p[q], 0u, 0x4442);
            xx = dp4a_uu(p[q], p[q], xx);
        }
        ta = __reduce_add_sync(0xffffffffu, ta);
        tb = __reduce_add_sync(0xffffffffu, tb);
        xx = __reduce_add_sync(0xffffffffu, xx);
        if (lane == 0) { ctot[lc] = make_uint2(ta, tb); cxx[lc] = xx; }
    }
    __syncthreads();
    KC_TICK(0);   // load + chunk totals
    cluster.sync();   // every CTA of the cluster has started: its shared memory may be written remotely from here on
    KC_TICK(1);   // first cluster barrier (cluster start-up skew)

    // ---- k-means++ seeding (the three attempts side by side)
#ifdef KC_TIMING
    long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tp_prev = clock64();
#define KC_PTICK(slot) do { const long long t_now = clock64(); tp[slot] += t_now - tp_prev; tp_prev = t_now; } while (0)
#else
#define KC_PTICK(slot) do { } while (0)
#endif
    pp_pass_cluster<NW>(cluster, sh, pix, cxx, seg, cpc, nch, ch_lo, N, 0, 1, 0, rank, C, warp, swarp, lane);
    KC_PTICK(0);
    if (tid < 3) sh.cenw[tid][0] = sh.candw[tid][0];
    for (int e = tid; e < 3 * nch; e += NT) {
        const int a = e / nch, lc = e - a * nch;
        seg[(a * 4) * cpc + lc] = seg[(a * 4 + 1) * cpc + lc];
    }
    __syncthreads();
    for (int k = 1; k < (KC_EXP == 4 ? 1 : 3); k++) {
        KC_PTICK(1);
        pp_sample_cluster(cluster, sh, pix, seg, cpc, nch, ch_lo, N, k, k - 1, C, rank, k - 1, swarp, lane);
        KC_PTICK(2);
        if (tid < 9) sh.candw[tid / 3][tid % 3] = region_px(img, S, rg, sh.cand_idx[k - 1][tid / 3][tid % 3]);
        __syncthreads();
        KC_PTICK(3);
        pp_pass_cluster<NW>(cluster, sh, pix, cxx, seg, cpc, nch, ch_lo, N, k, 3, k, rank, C, warp, swarp, lane);
        KC_PTICK(4);
        if (tid < 3) {
            // best trial: strict '<' in trial order
            const int a = tid;
            double s3[3];
#pragma unroll
            for (int t = 0; t < 3; t++) {
                long long s = 0;
                for (int r = 0; r < C; r++) s += sh.xt[k][r][a][t];
                s3[t] = (double)s;
            }
            int best = 0;
            double bs = s3[0];
            if (s3[1] < bs) { bs = s3[1]; best = 1; }
            if (s3[2] < bs) { bs = s3[2]; best = 2; }
            sh.best[a] = best;
            sh.cenw[a][k] = sh.candw[a][best];
        }
        __syncthreads();
        for (int e = tid; e < 3 * nch; e += NT) {
            const int a = e / nch, lc = e - a * nch;
            seg[(a * 4) * cpc + lc] = seg[(a * 4 + 1 + sh.best[a]) * cpc + lc];
        }
        __syncthreads();
    }
    if (tid < 27) sh.cen[tid / 9][tid % 9] = (float)((sh.cenw[tid / 9][(tid % 9) / 3] >> (8 * (tid % 3))) & 0xffu);
    __syncthreads();
    KC_TICK(2);   // k-means++

    // ---- Lloyd iterations: the attempts that have not converged yet advance together
    for (int iter = 1;; iter++) {      // iteration 0 was the seeding
        const int par = iter & 1;
        if (tid < 27 && sh.state[tid / 9] == 0) sh.oldc[tid / 9][tid % 9] = sh.cen[tid / 9][tid % 9];
        if (tid >= 32 && tid < 41 && sh.state[(tid - 32) / 3] == 0) {
            // filter of the pair (ca, cb) of attempt a: F(x) = 64 (|c_a|^2 - |c_b|^2) + x . 128 (c_b - c_a) ~ 128 (d_a(x) - d_b(x)) / 2
            const int a = (tid - 32) / 3, pr = (tid - 32) - 3 * a;
            const int ia = pr == 2 ? 1 : 0, ib = pr == 0 ? 1 : 2;
            double ka = 0.0, kb = 0.0;
            uint32_t lo = 0u, hi = 0u;
            for (int j = 0; j < 3; j++) {
                const float ca = sh.cen[a][3 * ia + j], cb = sh.cen[a][3 * ib + j];
                ka += (double)ca * (double)ca;
                kb += (double)cb * (double)cb;
                const int w = __float2int_rn(__fmul_rn(__fsub_rn(cb, ca), 128.f));
                lo |= (uint32_t)(w & 255) << (8 * j);
                hi |= (uint32_t)((w >> 8) & 255) << (8 * j);
            }
            sh.filt[a].lo[pr] = lo;
            sh.filt[a].hi[pr] = hi;
            sh.filt[a].k[pr] = (int)__double2ll_rn(64.0 * (ka - kb));
        }
        __syncthreads();
        const int act0 = sh.state[0] == 0, act1 = sh.state[1] == 0, act2 = sh.state[2] == 0;

        // labels + packed exact sums per chunk, attempt by attempt
#pragma unroll 1
        for (int a = 0; a < 3; a++) {
            if (!(a == 0 ? act0 : (a == 1 ? act1 : act2))) continue;
            const KcFilter F = sh.filt[a];
            const float *oc = sh.oldc[a];
            const int lsh = 24 + 2 * a;
            uint4 *csum_a = csum + a * cpc;
            for (int lc = warp; lc < nch; lc += NW) {
                uint4 *slot = (uint4 *)pix + lc * 32 + lane;
                uint4 sums;
                if ((ch_lo + lc + 1) * KC_CH <= N) {
                    sums = lloyd_chunk<true>(slot, 0xfu, F, oc, lsh);
                } else {
                    const int rem = N - ((ch_lo + lc) * KC_CH + lane * 4);
                    sums = lloyd_chunk<false>(slot, rem >= 4 ? 0xfu : (rem <= 0 ? 0u : ((1u << rem) - 1u)), F, oc, lsh);
                }
                if (lane == 0) csum_a[lc] = sums;
            }
        }
        KC_TICK(3);   // Lloyd pass (this warp's chunks)
        __syncthreads();
        KC_TICK(4);   // wait for the CTA's other warps
        // slice totals -> every CTA (warp a for attempt a)
        if (swarp < 3 && sh.state[swarp] == 0) {
            const int a = swarp;
            const uint4 *csum_a = csum + a * cpc;
            int t[12];
#pragma unroll
            for (int k = 0; k < 12; k++) t[k] = 0;
            for (int lc = lane; lc < nch; lc += 32) {
                const uint4 c = csum_a[lc];
                const uint2 g = ctot[lc];
                const int x0 = c.x & 0xffffu, y0 = c.x >> 16, z0 = c.y & 0xffffu, n0 = c.y >> 16;
                const int x1 = c.z & 0xffffu, y1 = c.z >> 16, z1 = c.w & 0xffffu, n1 = c.w >> 16;
                t[0] += x0; t[1] += y0; t[2] += z0;
                t[3] += x1; t[4] += y1; t[5] += z1;
                t[6] += (int)(g.x & 0xffffu) - x0 - x1;
                t[7] += (int)(g.x >> 16) - y0 - y1;
                t[8] += (int)(g.y & 0xffffu) - z0 - z1;
                t[9] += n0; t[10] += n1;
                t[11] += (int)(g.y >> 16) - n0 - n1;
            }
            int mine = 0;
#pragma unroll
            for (int k = 0; k < 12; k++) {
                const int s = __reduce_add_sync(0xffffffffu, t[k]);
                if (lane == k) mine = s;
            }
            if (lane < 12)
                for (int r = 0; r < C; r++) dsmem_st32(&sh.xl[par][rank][a][lane], r, (uint32_t)mine);
        }
        cluster.sync();
        KC_TICK(5);   // slice totals + cluster exchange
        if (swarp < 3 && sh.state[swarp] == 0) {
            const int a = swarp;
            int T = 0, P = 0;
            if (lane < 12) {
                for (int r = 0; r < C; r++) {
                    const int s = sh.xl[par][r][a][lane];
                    if (r < rank) P += s;
                    T += s;
                }
                sh.T[a][lane] = T;
                sh.P[a][lane] = P;
            }
            const unsigned tm = __ballot_sync(0xffffffffu, lane < 9 && T > KC_2_24);
            const unsigned fb = __ballot_sync(0xffffffffu, (lane < 9 && T + KC_GUARD >= (1 << 25)) ||
                                                               (lane >= 9 && lane < 12 && T == 0));
            if (lane == 0) {
                sh.tailmask[a] = fb ? 0 : (int)tm;
                if (fb) sh.state[a] = 2;        // identical in every CTA of the cluster
            }
        } else if (swarp < 3 && lane == 0) {
            sh.tailmask[swarp] = 0;
        }
        __syncthreads();
#if KC_EXP == 1
        const int tm0 = 0, tm1 = 0, tm2 = 0;      // ablation: no tail (wrong sums, timing only)
#else
        const int tm0 = sh.tailmask[0], tm1 = sh.tailmask[1], tm2 = sh.tailmask[2];
#endif
        if (tm0 | tm1 | tm2) {
            // ---- sums that leave the exact range ((attempt, chain) pairs one after the other; normally one per attempt).
            // For chain c, rank rc holds the chunk gx in which the float32 sum passes 2^24: one of its warps walks that
            // chunk serially from the exact sum before it; the chunks after gx (to the end of the region, whoever holds
            // them) are shared out over ALL warps of ALL CTAs of the cluster, pixels read through distributed shared
            // memory, and each warp composes the rounding automaton over its run of chunks for both entry parities.
            // Crossing chunks first: the last warp of each CTA searches the chunk sums of the rank that holds the crossing
            // (remote reads, so no extra cluster barrier; one warp per CTA, the DSMEM port is narrow).
            if (swarp < 3) {                       // one warp per attempt; its chains one after the other
                const int a = swarp, tma = a == 0 ? tm0 : (a == 1 ? tm1 : tm2);
#pragma unroll 1
                for (int c = 0; c < 9; c++) {
                    if (!((tma >> c) & 1)) continue;
                    int rc = 0, before_rc = 0;       // rank that holds the crossing, and the exact sum before its slice
                    for (; rc < C - 1; rc++) {
                        const int s = sh.xl[par][rc][a][c];
                        if (before_rc + s > KC_2_24) break;
                        before_rc += s;
                    }
                    // this lane's run of (at most 5) consecutive chunks of rank rc: all the remote reads are issued before
                    // anything depends on them
                    const int nch_rc = max(0, min(cpc, nchunk - rc * cpc));
                    const int per = (nch_rc + 31) >> 5, c_lo = lane * per;
                    int cs5[5];
#pragma unroll
                    for (int i = 0; i < 5; i++) {
                        const int lc = c_lo + i;
                        cs5[i] = 0;
                        if (i < per && lc < nch_rc)
                            cs5[i] = chain_chunk_sum(dsmem_ld128(csum + a * cpc + lc, rc), dsmem_ld64(ctot + lc, rc), c);
                    }
                    const int loc = cs5[0] + cs5[1] + cs5[2] + cs5[3] + cs5[4];
                    int incl = loc;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    int run = before_rc + incl - loc, found = -1, before = 0;
#pragma unroll
                    for (int i = 0; i < 5; i++) {
                        if (found < 0 && i < per && c_lo + i < nch_rc) {
                            const int nb = run + cs5[i];
                            if (nb > KC_2_24) { found = c_lo + i; before = run; }
                            run = nb;
                        }
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, found >= 0);
                    const int src = bal ? __ffs(bal) - 1 : 0;          // bal != 0 by the choice of rc
                    const int xc = max(0, __shfl_sync(0xffffffffu, found, src)), xs = __shfl_sync(0xffffffffu, before, src);
                    if (lane == 0) { sh.rc[a][c] = rc; sh.xc[a][c] = xc; sh.xstart[a][c] = xs; }
                }
            }
            __syncthreads();
            KC_TICK(9);   // tail: crossing search
#pragma unroll 1
            for (int e = 0; e < 27; e++) {
                const int a = e / 9, c = e - 9 * a;
                if (!(((a == 0 ? tm0 : (a == 1 ? tm1 : tm2)) >> c) & 1)) continue;
                const int k = c / 3, sft = 8 * (c - 3 * k);
                const int rc = sh.rc[a][c], xc = sh.xc[a][c], xstart = sh.xstart[a][c];
                const int lsh = 24 + 2 * a;
                const int gx = rc * cpc + xc;                            // global index of the crossing chunk
                if (rank == rc && warp == (e % NW)) {
                    // serial float32 walk of the crossing chunk, in pixel order, from the exact sum before it
                    uint32_t p[4];
                    unsigned valid = 0u;
#pragma unroll
                    for (int q = 0; q < 4; q++) {   // lane-strided: pixels 32 q + lane of the chunk
                        p[q] = pix[xc * KC_CH + q * 32 + lane];
                        if (gx * KC_CH + q * 32 + lane < N) valid |= 1u << q;
                    }
                    float vals[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const bool mem = ((valid >> q) & 1u) && (int)((p[q] >> lsh) & 3u) == k;   // label cached by the pass
                        vals[q] = mem ? (float)((p[q] >> sft) & 0xffu) : 0.f;     // adding +0 for non-members is exact
                    }
                    // the 128 values in pixel order through shared memory (a shuffle per step would put its latency on
                    // the chain): broadcast 128-bit reads, one dependent add per pixel
                    float *wb = sh.walkbuf[warp];
#pragma unroll
                    for (int q = 0; q < 4; q++) wb[q * 32 + lane] = vals[q];
                    __syncwarp();
                    float acc = (float)xstart;
#pragma unroll 8
                    for (int i4 = 0; i4 < KC_CH / 4; i4++) {
                        const float4 w4 = ((const float4 *)wb)[i4];
                        acc = __fadd_rn(acc, w4.x);
                        acc = __fadd_rn(acc, w4.y);
                        acc = __fadd_rn(acc, w4.z);
                        acc = __fadd_rn(acc, w4.w);
                    }
                    __syncwarp();
                    if (lane < C) dsmem_st32(&sh.walk_u[par][a][c], lane, (uint32_t)((int)acc >> 1));
                }
                // automaton over this warp's run of the chunks after the crossing: u -> u + add[u & 1]
                const int tail_n = nchunk - (gx + 1);
                const int perw = (tail_n + C * NW - 1) / (C * NW);
                const int g_lo = gx + 1 + (rank * NW + warp) * perw, g_hi = min(nchunk, g_lo + perw);
                int add0 = 0, add1 = 0;
                for (int g = g_lo; g < g_hi; g++) {
                    const int owner = g / cpc, lc = g - owner * cpc;
                    const uint4 v = dsmem_ld128((const uint4 *)pix + lc * 32 + lane, owner);
                    const uint32_t p[4] = {v.x, v.y, v.z, v.w};
                    const int rem = N - (g * KC_CH + lane * 4);
                    const unsigned valid = rem >= 4 ? 0xfu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
                    // this lane's 4 pixels in order, for entry parity 0 (i0) and 1 (i1)
                    int i0 = 0, i1 = 0;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const bool mem = ((valid >> q) & 1u) && (int)((p[q] >> lsh) & 3u) == k;   // label cached by the pass
                        const int x = (int)((p[q] >> sft) & 0xffu), ha = x >> 1;
                        if (mem) {
                            i0 += ha + ((x & 1) ? ((i0 + ha) & 1) : 0);
                            i1 += ha + ((x & 1) ? ((1 + i1 + ha) & 1) : 0);
                        }
                    }
                    // ordered composition across the lanes (earlier (L) then later (R)): inc_h = L_h + R[(h + L_h) & 1]
                    uint32_t T = (uint32_t)i0 | ((uint32_t)i1 << 16);     // <= 128 * 128 per entry: 16 bits each
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t R = __shfl_down_sync(0xffffffffu, T, d);
                        const uint32_t l0 = T & 0xffffu, l1 = T >> 16, r0 = R & 0xffffu, r1 = R >> 16;
                        const uint32_t n0 = l0 + ((l0 & 1u) ? r1 : r0), n1 = l1 + ((l1 & 1u) ? r0 : r1);
                        T = n0 | (n1 << 16);
                    }
                    const int ch0 = (int)(T & 0xffffu), ch1 = (int)(T >> 16);   // valid on lane 0
                    add0 += (add0 & 1) ? ch1 : ch0;
                    add1 += ((1 + add1) & 1) ? ch1 : ch0;
                }
                if (lane == 0) sh.wt[a][c][warp] = make_int2(add0, add1);
            }
            KC_TICK(10);  // tail: walk, automaton sweep of this warp
            __syncthreads();
            KC_TICK(11);  // tail: wait for the CTA's other warps
            if (tid < 27 && (((tid / 9 == 0 ? tm0 : (tid / 9 == 1 ? tm1 : tm2)) >> (tid % 9)) & 1)) {
                const int a = tid / 9, c = tid % 9;
                int t0 = 0, t1 = 0;
                for (int w = 0; w < NW; w++) {
                    const int2 e2 = sh.wt[a][c][w];
                    t0 += (t0 & 1) ? e2.y : e2.x;
                    t1 += ((1 + t1) & 1) ? e2.y : e2.x;
                }
                for (int r = 0; r < C; r++) dsmem_st64(&sh.tl[par][rank][a][c], r, (unsigned long long)(uint32_t)t0 | ((unsigned long long)(uint32_t)t1 << 32));
            }
            cluster.sync();
            if (tid < 27 && (((tid / 9 == 0 ? tm0 : (tid / 9 == 1 ? tm1 : tm2)) >> (tid % 9)) & 1)) {
                const int a = tid / 9, c = tid % 9;
                int uu = sh.walk_u[par][a][c];
                for (int r = 0; r < C; r++) uu += (uu & 1) ? sh.tl[par][r][a][c].y : sh.tl[par][r][a][c].x;
                sh.T[a][c] = uu;          // half of the float32 sum (an even integer below 2^25)
            }
            __syncthreads();
            KC_TICK(6);   // tail (sums past 2^24)
        }

        // new centres, shift, stop rule per attempt (every CTA computes the same values)
        if (swarp < 3 && lane == 0 && sh.state[swarp] == 0) {
            const int a = swarp, tmask = sh.tailmask[a];
            double max_shift = 0.0;
            for (int k = 0; k < 3; k++) {
                const float scale = __fdiv_rn(1.f, (float)sh.T[a][9 + k]);
                double dist = 0.0;
                for (int j = 0; j < 3; j++) {
                    const int c = 3 * k + j;
                    const float sum = ((tmask >> c) & 1) ? (float)(sh.T[a][c] << 1) : (float)sh.T[a][c];
                    const float cn = __fmul_rn(sum, scale);
                    sh.cen[a][c] = cn;
                    const double t = (double)__fsub_rn(cn, sh.oldc[a][c]);
                    dist = __dadd_rn(dist, __dmul_rn(t, t));
                }
                max_shift = dist > max_shift ? dist : max_shift;
            }
            sh.iters[a] = iter + 1;
            if ((iter + 1 == KM_MAX_ITER) || (max_shift <= KM_EPS2) || KC_EXP == 3) sh.state[a] = 1;
        }
        __syncthreads();
        KC_TICK(7);   // centres
        if (sh.state[0] != 0 && sh.state[1] != 0 && sh.state[2] != 0) break;
    }

    // ---- compactness of the converged attempts: labels stay those assigned against oldc; distances to the final centres
    double *cdbl = (double *)seg;          // [3][cpc] per-chunk sums (the k-means++ chunk sums are no longer needed)
#pragma unroll 1
    for (int a = 0; a < 3; a++) {
        if (sh.state[a] != 1 || KC_EXP == 2) continue;
        const int lsh = 24 + 2 * a;
        float nc[9];
#pragma unroll
        for (int k = 0; k < 9; k++) nc[k] = sh.cen[a][k];
        for (int lc = warp; lc < nch; lc += NW) {
            const uint4 v = ((const uint4 *)pix)[lc * 32 + lane];
            const uint32_t p[4] = {v.x, v.y, v.z, v.w};
            const int rem = (ch_lo + lc + 1) * KC_CH <= N ? 4 : N - ((ch_lo + lc) * KC_CH + lane * 4);
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (q < rem) {
                    const float3 x = unpack_px(p[q]);
                    const uint32_t lab = (p[q] >> lsh) & 3u;      // assigned against oldc by the attempt's last pass
                    const float cx = lab == 0u ? nc[0] : (lab == 1u ? nc[3] : nc[6]);
                    const float cy = lab == 0u ? nc[1] : (lab == 1u ? nc[4] : nc[7]);
                    const float cz = lab == 0u ? nc[2] : (lab == 1u ? nc[5] : nc[8]);
                    acc += (double)dist3(x.x, x.y, x.z, cx, cy, cz);
                }
            }
            acc = warp_sum_d(acc);
            if (lane == 0) cdbl[a * cpc + lc] = acc;
        }
    }
    __syncthreads();
    if (swarp < 3 && sh.state[swarp] == 1) {      // fixed order: the result does not depend on which warp took which chunk
        double t = 0.0;
        for (int lc = lane; lc < nch; lc += 32) t += cdbl[swarp * cpc + lc];
        t = warp_sum_d(t);
        if (lane == 0) dsmem_st64(&sh.xcomp[rank][swarp], 0, (unsigned long long)__double_as_longlong(t));
    }
    cluster.sync();
    if (rank == 0 && tid < 3) {
        const int a = tid;
        KmAttempt &res = results[unit * 3 + a];
        if (sh.state[a] == 1) {
            double t = 0.0;
            for (int r = 0; r < C; r++) t += sh.xcomp[r][a];
            res.compactness = t;
            for (int k = 0; k < 9; k++) { res.centers[k] = sh.cen[a][k]; res.old_centers[k] = sh.oldc[a][k]; }
            res.n_fix = 0;
            for (int q = 0; q < 2; q++) { res.fix_idx[q] = 0; res.fix_k[q] = 0; }
            res.iters = sh.iters[a];
        } else {
            res.iters = KM_ITERS_FALLBACK;
        }
    }
    KC_TICK(8);   // compactness
#ifdef KC_TIMING
    if (tid == 0 && unit == (gridDim.x / C > 50 ? 50 : 0))
        printf("rank %d pp: passA %lld  copy %lld  sample %lld  fetch %lld  passBC %lld\n", rank, tp[0], tp[1], tp[2], tp[3], tp[4]);
    if (tid == 0 && unit == (gridDim.x / C > 50 ? 50 : 0))
        printf("rank %d iters %d %d %d: load %lld  sync0 %lld  pp %lld  pass %lld  wait %lld  xchg %lld  tail %lld (search %lld sweep %lld wait %lld)  centres %lld  compact %lld  total %lld\n",
               rank, sh.iters[0], sh.iters[1], sh.iters[2], tk[0], tk[1], tk[2], tk[3], tk[4], tk[5], tk[6], tk[9], tk[10], tk[11], tk[7], tk[8], clock64() - t_begin);
#endif
}

// -------------------------------------------------------------------------------------------------------------- launcher
// Cluster size: at most 141 chunks (72 KB of pixels) per CTA so that two 512-thread CTAs fit an SM, and no fewer CTAs than
// keep ~6000 pixels of work each.
static int kc_cluster_size(int N)
{
    const int nchunk = (N + KC_CH - 1) / KC_CH;
    int C = 1;
    while (C < KC_MAXC && (nchunk + C - 1) / C > 141) C *= 2;
    while (C < KC_MAXC && (nchunk + C - 1) / C > 48) C *= 2;
    return C;
}

#define KC_BYTES_PER_CHUNK (KC_CH * 4 + 3 * 16 + 8 + 12 * 4 + 4)

// n frames x regs.n regions: one cluster per (frame, region); d_rng_states / d_results are indexed by frame * regs.n + region
int ckb_launch_kmeans_cluster(ckb_ctx *ctx, const uint8_t *d_imgs, int n, const RegionSet &regs, const uint64_t *d_rng_states,
                              KmAttempt *d_results, cudaStream_t st)
{
    constexpr int NT = 512;
    int maxN = 1;
    for (int i = 0; i < regs.n; i++) maxN = regs.r[i].N > maxN ? regs.r[i].N : maxN;
    const int C = kc_cluster_size(maxN);                 // sized for the largest region of the call
    const int nchunk = (maxN + KC_CH - 1) / KC_CH;
    const int cpc = (nchunk + C - 1) / C;
    if (cpc > 141) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_find_stones: region too large for a cluster of %d", C);
    const size_t dyn = (size_t)cpc * KC_BYTES_PER_CHUNK;
    static bool attr_set[64] = {false};
    if (ctx->device >= 0 && ctx->device < 64 && !attr_set[ctx->device]) {
        CKB_CUDA(ctx, cudaFuncSetAttribute(ckb_kmeans_cluster_u8<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           141 * KC_BYTES_PER_CHUNK));
        attr_set[ctx->device] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n * regs.n * C), 1, 1);
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int S = ctx->S;
    CKB_CUDA(ctx, cudaLaunchKernelEx(&cfg, ckb_kmeans_cluster_u8<NT>, d_imgs, S, regs, cpc, d_rng_states, d_results));
    CKB_LAUNCH_CHECK(ctx, "ckb_kmeans_cluster");
    return CKB_OK;
}
