// K3 for uint8 images — one THREAD-BLOCK CLUSTER per (frame, attempt), region pixels resident in distributed shared memory.
//
// Replaces cv2.kmeans(pixels, 3, None, (TERM_CRITERIA_EPS, 15, 3), 3, KMEANS_PP_CENTERS)   (sf_clustering.py:103-104)
// for the canonical uint8 image, bit for bit (same contract as ckb_kmeans_attempt in kmeans.cu, which stays the path for
// float32 images and the fallback for the rare cases this kernel declines).
//
// Why a second kernel. One CTA per (frame, attempt) left 3 n work units for 148 SMs (64 frames: 1.3 waves) and re-read the
// region from L2 in every one of its ~8 passes. Here a cluster of C CTAs (C = 8 for a full board) owns one unit: CTA r
// keeps pixels [r L, (r+1) L) of the region in its shared memory (packed uchar4, loaded once, straight from the image: no
// pack kernel, no scratch), every pass reads shared memory only, and the CTAs exchange their partial sums through
// distributed shared memory (remote st.shared::cluster + barrier.cluster), one exchange per pass. 3 n C CTAs of 512
// threads, two per SM, balance the chip.
//
// Arithmetic. For uint8 pixels every quantity of k-means++ is an integer: distances by dp4a, sums exact in any order.
// Lloyd iterations: the label of a pixel is the first minimum of three float32 distances (8 roundings each). They are
// only evaluated where they can matter: a fixed-point filter F_ab(x) ~ 128 (d_a(x) - d_b(x)) / 2 (weights 128 (c_b - c_a)
// rounded to integers and split into two bytes: two dp4a + one shift-add per pair of centres) is within 3.06 of the
// float32 value, so |F_ab| > 3.5 * 128 decides the comparison; a warp in which some pixel is left undecided recomputes
// that group of pixels with the exact float32 chain. Centre sums are OpenCV's sequential float32 sums in pixel order:
// exact integers up to 2^24 (taken as packed 16-bit warp reductions per 128-pixel chunk, third cluster = chunk total
// minus the other two), the chunk in which a sum crosses 2^24 is walked serially by one warp, and in [2^24, 2^25) float32
// addition of an integer is a two-state automaton on the parity of the half-sum; every CTA composes the automaton over
// its own chunks for both entry parities in parallel, so the hand-over between CTAs is a table lookup. A unit with an
// empty cluster or a sum within reach of 2^25 is handed to ckb_kmeans_attempt (KmAttempt.iters = KM_ITERS_FALLBACK).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "kmeans_common.cuh"

namespace cg = cooperative_groups;

#define KC_CH 128                 // pixels per chunk: one warp iteration, 4 consecutive pixels per lane
#define KC_MAXC 16                // largest cluster size used (8 is the portable limit; 16 needs the opt-in attribute)
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
    long long xt[3][KC_MAXC][4];  // k-means++: per-rank totals of the candidate distance sums (pass A, B, C)
    int cand_idx[2][3];           // sampled candidates (round 1, 2), written by the owning CTA into every CTA
    int xl[2][KC_MAXC][12];       // Lloyd: per-rank slice sums (9 chains) and counts (3), double buffered by iteration
    int2 tl[2][KC_MAXC][9];       // Lloyd tail: per-rank automaton tables (add for entry parity 0 / 1) or the value itself
    double xcomp[KC_MAXC];        // compactness partials (rank 0's copy is the one that is read)
    KcFilter filt;
    float cen[9], oldc[9];
    uint32_t cenw[3];             // k-means++ centres as packed pixels
    uint32_t candw[3];
    int mine[12];
    int T[12], P[12];             // totals over the cluster / prefix before this rank
    int tailmask, fallback, flag;
    int next[2];                  // chunk claim counters, used alternately by successive passes
    int rc[9], xc[9], xstart[9];  // per tail chain: rank / local chunk of the 2^24 crossing, exact sum before that chunk
    int walk_u[2][9];             // half of the float32 sum after the crossing chunk (written by the rank that walks it)
    double u[6];                  // the attempt's six uniform draws (k-means++ sampling)
    int2 wt[9][16];               // per-warp automaton tables
    double red_d[16];
    long long red_l[16];
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

// Dynamic chunk claiming: the warps of a CTA take the CTA's chunks first come, first served (one shared-memory atomic per
// chunk). The hardware warp scheduler is priority based, so with a static split the CTA - and through the next cluster
// barrier the whole cluster - waits for its least favoured warp; every per-chunk result lands in a per-chunk slot, so
// the outcome does not depend on who processed which chunk.
__device__ __forceinline__ int claim_chunk(int *counter, int lane)
{
    int lc = 0;
    if (lane == 0) lc = atomicAdd(counter, 1);
    return __shfl_sync(0xffffffffu, lc, 0);
}

// Labels of four pixels against the centres `oc` (first minimum of the float32 distances): k0[q] / k1[q] = pixel q belongs
// to cluster 0 / 1 (neither: cluster 2). `valid` has bit q set for pixels inside the region. Warp-collective.
__device__ __forceinline__ void classify4(const uint32_t p[4], unsigned valid, const KcFilter &F, const float *oc,
                                          bool k0[4], bool k1[4])
{
    bool unsure = false;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int f01 = dp4a_uu(p[q], F.lo[0], F.k[0]) + (dp4a_us(p[q], F.hi[0], 0) * 256);
        const int f02 = dp4a_uu(p[q], F.lo[1], F.k[1]) + dp4a_us(p[q], F.hi[1], 0) * 256;
        const int f12 = dp4a_uu(p[q], F.lo[2], F.k[2]) + dp4a_us(p[q], F.hi[2], 0) * 256;
        const bool a = f01 < -KC_T && f02 < -KC_T;
        const bool b = f01 > KC_T && f12 < -KC_T;
        const bool c = f02 > KC_T && f12 > KC_T;
        const bool v = (valid >> q) & 1u;
        k0[q] = a && v;
        k1[q] = b && v;
        unsure |= v && !(a || b || c);
    }
    if (__any_sync(0xffffffffu, unsure)) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int lab = argmin3(unpack_px(p[q]), oc);
            const bool v = (valid >> q) & 1u;
            k0[q] = v && lab == 0;
            k1[q] = v && lab == 1;
        }
    }
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

// k-means++ pass over this CTA's chunks: per chunk, for each of ncand candidates, sum over the pixels of
// min(d(x, cand), d(x, nearest chosen centre)) -> seg[(1 + t) * cpc + chunk]; per-CTA totals -> every CTA's xt[phase][rank][t].
template <int NW>
__device__ __forceinline__ void pp_pass_cluster(cg::cluster_group &cluster, KcShared &sh, const uint32_t *pix, const int *cxx,
                                                int *seg, int cpc, int nch, int ch_lo, int N, int ncen, int ncand, int phase,
                                                int rank, int C, int warp, int lane)
{
    uint32_t cw[3], bw[2] = {0u, 0u};
    int cc[3], bc[2] = {0, 0};
#pragma unroll
    for (int t = 0; t < 3; t++) {
        cw[t] = sh.candw[t < ncand ? t : 0];
        cc[t] = dp4a_uu(cw[t], cw[t], 0);
    }
#pragma unroll
    for (int b = 0; b < 2; b++) {
        if (b < ncen) {
            bw[b] = sh.cenw[b];
            bc[b] = dp4a_uu(bw[b], bw[b], 0);
        }
    }
    for (;;) {
        const int lc = claim_chunk(&sh.next[phase & 1], lane);
        if (lc >= nch) break;
        const uint4 v = ((const uint4 *)pix)[lc * 32 + lane];
        const uint32_t p[4] = {v.x, v.y, v.z, v.w};
        const int rem = N - ((ch_lo + lc) * KC_CH + lane * 4);
        int acc[3] = {0, 0, 0};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (q < rem) {
                // distances relative to |x|^2, which is added per chunk (cxx): d' = |c|^2 - 2 x.c
                int base = 0x7fffffff;
                if (ncen > 0) base = bc[0] - 2 * dp4a_uu(p[q], bw[0], 0);
                if (ncen > 1) base = min(base, bc[1] - 2 * dp4a_uu(p[q], bw[1], 0));
#pragma unroll
                for (int t = 0; t < 3; t++)
                    if (t < ncand) acc[t] += min(cc[t] - 2 * dp4a_uu(p[q], cw[t], 0), base);
            }
        }
#pragma unroll
        for (int t = 0; t < 3; t++) {
            if (t < ncand) {
                const int s = __reduce_add_sync(0xffffffffu, acc[t]);
                if (lane == 0) seg[(1 + t) * cpc + lc] = s + cxx[lc];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) sh.next[phase & 1] = 0;      // free for the pass after the next one
    if (warp < ncand) {
        long long tot = 0;
        for (int lc = lane; lc < nch; lc += 32) tot += seg[(1 + warp) * cpc + lc];
        tot = warp_sum_ll(tot);
        if (lane < C) *cluster.map_shared_rank(&sh.xt[phase][rank][warp], lane) = tot;
    }
    cluster.sync();
}

// k-means++ sampling (generateCentersPP): first index ci in [0, N-1) with p - sum_{i<=ci} dist[i] <= 0, else N-1, for the
// three trials of one round; dist = distance to the nearest of the ncen chosen centres, whose per-chunk sums are
// seg[0..cpc) here and whose per-rank totals are sh.xt[phase][r][slot]. The CTA that owns the index publishes it to every CTA.
__device__ __forceinline__ void pp_sample_cluster(cg::cluster_group &cluster, KcShared &sh, const uint32_t *pix, const int *seg,
                                                  int nch, int ch_lo, int N, int ncen, int phase, int slot, int C,
                                                  int rank, int round, int warp, int lane)
{
    if (warp < 3) {
        long long off = 0, total = 0, mytot = 0;
        for (int r = 0; r < C; r++) {
            const long long t = sh.xt[phase][r][slot];
            if (r < rank) off += t;
            if (r == rank) mytot = t;
            total += t;
        }
        const double p = __dmul_rn(sh.u[round * 3 + warp], (double)total);
        const bool before = rank > 0 && (double)off >= p;                 // an earlier CTA owns it
        const bool mine = !before && (double)(off + mytot) >= p;
        const bool nobody = rank == C - 1 && (double)total < p;           // cannot happen (u <= 1); kept for safety
        int ci = -1;
        if (nobody) ci = N - 1;
        if (mine && nch > 0) {
            // chunk: every lane owns a run of consecutive chunks
            const int per = (nch + 31) >> 5, c_lo = lane * per, c_hi = min(nch, c_lo + per);
            long long loc = 0;
            for (int lc = c_lo; lc < c_hi; lc++) loc += seg[lc];
            const long long incl = warp_incl_scan_ll(loc, lane);
            long long run = off + incl - loc;
            int found = -1;
            long long before_chunk = 0;
            for (int lc = c_lo; lc < c_hi; lc++) {
                const long long nb = run + seg[lc];
                if ((double)nb >= p) { found = lc; before_chunk = run; break; }
                run = nb;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, found >= 0);
            // bal != 0: (off + mytot) >= p and the chunk sums add up to mytot
            const int src = __ffs(bal) - 1;
            const int lc = __shfl_sync(0xffffffffu, found, src);
            const long long base_sum = __shfl_sync(0xffffffffu, before_chunk, src);
            // pixel inside the chunk: 4 consecutive pixels per lane
            const uint4 v = ((const uint4 *)pix)[lc * 32 + lane];
            const uint32_t px[4] = {v.x, v.y, v.z, v.w};
            const int i0 = (ch_lo + lc) * KC_CH + lane * 4;
            int d[4];
            int lsum = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int b = 0;
                if (i0 + q < N) {
                    const int xx = dp4a_uu(px[q], px[q], 0);
                    const uint32_t c0 = sh.cenw[0];
                    b = xx + dp4a_uu(c0, c0, 0) - 2 * dp4a_uu(px[q], c0, 0);
                    if (ncen > 1) {
                        const uint32_t c1 = sh.cenw[1];
                        b = min(b, xx + dp4a_uu(c1, c1, 0) - 2 * dp4a_uu(px[q], c1, 0));
                    }
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
        if (ci >= 0 && lane < C) *cluster.map_shared_rank(&sh.cand_idx[round][warp], lane) = ci;
    }
    cluster.sync();
}

template <int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) ckb_kmeans_cluster_u8(const uint8_t *__restrict__ imgs, int S, Region rg, int cpc,
                                                              const uint64_t *__restrict__ rng_states,
                                                              KmAttempt *__restrict__ results)
{
    constexpr int NW = NT / 32;
    const size_t img_bytes = (size_t)S * S * 3;     // the vector loads never read past the frame they belong to
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int unit = blockIdx.x / C, frame = unit / 3, attempt = unit - 3 * frame;
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
    uint32_t *pix = (uint32_t *)kc_dyn;                 // [cpc * 128] packed pixels (byte 3 = 0), zero beyond the region
    uint4 *csum = (uint4 *)(pix + (size_t)cpc * KC_CH);  // [cpc] Lloyd: packed sums of clusters 0 and 1 (x | y << 16, z | count << 16)
    uint2 *ctot = (uint2 *)(csum + cpc);                // [cpc] chunk totals, same packing
    int *seg = (int *)(ctot + cpc);                     // [4][cpc] k-means++ chunk sums: current dist, three candidates
    int *cxx = seg + 4 * cpc;                           // [cpc] sum of |x|^2

    const uint8_t *img = imgs + (size_t)frame * S * S * 3;

    // ---- load this CTA's slice of the region (fused "pack") and take the chunk totals. A lane owns 4 consecutive region
    // pixels = 12 consecutive bytes of one image row (unless they straddle the region's right edge): three or four aligned
    // 32-bit loads realigned with funnel shifts, one 128-bit store to shared memory.
    const bool vec_ok = (((uintptr_t)imgs) & 3) == 0;
    if (tid == 0) { sh.next[0] = 0; sh.next[1] = 0; sh.fallback = 0; }
    __syncthreads();
    // cv::RNG draws of this attempt: 1 integer + 6 doubles = 13 draws
    uint64_t st = rng_states[frame];
    for (int k = 0; k < 13 * attempt; k++) rng_next(st);
    const int c0 = (int)(rng_next(st) % (uint32_t)N);
    if (tid == 0) {
        for (int k = 0; k < 6; k++) sh.u[k] = rng_double(st);
        const uint32_t w0 = region_px(img, S, rg, c0);
        sh.candw[0] = w0;
        sh.candw[1] = w0;
        sh.candw[2] = w0;
        sh.fallback = 0;
    }
    for (;;) {
        const int lc = claim_chunk(&sh.next[1], lane);
        if (lc >= nch) break;
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
    if (tid == 0) sh.next[1] = 0;
    KC_TICK(0);   // load + chunk totals
    cluster.sync();   // every CTA of the cluster has started: its shared memory may be written remotely from here on
    KC_TICK(1);   // first cluster barrier (cluster start-up skew)

    // ---- k-means++ seeding
    pp_pass_cluster<NW>(cluster, sh, pix, cxx, seg, cpc, nch, ch_lo, N, 0, 1, 0, rank, C, swarp, lane);
    if (tid == 0) sh.cenw[0] = sh.candw[0];
    for (int lc = tid; lc < nch; lc += NT) seg[lc] = seg[cpc + lc];
    int best = 0;
    __syncthreads();
    for (int k = 1; k < 3; k++) {
        pp_sample_cluster(cluster, sh, pix, seg, nch, ch_lo, N, k, k - 1, best, C, rank, k - 1, swarp, lane);
        if (tid < 3) sh.candw[tid] = region_px(img, S, rg, sh.cand_idx[k - 1][tid]);
        __syncthreads();
        pp_pass_cluster<NW>(cluster, sh, pix, cxx, seg, cpc, nch, ch_lo, N, k, 3, k, rank, C, swarp, lane);
        // best trial: strict '<' in trial order
        double s3[3];
#pragma unroll
        for (int t = 0; t < 3; t++) {
            long long s = 0;
            for (int r = 0; r < C; r++) s += sh.xt[k][r][t];
            s3[t] = (double)s;
        }
        best = 0;
        if (s3[1] < s3[0]) best = 1;
        if (s3[2] < s3[best]) best = 2;
        if (tid == 0) sh.cenw[k] = sh.candw[best];
        for (int lc = tid; lc < nch; lc += NT) seg[lc] = seg[(1 + best) * cpc + lc];
        __syncthreads();
    }
    if (tid < 9) sh.cen[tid] = (float)((sh.cenw[tid / 3] >> (8 * (tid % 3))) & 0xffu);
    __syncthreads();
    KC_TICK(2);   // k-means++

    // ---- Lloyd iterations
    int iter = 1;   // iteration 0 was the seeding
    for (;;) {
        const int par = iter & 1;
        if (tid < 9) sh.oldc[tid] = sh.cen[tid];
        if (tid < 3) {
            // filter of the pair (a, b): F(x) = 64 (|c_a|^2 - |c_b|^2) + x . 128 (c_b - c_a)  ~  128 (d_a(x) - d_b(x)) / 2
            const int a = tid == 2 ? 1 : 0, b = tid == 0 ? 1 : 2;
            double ka = 0.0, kb = 0.0;
            uint32_t lo = 0u, hi = 0u;
            for (int j = 0; j < 3; j++) {
                const float ca = sh.cen[3 * a + j], cb = sh.cen[3 * b + j];
                ka += (double)ca * (double)ca;
                kb += (double)cb * (double)cb;
                const int w = __float2int_rn(__fmul_rn(__fsub_rn(cb, ca), 128.f));
                lo |= (uint32_t)(w & 255) << (8 * j);
                hi |= (uint32_t)((w >> 8) & 255) << (8 * j);
            }
            sh.filt.lo[tid] = lo;
            sh.filt.hi[tid] = hi;
            sh.filt.k[tid] = (int)__double2ll_rn(64.0 * (ka - kb));
        }
        __syncthreads();
        KcFilter F = sh.filt;
        float oc[9];
#pragma unroll
        for (int k = 0; k < 9; k++) oc[k] = sh.oldc[k];

        // labels + packed exact sums per chunk
        for (;;) {
            const int lc = claim_chunk(&sh.next[par], lane);
            if (lc >= nch) break;
            const uint4 v = ((const uint4 *)pix)[lc * 32 + lane];
            const uint32_t p[4] = {v.x, v.y, v.z, v.w};
            const int rem = N - ((ch_lo + lc) * KC_CH + lane * 4);
            const unsigned valid = rem >= 4 ? 0xfu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
            bool k0[4], k1[4];
            classify4(p, valid, F, oc, k0, k1);
            uint32_t a0 = 0u, b0 = 0u, a1 = 0u, b1 = 0u;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t pa = __byte_perm(p[q], 0u, 0x4140);       // x | y << 16
                const uint32_t pb = __byte_perm(p[q], 0x100u, 0x4542);   // z | 1 << 16
                if (k0[q]) { a0 += pa; b0 += pb; }
                if (k1[q]) { a1 += pa; b1 += pb; }
            }
            a0 = __reduce_add_sync(0xffffffffu, a0);
            b0 = __reduce_add_sync(0xffffffffu, b0);
            a1 = __reduce_add_sync(0xffffffffu, a1);
            b1 = __reduce_add_sync(0xffffffffu, b1);
            if (lane == 0) csum[lc] = make_uint4(a0, b0, a1, b1);
        }
        KC_TICK(3);   // Lloyd pass (this warp's chunks)
        __syncthreads();
        KC_TICK(4);   // wait for the CTA's other warps
        if (tid == 0) sh.next[par] = 0;
        // slice totals -> every CTA
        if (swarp == 0) {
            int t[12];
#pragma unroll
            for (int k = 0; k < 12; k++) t[k] = 0;
            for (int lc = lane; lc < nch; lc += 32) {
                const uint4 c = csum[lc];
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
#pragma unroll
            for (int k = 0; k < 12; k++) {
                const int s = __reduce_add_sync(0xffffffffu, t[k]);
                if (lane == 0) sh.mine[k] = s;
            }
            __syncwarp();
            if (lane < 12) {
                const int val = sh.mine[lane];
                for (int r = 0; r < C; r++) cluster.map_shared_rank(&sh.xl[par][rank][0], r)[lane] = val;
            }
        }
        cluster.sync();
        KC_TICK(5);   // slice totals + cluster exchange
        if (swarp == 0) {
            int T = 0, P = 0;
            if (lane < 12) {
                for (int r = 0; r < C; r++) {
                    const int s = sh.xl[par][r][lane];
                    if (r < rank) P += s;
                    T += s;
                }
                sh.T[lane] = T;
                sh.P[lane] = P;
            }
            const unsigned tm = __ballot_sync(0xffffffffu, lane < 9 && T > KC_2_24);
            const unsigned fb = __ballot_sync(0xffffffffu, (lane < 9 && T + KC_GUARD >= (1 << 25)) ||
                                                               (lane >= 9 && lane < 12 && T == 0));
            if (lane == 0) { sh.tailmask = (int)tm; if (fb) sh.fallback = 1; }
        }
        __syncthreads();
        if (sh.fallback) break;       // identical in every CTA of the cluster
        const int tailmask = sh.tailmask;
        if (tailmask) {
            // ---- sums that leave the exact range (one chain after the other; normally one). For chain c, rank rc holds the
            // chunk gx in which the float32 sum passes 2^24: one of its warps walks that chunk serially from the exact sum
            // before it; the chunks after gx (to the end of the region, whoever holds them) are shared out over ALL warps
            // of ALL CTAs of the cluster, pixels read through distributed shared memory, and each warp composes the
            // rounding automaton over its run of chunks for both entry parities.
            // crossing chunk of every tail chain: the last warp of each CTA searches the chunk sums of the rank that holds
            // the crossing (remote reads, so no extra cluster barrier; one warp per CTA, the DSMEM port is narrow)
            if (warp == NW - 1) {
#pragma unroll 1
                for (int c = 0; c < 9; c++) {
                    if (!((tailmask >> c) & 1)) continue;
                    int rc = 0, before_rc = 0;       // rank that holds the crossing, and the exact sum before its slice
                    for (; rc < C - 1; rc++) {
                        const int s = sh.xl[par][rc][c];
                        if (before_rc + s > KC_2_24) break;
                        before_rc += s;
                    }
                    const uint4 *r_csum = cluster.map_shared_rank(csum, rc);
                    const uint2 *r_ctot = cluster.map_shared_rank(ctot, rc);
                    const int nch_rc = max(0, min(cpc, nchunk - rc * cpc));
                    const int per = (nch_rc + 31) >> 5, c_lo = lane * per, c_hi = min(nch_rc, c_lo + per);
                    int loc = 0;
                    for (int lc = c_lo; lc < c_hi; lc++) loc += chain_chunk_sum(r_csum[lc], r_ctot[lc], c);
                    int incl = loc;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    int run = before_rc + incl - loc, found = -1, before = 0;
                    for (int lc = c_lo; lc < c_hi; lc++) {
                        const int nb = run + chain_chunk_sum(r_csum[lc], r_ctot[lc], c);
                        if (nb > KC_2_24) { found = lc; before = run; break; }
                        run = nb;
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, found >= 0);
                    const int src = bal ? __ffs(bal) - 1 : 0;          // bal != 0 by the choice of rc
                    const int xc = __shfl_sync(0xffffffffu, found, src), xs = __shfl_sync(0xffffffffu, before, src);
                    if (lane == 0) { sh.rc[c] = rc; sh.xc[c] = xc; sh.xstart[c] = xs; }
                }
            }
            __syncthreads();
            KC_TICK(9);   // tail: crossing search
#pragma unroll 1
            for (int c = 0; c < 9; c++) {
                if (!((tailmask >> c) & 1)) continue;
                const int k = c / 3, sft = 8 * (c - 3 * k);
                const int rc = sh.rc[c], xc = sh.xc[c], xstart = sh.xstart[c];
                const int gx = rc * cpc + xc;                            // global index of the crossing chunk
                if (rank == rc && warp == (c % NW)) {
                    // serial float32 walk of the crossing chunk, in pixel order, from the exact sum before it
                    uint32_t p[4];
                    unsigned valid = 0u;
#pragma unroll
                    for (int q = 0; q < 4; q++) {   // lane-strided: pixels 32 q + lane of the chunk
                        p[q] = pix[xc * KC_CH + q * 32 + lane];
                        if (gx * KC_CH + q * 32 + lane < N) valid |= 1u << q;
                    }
                    bool k0[4], k1[4];
                    classify4(p, valid, F, oc, k0, k1);
                    float vals[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const bool v = (valid >> q) & 1u;
                        const bool mem = v && (k == 0 ? k0[q] : (k == 1 ? k1[q] : !(k0[q] || k1[q])));
                        vals[q] = mem ? (float)((p[q] >> sft) & 0xffu) : 0.f;     // adding +0 for non-members is exact
                    }
                    float acc = (float)xstart;
#pragma unroll
                    for (int q = 0; q < 4; q++)
#pragma unroll
                        for (int l = 0; l < 32; l++) acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, vals[q], l));
                    if (lane < C) *cluster.map_shared_rank(&sh.walk_u[par][c], lane) = (int)acc >> 1;
                }
                // automaton over this warp's run of the chunks after the crossing: u -> u + add[u & 1]
                const int tail_n = nchunk - (gx + 1);
                const int perw = (tail_n + C * NW - 1) / (C * NW);
                const int g_lo = gx + 1 + (rank * NW + warp) * perw, g_hi = min(nchunk, g_lo + perw);
                int add0 = 0, add1 = 0;
                for (int g = g_lo; g < g_hi; g++) {
                    const int owner = g / cpc, lc = g - owner * cpc;
                    const uint4 v = ((const uint4 *)cluster.map_shared_rank(pix, owner))[lc * 32 + lane];
                    const uint32_t p[4] = {v.x, v.y, v.z, v.w};
                    const int rem = N - (g * KC_CH + lane * 4);
                    const unsigned valid = rem >= 4 ? 0xfu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
                    bool k0[4], k1[4];
                    classify4(p, valid, F, oc, k0, k1);
                    // this lane's 4 pixels in order, for entry parity 0 (i0) and 1 (i1): float32 addition of an integer x
                    // to the even integer 2u is u += (x >> 1) + (x odd ? (u + (x >> 1)) & 1 : 0)
                    int i0 = 0, i1 = 0;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const bool v = (valid >> q) & 1u;
                        const bool mem = v && (k == 0 ? k0[q] : (k == 1 ? k1[q] : !(k0[q] || k1[q])));
                        const int x = (int)((p[q] >> sft) & 0xffu), a = x >> 1;
                        if (mem) {
                            i0 += a + ((x & 1) ? ((i0 + a) & 1) : 0);
                            i1 += a + ((x & 1) ? ((1 + i1 + a) & 1) : 0);
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
                if (lane == 0) sh.wt[c][warp] = make_int2(add0, add1);
            }
            KC_TICK(10);  // tail: crossing search, walk, automaton sweep of this warp
            __syncthreads();
            KC_TICK(11);  // tail: wait for the CTA's other warps
            if (tid < 9 && ((tailmask >> tid) & 1)) {
                const int c = tid;
                int t0 = 0, t1 = 0;
                for (int w = 0; w < NW; w++) {
                    const int2 e = sh.wt[c][w];
                    t0 += (t0 & 1) ? e.y : e.x;
                    t1 += ((1 + t1) & 1) ? e.y : e.x;
                }
                for (int r = 0; r < C; r++) *cluster.map_shared_rank(&sh.tl[par][rank][c], r) = make_int2(t0, t1);
            }
            cluster.sync();
            if (tid < 9 && ((tailmask >> tid) & 1)) {
                const int c = tid;
                int uu = sh.walk_u[par][c];
                for (int r = 0; r < C; r++) uu += (uu & 1) ? sh.tl[par][r][c].y : sh.tl[par][r][c].x;
                sh.T[c] = uu;          // half of the float32 sum (an even integer below 2^25)
            }
            __syncthreads();
            KC_TICK(6);   // tail (sums past 2^24)
        }

        // new centres, shift, stop rule (every CTA computes the same values)
        if (tid == NT - 32) {
            double max_shift = 0.0;
            for (int k = 0; k < 3; k++) {
                const float scale = __fdiv_rn(1.f, (float)sh.T[9 + k]);
                double dist = 0.0;
                for (int j = 0; j < 3; j++) {
                    const int c = 3 * k + j;
                    const float sum = ((tailmask >> c) & 1) ? (float)(sh.T[c] << 1) : (float)sh.T[c];
                    const float cn = __fmul_rn(sum, scale);
                    sh.cen[c] = cn;
                    const double t = (double)__fsub_rn(cn, sh.oldc[c]);
                    dist = __dadd_rn(dist, __dmul_rn(t, t));
                }
                max_shift = dist > max_shift ? dist : max_shift;
            }
            sh.flag = (iter + 1 == KM_MAX_ITER) || (max_shift <= KM_EPS2);
        }
        ++iter;
        __syncthreads();
        KC_TICK(7);   // centres
        if (sh.flag) break;
    }

    if (sh.fallback) {
        if (rank == 0 && tid == 0) results[frame * 3 + attempt].iters = KM_ITERS_FALLBACK;
        cluster.sync();
        return;
    }

    // ---- compactness: labels stay those assigned against oldc; distances to the final centres
    {
        KcFilter F = sh.filt;
        float oc[9], nc[9];
#pragma unroll
        for (int k = 0; k < 9; k++) { oc[k] = sh.oldc[k]; nc[k] = sh.cen[k]; }
        double *cdbl = (double *)seg;          // per-chunk sums (the k-means++ chunk sums are no longer needed)
        const int cnt_i = iter & 1;            // the counter the last Lloyd pass did not use
        for (;;) {
            const int lc = claim_chunk(&sh.next[cnt_i], lane);
            if (lc >= nch) break;
            const uint4 v = ((const uint4 *)pix)[lc * 32 + lane];
            const uint32_t p[4] = {v.x, v.y, v.z, v.w};
            const int rem = N - ((ch_lo + lc) * KC_CH + lane * 4);
            const unsigned valid = rem >= 4 ? 0xfu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
            bool k0[4], k1[4];
            classify4(p, valid, F, oc, k0, k1);
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if ((valid >> q) & 1u) {
                    const float3 x = unpack_px(p[q]);
                    const float cx = k0[q] ? nc[0] : (k1[q] ? nc[3] : nc[6]);
                    const float cy = k0[q] ? nc[1] : (k1[q] ? nc[4] : nc[7]);
                    const float cz = k0[q] ? nc[2] : (k1[q] ? nc[5] : nc[8]);
                    acc += (double)dist3(x.x, x.y, x.z, cx, cy, cz);
                }
            }
            acc = warp_sum_d(acc);
            if (lane == 0) cdbl[lc] = acc;
        }
        __syncthreads();
        if (swarp == 0) {                      // fixed order: the result does not depend on the chunk claiming
            double t = 0.0;
            for (int lc = lane; lc < nch; lc += 32) t += cdbl[lc];
            t = warp_sum_d(t);
            if (lane == 0) *cluster.map_shared_rank(&sh.xcomp[rank], 0) = t;
        }
        cluster.sync();
        if (rank == 0 && tid == 0) {
            double t = 0.0;
            for (int r = 0; r < C; r++) t += sh.xcomp[r];
            KmAttempt &res = results[frame * 3 + attempt];
            res.compactness = t;
            for (int k = 0; k < 9; k++) { res.centers[k] = sh.cen[k]; res.old_centers[k] = sh.oldc[k]; }
            res.n_fix = 0;
            for (int q = 0; q < 2; q++) { res.fix_idx[q] = 0; res.fix_k[q] = 0; }
            res.iters = iter;
        }
    }
    KC_TICK(8);   // compactness
#ifdef KC_TIMING
    if (tid == 0 && unit == (gridDim.x / C > 150 ? 150 : 1))
        printf("rank %d iters %d: load %lld  sync0 %lld  pp %lld  pass %lld  wait %lld  xchg %lld  tail %lld (search %lld sweep %lld wait %lld)  centres %lld  compact %lld  total %lld\n",
               rank, iter, tk[0], tk[1], tk[2], tk[3], tk[4], tk[5], tk[6], tk[9], tk[10], tk[11], tk[7], tk[8], clock64() - t_begin);
#endif
}

// -------------------------------------------------------------------------------------------------------------- launcher
// Cluster size: at most 141 chunks (72 KB of pixels) per CTA so that two 512-thread CTAs fit an SM, and no fewer CTAs than
// keep ~6000 pixels of work each. CKB_KM_CLUSTER / CKB_KM_THREADS (environment, tuning only) override the defaults.
static int kc_cluster_size(int N, int max_c)
{
    const int nchunk = (N + KC_CH - 1) / KC_CH;
    int C = 1;
    while (C < max_c && (nchunk + C - 1) / C > 141) C *= 2;
    while (C < max_c && (nchunk + C - 1) / C > 48) C *= 2;
    return C;
}

template <int NT>
static int kc_launch(ckb_ctx *ctx, const uint8_t *d_imgs, int n, const Region &rg, const uint64_t *d_rng_states,
                     KmAttempt *d_results, cudaStream_t st, int C)
{
    const int nchunk = (rg.N + KC_CH - 1) / KC_CH;
    const int cpc = (nchunk + C - 1) / C;
    const size_t dyn = (size_t)cpc * (KC_CH * 4 + 16 + 8 + 4 + 16);
    static bool attr_set[64] = {false};
    if (ctx->device >= 0 && ctx->device < 64 && !attr_set[ctx->device]) {
        CKB_CUDA(ctx, cudaFuncSetAttribute(ckb_kmeans_cluster_u8<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 141 * 556));
        CKB_CUDA(ctx, cudaFuncSetAttribute(ckb_kmeans_cluster_u8<NT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr_set[ctx->device] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(3 * n * C), 1, 1);
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
    CKB_CUDA(ctx, cudaLaunchKernelEx(&cfg, ckb_kmeans_cluster_u8<NT>, d_imgs, S, rg, cpc, d_rng_states, d_results));
    CKB_LAUNCH_CHECK(ctx, "ckb_kmeans_cluster");
    return CKB_OK;
}

int ckb_launch_kmeans_cluster(ckb_ctx *ctx, const uint8_t *d_imgs, int n, const Region &rg, const uint64_t *d_rng_states,
                              KmAttempt *d_results, cudaStream_t st)
{
    static int env_c = -1, env_nt = -1;
    if (env_c < 0) {
        const char *e = getenv("CKB_KM_CLUSTER");
        env_c = e ? atoi(e) : 0;
        e = getenv("CKB_KM_THREADS");
        env_nt = e ? atoi(e) : 0;
    }
    const int max_c = (env_c == 1 || env_c == 2 || env_c == 4 || env_c == 8 || env_c == 16) ? env_c : 8;
    const int C = kc_cluster_size(rg.N, max_c);
    if (((rg.N + KC_CH - 1) / KC_CH + C - 1) / C > 141)
        CKB_FAIL(ctx, CKB_E_INVALID, "ckb_find_stones: region too large for a cluster of %d", C);
    if (env_nt == 256) return kc_launch<256>(ctx, d_imgs, n, rg, d_rng_states, d_results, st, C);
    return kc_launch<512>(ctx, d_imgs, n, rg, d_rng_states, d_results, st, C);
}
