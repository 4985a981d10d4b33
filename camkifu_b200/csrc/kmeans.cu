// K3 + K2 — OpenCV-compatible 3-means colour clustering of the canonical goban image and the per-intersection
// classification that follows it.
//
// Replaces SfClustering.find_stones (src/camkifu/stone/sf_clustering.py:48-178):
//   cv2.kmeans(pixels, 3, None, (TERM_CRITERIA_EPS, 15, 3), 3, KMEANS_PP_CENTERS)                      :103-104
//   labels+1, *mask, 361 x np.unique -> ratios (uint8 percentages)                                     :105-129
//   interpret_ratios (centre grey -> B/E/W, argmax) and check_density                                  :131-178
//
// Parity design. cv2.kmeans' result depends on (a) the cv::RNG stream (multiply-with-carry), (b) k-means++ seeding:
// float32 squared distances, float64 running sums, a sequential "p -= dist[i]" sampling scan, best of 3 trials,
// (c) Lloyd iterations whose centre update is a SEQUENTIAL float32 sum in pixel order, (d) the stop rule
// max centre shift^2 <= 9 with labels NOT re-assigned on the last iteration, (e) best of 3 attempts by compactness.
// All of it is reproduced bit for bit. Distances and all the independent per-pixel work run in parallel; float64
// sums are taken in a fixed (deterministic) tree order — they are exact whenever OpenCV's are, which holds for pixel
// data (float32 terms spanning < 29 binades) — and the one inherently serial piece, the float32 centre sums, is
// executed serially where it has to be: the CTA computes the labels of a 512-pixel chunk in parallel, writes the
// per-(cluster, channel) masked values to shared memory, and nine lanes of warp 0 then walk the chunk in pixel order
// with one dependent FADD per pixel each (adding +0 for non-members is exact). The dependent-add latency (4 cycles)
// bounds such an iteration at ~0.3 ms for a whole board. For uint8 images (the canonical image itself, as opposed
// to SfClustering's float32 running average) the sums are integers that float32 represents exactly below 2^24, so
// the serial walk is only needed in the chunk where a running sum passes 2^24 (beyond it, up to 2^25, the rounding
// is a two-state parity automaton that is evaluated in parallel) — see the Lloyd loop.
// Throughput comes from running many (frame, attempt) CTAs side by side: grid = 3 attempts x n frames.
//
// Since round 2 the uint8 path is kmeans_cluster.cu (one thread-block cluster per frame, pixels resident in distributed
// shared memory); the kernel below serves float32 images (SfClustering's running average) and the attempts the cluster
// kernel declines (KM_ITERS_FALLBACK: an empty cluster, sums within reach of 2^25), for which it is launched after it and
// exits at once otherwise.
//
// Kernels:  ckb_pack_region     region pixels -> linear, vector-loadable scratch (uchar4 / float4 per pixel)
//           ckb_kmeans_attempt  one CTA per (frame, attempt)
//           ckb_zone_classify   12 CTAs per (frame, region): best attempt, labels of each zone's disk read from the image
//                               in place, zone histograms (warp per zone), ratios, stones; the last CTA of a unit
//                               (ticket) runs the density check
// Launchers: ckb_find_stones (one region) and ckb_find_stones_regions (SfMeta's 3 x 3 regions x n frames at once).
#include "kmeans_common.cuh"

#define KM_THREADS_F32 256             // float32 input: the serial centre sums dominate, many small CTAs per SM
#define KM_MAX_WARPS 32
#define KM_SEG 256                     // pixels per warp segment in the k-means++ passes
#define KM_MAX_N (380 * 380)
#define KM_MAX_SEG ((KM_MAX_N + KM_SEG - 1) / KM_SEG)   // 565
#define KM_CH 512                      // pixels per Lloyd chunk
#define KM_PLANE (KM_CH + 4)           // +4 words: the nine lanes' 128-bit reads fall in distinct banks
#define KM_MAX_CHUNK ((KM_MAX_N + KM_CH - 1) / KM_CH)   // 283

// ------------------------------------------------------------------------------------------------------------- pack
template <bool F32>
__global__ void __launch_bounds__(256) ckb_pack_region(const void *__restrict__ imgs, int S, Region rg,
                                                       void *__restrict__ scratch, size_t scratch_stride,
                                                       const KmAttempt *__restrict__ only_flagged, int nreg, int reg)
{
    const int f = blockIdx.y;
    // uint8 path: only the frames with an attempt the cluster kernel declined are packed (normally none)
    const int u3 = (f * nreg + reg) * 3;
    if (only_flagged && only_flagged[u3].iters != KM_ITERS_FALLBACK && only_flagged[u3 + 1].iters != KM_ITERS_FALLBACK &&
        only_flagged[u3 + 2].iters != KM_ITERS_FALLBACK)
        return;
    char *dst = (char *)scratch + (size_t)f * scratch_stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rg.N; i += gridDim.x * blockDim.x) {
        const int row = i / rg.w, col = i - row * rg.w;
        const size_t src = ((size_t)f * S * S + (size_t)(rg.x0 + row) * S + rg.y0 + col) * 3;
        if (F32) {
            const float *p = (const float *)imgs + src;
            ((float4 *)dst)[i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
        } else {
            const uint8_t *p = (const uint8_t *)imgs + src;
            ((uchar4 *)dst)[i] = make_uchar4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0);
        }
    }
}

// ---------------------------------------------------------------------------------------------------- k-means attempt
struct __align__(16) KmShared {
    union {
        double segsum[4][KM_MAX_SEG + 3];  // [0] current dist, [1..3] the three trial candidates   (k-means++)
        float planes[9][KM_PLANE];         // masked values per (cluster, channel) of one chunk     (Lloyd)
    } u;
    int csum[9][KM_MAX_CHUNK];             // uint8 input: exact integer member sums per chunk         (Lloyd)
    int ch0;                               // first chunk whose float32 running sums must be taken serially
    int seg_a[9][KM_CH / 32];              // uint8 input, sums in [2^24, 2^25): per 32-pixel segment sum of (x >> 1) ...
    int seg_f[9][KM_CH / 32];              // ... and the segment's rounding automaton (see the Lloyd loop)
    int mode_serial, mode_slow;
    double red_d[KM_MAX_WARPS * 3];
    int red_i[KM_MAX_WARPS * 4];
    float cen[9];
    float oldc[9];
    float sums[9];
    int cnt[3];
    int cand[3];
    int fix_idx[2];
    int fix_k[2];
    int n_fix;
    int flag;
    double dtmp[4];
};

// nearest already-chosen centre distance (k-means++ "dist" array, recomputed instead of stored)
__device__ __forceinline__ float pp_base(float3 x, const float *cen, int ncen)
{
    float b = dist3(x.x, x.y, x.z, cen[0], cen[1], cen[2]);
    if (ncen > 1) {
        const float d1 = dist3(x.x, x.y, x.z, cen[3], cen[4], cen[5]);
        b = d1 < b ? d1 : b;  // std::min(d, dist[i])
    }
    return b;
}

template <bool F32>
__device__ int pp_sample(const void *px, int N, int nseg, const double *segsum, double p, const float *cen, int ncen,
                         int lane)
{
    // first index ci in [0, N-1) with p - sum_{i<=ci} dist[i] <= 0, else N-1   (generateCentersPP)
    const int per = (nseg + 31) >> 5;
    const int s0 = lane * per, s1 = min(nseg, s0 + per);
    double loc = 0.0;
    for (int s = s0; s < s1; s++) loc += segsum[s];
    const double incl = warp_incl_scan_d(loc, lane);
    double excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 0.0;
    double r = p - excl;
    int found = -1;
    double r_before = 0.0;
    for (int s = s0; s < s1; s++) {
        const double rb = r;
        r -= segsum[s];
        if (r <= 0.0) { found = s; r_before = rb; break; }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, found >= 0);
    if (bal == 0u) return N - 1;
    const int src = __ffs(bal) - 1;
    const int seg = __shfl_sync(0xffffffffu, found, src);
    double resid = __shfl_sync(0xffffffffu, r_before, src);
    int ci = min(N - 1, seg * KM_SEG + KM_SEG - 1);
    for (int j = 0; j < KM_SEG / 32; j++) {
        const int i = seg * KM_SEG + j * 32 + lane;
        double d = 0.0;
        if (i < N) d = (double)pp_base(load_px<F32>(px, i), cen, ncen);
        const double pre = warp_incl_scan_d(d, lane);
        const unsigned hit = __ballot_sync(0xffffffffu, (i < N) && (resid - pre <= 0.0));
        if (hit) { ci = seg * KM_SEG + j * 32 + (__ffs(hit) - 1); break; }
        resid -= __shfl_sync(0xffffffffu, pre, 31);
    }
    return min(ci, N - 1);
}

// one k-means++ pass: for ncand candidate centres (pixel indices in sh.cand) compute min(d(x, cand), base) summed per
// segment into segsum[1 + c][seg]; ncen == 0 means "no base" (the very first centre).
template <bool F32, int NW>
__device__ void pp_pass(const void *px, int N, int nseg, KmShared &sh, int ncen, int ncand, int warp, int lane)
{
    float3 cd[3];
#pragma unroll
    for (int c = 0; c < 3; c++) cd[c] = load_px<F32>(px, sh.cand[c < ncand ? c : 0]);
    for (int seg = warp; seg < nseg; seg += NW) {
        double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < KM_SEG / 32; j++) {
            const int i = seg * KM_SEG + j * 32 + lane;
            if (i < N) {
                const float3 x = load_px<F32>(px, i);
                const float base = ncen > 0 ? pp_base(x, sh.cen, ncen) : FLT_MAX;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    if (c < ncand) {
                        const float d = dist3(x.x, x.y, x.z, cd[c].x, cd[c].y, cd[c].z);
                        acc[c] += (double)(ncen > 0 ? (d < base ? d : base) : d);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (c < ncand) {
                const double s = warp_sum_d(acc[c]);
                if (lane == 0) sh.u.segsum[1 + c][seg] = s;
            }
        }
    }
    __syncthreads();
    // totals: warp c sums the candidate-c segment sums
    if (warp < ncand) {
        double t = 0.0;
        for (int s = lane; s < nseg; s += 32) t += sh.u.segsum[1 + warp][s];
        t = warp_sum_d(t);
        if (lane == 0) sh.dtmp[1 + warp] = t;
    }
    __syncthreads();
}

// The same pass for uint8 pixels in integer arithmetic. Pixels and k-means++ centres (which are pixels) are packed
// bytes, d(x, c) = x.x + c.c - 2 x.c with three dp4a; every float32 / float64 quantity of the float version is an
// integer below 2^24 / 2^53 here, so the results are identical whatever the summation order.
template <int NW>
__device__ void pp_pass_u8(const void *px, int N, int nseg, KmShared &sh, int ncen, int ncand, int warp, int lane)
{
    const uint32_t *pw = (const uint32_t *)px;
    uint32_t cw[3], cc[3], bw[2] = {0u, 0u}, bc[2] = {0u, 0u};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        cw[c] = __ldg(pw + sh.cand[c < ncand ? c : 0]) & 0x00ffffffu;
        cc[c] = __dp4a(cw[c], cw[c], 0u);
    }
#pragma unroll
    for (int c = 0; c < 2; c++) {
        if (c < ncen) {
            bw[c] = (uint32_t)sh.cen[3 * c] | ((uint32_t)sh.cen[3 * c + 1] << 8) | ((uint32_t)sh.cen[3 * c + 2] << 16);
            bc[c] = __dp4a(bw[c], bw[c], 0u);
        }
    }
    for (int seg = warp; seg < nseg; seg += NW) {
        uint32_t acc[3] = {0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < KM_SEG / 32; j++) {
            const int i = seg * KM_SEG + j * 32 + lane;
            if (i < N) {
                const uint32_t x = __ldg(pw + i) & 0x00ffffffu;
                const uint32_t xx = __dp4a(x, x, 0u);
                uint32_t base = 0xffffffffu;
                if (ncen > 0) base = xx + bc[0] - 2u * __dp4a(x, bw[0], 0u);
                if (ncen > 1) base = min(base, xx + bc[1] - 2u * __dp4a(x, bw[1], 0u));
#pragma unroll
                for (int c = 0; c < 3; c++)
                    if (c < ncand) acc[c] += min(xx + cc[c] - 2u * __dp4a(x, cw[c], 0u), base);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (c < ncand) {
                const uint32_t t = __reduce_add_sync(0xffffffffu, acc[c]);   // <= 256 x 195075: fits
                if (lane == 0) sh.u.segsum[1 + c][seg] = (double)t;
            }
        }
    }
    __syncthreads();
    if (warp < ncand) {
        double t = 0.0;
        for (int s2 = lane; s2 < nseg; s2 += 32) t += sh.u.segsum[1 + warp][s2];
        t = warp_sum_d(t);
        if (lane == 0) sh.dtmp[1 + warp] = t;
    }
    __syncthreads();
}

template <bool F32, int NT>
__global__ void __launch_bounds__(NT) ckb_kmeans_attempt(const void *__restrict__ scratch,
                                                                 size_t scratch_stride, int N,
                                                                 const uint64_t *__restrict__ rng_states,
                                                                 KmAttempt *__restrict__ results, int only_flagged,
                                                                 int nreg, int reg)
{
    constexpr int NW = NT / 32;
    // batched multi-region calls: rng_states / results are indexed by unit = frame * nreg + region (scratch by frame)
    const int unit = blockIdx.y * nreg + reg;
    // uint8 path: this kernel is the fallback of ckb_kmeans_cluster_u8 and runs only the attempts that one declined
    if (only_flagged && results[unit * 3 + blockIdx.x].iters != KM_ITERS_FALLBACK) return;
    // uint8 input: the labels of the current iteration are kept (1 byte per pixel, behind the packed pixels in this
    // frame's scratch slot, one array per attempt) so that the compactness pass need not recompute the three distances
    // (7 N bytes of the 16 S^2-byte slot; the pixel part is only ever read, the label part only by this CTA)
    uint8_t *lab_cache = F32 ? nullptr
                             : (uint8_t *)const_cast<void *>(scratch) + (size_t)blockIdx.y * scratch_stride +
                                   (((size_t)N * 4 + 15) & ~(size_t)15) + (size_t)blockIdx.x * (((size_t)N + 15) & ~(size_t)15);
    __shared__ KmShared sh;
    const int attempt = blockIdx.x, frame = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const void *px = (const char *)scratch + (size_t)frame * scratch_stride;
    const int nseg = (N + KM_SEG - 1) / KM_SEG;

    // ---- cv::RNG draws of this attempt: 1 integer + 6 doubles = 13 draws
    uint64_t st = rng_states[unit];
    for (int k = 0; k < 13 * attempt; k++) rng_next(st);
    const int c0 = (int)(rng_next(st) % (uint32_t)N);
    double u[6];
#pragma unroll
    for (int k = 0; k < 6; k++) u[k] = rng_double(st);

#ifdef KM_TIMING
    long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t_prev = clock64();
#define KM_TICK(slot) do { const long long t_now = clock64(); tk[slot] += t_now - t_prev; t_prev = t_now; } while (0)
#else
#define KM_TICK(slot) do { } while (0)
#endif
    // ---- k-means++ seeding
    if (tid == 0) { sh.cand[0] = c0; sh.cand[1] = c0; sh.cand[2] = c0; }
    __syncthreads();
    if (F32) pp_pass<F32, NW>(px, N, nseg, sh, 0, 1, warp, lane);
    else pp_pass_u8<NW>(px, N, nseg, sh, 0, 1, warp, lane);
    if (tid < 3) {
        const float3 x = load_px<F32>(px, c0);
        sh.cen[tid] = tid == 0 ? x.x : (tid == 1 ? x.y : x.z);
    }
    for (int s = tid; s < nseg; s += NT) sh.u.segsum[0][s] = sh.u.segsum[1][s];
    if (tid == 0) sh.dtmp[0] = sh.dtmp[1];  // sum0
    __syncthreads();
    for (int k = 1; k < 3; k++) {
        if (warp < 3) {
            const double p = __dmul_rn(u[(k - 1) * 3 + warp], sh.dtmp[0]);
            const int ci = pp_sample<F32>(px, N, nseg, sh.u.segsum[0], p, sh.cen, k, lane);
            if (lane == 0) sh.cand[warp] = ci;
        }
        __syncthreads();
        if (F32) pp_pass<F32, NW>(px, N, nseg, sh, k, 3, warp, lane);
        else pp_pass_u8<NW>(px, N, nseg, sh, k, 3, warp, lane);
        // best trial: strict '<' in trial order
        int best = 0;
        double bs = sh.dtmp[1];
        if (sh.dtmp[2] < bs) { bs = sh.dtmp[2]; best = 1; }
        if (sh.dtmp[3] < bs) { bs = sh.dtmp[3]; best = 2; }
        __syncthreads();
        if (tid < 3) {
            const float3 x = load_px<F32>(px, sh.cand[best]);
            sh.cen[3 * k + tid] = tid == 0 ? x.x : (tid == 1 ? x.y : x.z);
        }
        for (int s = tid; s < nseg; s += NT) sh.u.segsum[0][s] = sh.u.segsum[1 + best][s];
        if (tid == 0) sh.dtmp[0] = bs;
        __syncthreads();
    }

    KM_TICK(0);   // k-means++
    // ---- Lloyd iterations
    const int nchunk = (N + KM_CH - 1) / KM_CH;
    int iter = 1;  // iteration 0 was the seeding; labels are (conceptually) assigned against sh.cen
    for (;;) {
        if (tid < 9) sh.oldc[tid] = sh.cen[tid];
        if (tid == 0) sh.n_fix = 0;
        __syncthreads();
        float oc[9];
#pragma unroll
        for (int k = 0; k < 9; k++) oc[k] = sh.oldc[k];

        // centre sums. OpenCV adds the members of a cluster in pixel order in float32. For uint8-valued pixels every
        // partial sum below 2^24 is an integer that float32 holds exactly, so up to that point the order does not
        // matter: pass 1 labels all pixels and takes exact integer sums per 512-pixel chunk fully in parallel (one
        // warp per chunk, no block barrier), a nine-lane prefix over the chunk sums finds the first chunk in which
        // any (cluster, channel) running sum could pass 2^24, and only from that chunk on (`ch0`; typically the last
        // few chunks of the largest cluster's brightest channel, often none) are the adds executed serially as for
        // float32 input: the CTA labels a chunk in parallel, writes the per-(cluster, channel) masked values to shared
        // memory, and nine lanes of warp 0 walk it in pixel order with one dependent FADD per pixel each.
        float acc = 0.f;
        int c0n = 0, c1n = 0, c2n = 0;
        int ch0 = 0;
        if (!F32) {
            for (int ch = warp; ch < nchunk; ch += NW) {
                int s9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
                for (int j = 0; j < KM_CH / 32; j++) {
                    const int i = ch * KM_CH + j * 32 + lane;
                    if (i < N) {
                        const float3 x = load_px<F32>(px, i);
                        const int lab = argmin3(x, oc);
                        lab_cache[i] = (uint8_t)lab;
                        const int vx = (int)x.x, vy = (int)x.y, vz = (int)x.z;
                        c0n += lab == 0;
                        c1n += lab == 1;
                        c2n += lab == 2;
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const int m = lab == k;
                            s9[3 * k + 0] += m * vx;
                            s9[3 * k + 1] += m * vy;
                            s9[3 * k + 2] += m * vz;
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 9; c++) {
                    const int t = __reduce_add_sync(0xffffffffu, s9[c]);
                    if (lane == 0) sh.csum[c][ch] = t;
                }
            }
            __syncthreads();
            KM_TICK(1);   // parallel pass
            if (warp == 0) {
                // first chunk in which any of the nine running sums passes 2^24, and the exact sums before it: every
                // lane owns a run of consecutive chunks, warp scan over the lane totals, one chain after the other
                const int per = (nchunk + 31) >> 5, c_lo = lane * per, c_hi = min(nchunk, c_lo + per);
                int first = nchunk;
#pragma unroll 1
                for (int c = 0; c < 9; c++) {
                    int loc = 0;
                    for (int ch = c_lo; ch < c_hi; ch++) loc += sh.csum[c][ch];
                    int incl = loc;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    if (incl > (1 << 24)) {          // a partial sum inside one of this lane's chunks may need rounding
                        int run = incl - loc;
                        for (int ch = c_lo; ch < c_hi; ch++) {
                            run += sh.csum[c][ch];
                            if (run > (1 << 24)) { first = min(first, ch); break; }
                        }
                    }
                }
                first = __reduce_min_sync(0xffffffffu, first);
#pragma unroll 1
                for (int c = 0; c < 9; c++) {
                    int loc = 0;
                    for (int ch = c_lo; ch < min(c_hi, first); ch++) loc += sh.csum[c][ch];
                    loc = __reduce_add_sync(0xffffffffu, loc);
                    if (lane == c) sh.sums[c] = (float)loc;           // exact: loc <= 2^24
                }
                if (lane == 0) sh.ch0 = first;
            }
            __syncthreads();
            ch0 = sh.ch0;
            if (warp == 0 && lane < 9) acc = sh.sums[lane];
            KM_TICK(2);   // prefix over chunk sums
        }
        constexpr int Q = (KM_CH + NT - 1) / NT;     // chunk pixels per thread in the serial phase
        float3 xn[Q];
        bool vn[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const int p = q * NT + tid;
            const int i = ch0 * KM_CH + p;
            vn[q] = p < KM_CH && i < N;
            xn[q] = vn[q] ? load_px<F32>(px, i) : make_float3(0.f, 0.f, 0.f);
        }
        for (int ch = ch0; ch < nchunk; ch++) {
            int labq[Q];
            float3 xc[Q];
#pragma unroll
            for (int q = 0; q < Q; q++) {
                const int p = q * NT + tid;
                xc[q] = xn[q];
                labq[q] = (p < KM_CH && vn[q]) ? argmin3(xc[q], oc) : -1;
                const int i = (ch + 1) * KM_CH + p;      // fetch the next chunk's pixels while this one is summed
                vn[q] = p < KM_CH && i < N;
                xn[q] = vn[q] ? load_px<F32>(px, i) : make_float3(0.f, 0.f, 0.f);
            }
            // uint8 input: decide how this chunk's sums are taken. A running sum that stays <= 2^24 is exact (add the
            // chunk's integer sum); one in [2^24, 2^25) is an even integer 2u and float32 addition of an integer x
            // acts on u as  u += (x >> 1) + (x odd ? (u + (x >> 1)) & 1 : 0)  (ties go to the even significand), a
            // two-state automaton on the parity of u that is evaluated in parallel below; only the chunk in which a
            // sum crosses 2^24, or one that could reach 2^25, is walked serially.
            bool serial = true;
            unsigned slow = 0u;
            if (!F32) {
                if (warp == 0) {
                    int need = 0, sl = 0;
                    if (lane < 9) {
                        const int cur = (int)acc;                       // exact: acc is an integer below 2^25
                        if (cur < (1 << 24)) need = cur + sh.csum[lane][ch] > (1 << 24);
                        else { sl = 1; need = cur + KM_CH * 255 >= (1 << 25); }
                    }
                    const unsigned nb = __ballot_sync(0xffffffffu, need), sb = __ballot_sync(0xffffffffu, sl);
                    if (lane == 0) { sh.mode_serial = nb != 0u; sh.mode_slow = (int)(sb & 0x1ffu); }
                }
                __syncthreads();
                serial = sh.mode_serial != 0;
                slow = (unsigned)sh.mode_slow;
            }
            if (serial) {
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    const int p = q * NT + tid;
                    if (p >= KM_CH) continue;
                    const float3 x = xc[q];
                    const int lab = labq[q];
                    if (F32) {          // uint8 input counted its members in pass 1
                        c0n += lab == 0;
                        c1n += lab == 1;
                        c2n += lab == 2;
                    }
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const bool m = lab == k;
                        sh.u.planes[3 * k + 0][p] = m ? x.x : 0.f;
                        sh.u.planes[3 * k + 1][p] = m ? x.y : 0.f;
                        sh.u.planes[3 * k + 2][p] = m ? x.z : 0.f;
                    }
                }
                __syncthreads();
                if (warp == 0 && lane < 9) {
                    const float4 *pl = (const float4 *)sh.u.planes[lane];
#pragma unroll 8
                    for (int p = 0; p < KM_CH / 4; p++) {
                        const float4 v = pl[p];
                        acc = __fadd_rn(acc, v.x);
                        acc = __fadd_rn(acc, v.y);
                        acc = __fadd_rn(acc, v.z);
                        acc = __fadd_rn(acc, v.w);
                    }
                }
                __syncthreads();
            } else {
                // per 32-pixel segment (= one warp's pixels, in pixel order) and slow chain: A = sum of x >> 1 over the
                // members, and the segment's automaton: has_odd, the parity the segment leaves behind (c), the term
                // q of its first odd member (whose increment is p_in ^ q), and the increments D of its other odd
                // members, each of which sees parity 0 after the previous odd member plus the even members in between
                for (unsigned rest = slow; rest; rest &= rest - 1) {
                    const int c = __ffs(rest) - 1, k = c / 3, j = c - 3 * k;
#pragma unroll
                    for (int q = 0; q < Q; q++) {
                        const int p = q * NT + tid;
                        if (p >= KM_CH) continue;               // warp-uniform: NT and KM_CH are multiples of 32
                        const bool mem = labq[q] == k;
                        const int xi = (int)(j == 0 ? xc[q].x : (j == 1 ? xc[q].y : xc[q].z));
                        const int a = mem ? (xi >> 1) : 0;
                        const bool odd = mem && (xi & 1), cb = mem && !(xi & 1) && (a & 1);
                        const unsigned mo = __ballot_sync(0xffffffffu, odd), mc = __ballot_sync(0xffffffffu, cb);
                        const unsigned below = (1u << lane) - 1u, ob = mo & below;
                        int dl = 0, qf = 0;
                        if (odd) {
                            if (ob) {
                                const int last = 31 - __clz(ob);
                                dl = (__popc(mc & below & ~((2u << last) - 1u)) ^ a) & 1;
                            } else {
                                qf = (__popc(mc & below) ^ a) & 1;
                            }
                        }
                        const int A = __reduce_add_sync(0xffffffffu, a);
                        const int D = __popc(__ballot_sync(0xffffffffu, dl != 0));
                        const int qfirst = __ballot_sync(0xffffffffu, qf != 0) != 0u;
                        int cw;
                        if (mo) {
                            const int top = 31 - __clz(mo);
                            cw = top == 31 ? 0 : (__popc(mc >> (top + 1)) & 1);
                        } else {
                            cw = __popc(mc) & 1;
                        }
                        if (lane == 0) {
                            sh.seg_a[c][p >> 5] = A;
                            sh.seg_f[c][p >> 5] = (mo != 0u) | (cw << 1) | (qfirst << 2) | (D << 3);
                        }
                    }
                }
                __syncthreads();
                if (warp == 0 && lane < 9) {
                    const int cur = (int)acc;
                    if (!((slow >> lane) & 1u)) {
                        acc = (float)(cur + sh.csum[lane][ch]);        // exact regime
                    } else {
                        int u = cur >> 1, par = u & 1, add = 0;
                        for (int sg = 0; sg < KM_CH / 32; sg++) {
                            const int f = sh.seg_f[lane][sg];
                            add += sh.seg_a[lane][sg];
                            if (f & 1) {
                                add += (f >> 3) + (par ^ ((f >> 2) & 1));
                                par = (f >> 1) & 1;
                            } else {
                                par ^= (f >> 1) & 1;
                            }
                        }
                        acc = (float)((u + add) << 1);                 // even and below 2^25: exact
                    }
                }
                __syncthreads();
            }
        }
        KM_TICK(3);   // serial chunks
        if (warp == 0 && lane < 9) sh.sums[lane] = acc;
        // counts
        c0n = __reduce_add_sync(0xffffffffu, c0n);
        c1n = __reduce_add_sync(0xffffffffu, c1n);
        c2n = __reduce_add_sync(0xffffffffu, c2n);
        if (lane == 0) { sh.red_i[warp * 4 + 0] = c0n; sh.red_i[warp * 4 + 1] = c1n; sh.red_i[warp * 4 + 2] = c2n; }
        __syncthreads();
        if (tid < 3) {
            int t = 0;
            for (int w = 0; w < NW; w++) t += sh.red_i[w * 4 + tid];
            sh.cnt[tid] = t;
        }
        __syncthreads();

        // empty-cluster repair (rare): move the farthest point of the biggest cluster into the empty one
        for (int k = 0; k < 3; k++) {
            if (sh.cnt[k] != 0) continue;  // block-uniform
            int max_k = 0;
            for (int k1 = 1; k1 < 3; k1++) if (sh.cnt[max_k] < sh.cnt[k1]) max_k = k1;
            const float scale = __fdiv_rn(1.f, (float)sh.cnt[max_k]);
            const float bx = __fmul_rn(sh.sums[3 * max_k], scale), by = __fmul_rn(sh.sums[3 * max_k + 1], scale),
                        bz = __fmul_rn(sh.sums[3 * max_k + 2], scale);
            // farthest: `if (max_dist <= dist)` in index order => largest distance, last index among ties
            float bestd = -1.f;
            int besti = -1;
            for (int i = tid; i < N; i += NT) {
                const float3 x = load_px<F32>(px, i);
                int lab = argmin3(x, oc);
                for (int q = 0; q < sh.n_fix; q++) if (sh.fix_idx[q] == i) lab = sh.fix_k[q];
                if (lab != max_k) continue;
                const float d = dist3(x.x, x.y, x.z, bx, by, bz);
                if (d > bestd || (d == bestd && i > besti)) { bestd = d; besti = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bestd, o);
                const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                if (od > bestd || (od == bestd && oi > besti)) { bestd = od; besti = oi; }
            }
            if (lane == 0) { sh.red_d[warp] = (double)bestd; sh.red_i[warp] = besti; }
            __syncthreads();
            if (tid == 0) {
                double bd = sh.red_d[0];
                int bi = sh.red_i[0];
                for (int w = 1; w < NW; w++)
                    if (sh.red_d[w] > bd || (sh.red_d[w] == bd && sh.red_i[w] > bi)) { bd = sh.red_d[w]; bi = sh.red_i[w]; }
                const float3 x = load_px<F32>(px, bi);
                sh.cnt[max_k]--;
                sh.cnt[k]++;
                sh.fix_idx[sh.n_fix] = bi;
                sh.fix_k[sh.n_fix] = k;
                sh.n_fix++;
                sh.sums[3 * max_k + 0] = __fsub_rn(sh.sums[3 * max_k + 0], x.x);
                sh.sums[3 * max_k + 1] = __fsub_rn(sh.sums[3 * max_k + 1], x.y);
                sh.sums[3 * max_k + 2] = __fsub_rn(sh.sums[3 * max_k + 2], x.z);
                sh.sums[3 * k + 0] = __fadd_rn(sh.sums[3 * k + 0], x.x);
                sh.sums[3 * k + 1] = __fadd_rn(sh.sums[3 * k + 1], x.y);
                sh.sums[3 * k + 2] = __fadd_rn(sh.sums[3 * k + 2], x.z);
            }
            __syncthreads();
        }

        // new centres, shift, stop rule
        if (tid == 0) {
            double max_shift = 0.0;
            for (int k = 0; k < 3; k++) {
                const float scale = __fdiv_rn(1.f, (float)sh.cnt[k]);
                double dist = 0.0;
                for (int j = 0; j < 3; j++) {
                    const float c = __fmul_rn(sh.sums[3 * k + j], scale);
                    sh.cen[3 * k + j] = c;
                    const double t = (double)__fsub_rn(c, sh.oldc[3 * k + j]);
                    dist = __dadd_rn(dist, __dmul_rn(t, t));
                }
                max_shift = dist > max_shift ? dist : max_shift;
            }
            ++iter;
            sh.flag = (iter == KM_MAX_ITER) || (max_shift <= KM_EPS2);
        } else {
            ++iter;
        }
        __syncthreads();
        if (sh.flag) break;
    }

    KM_TICK(4);   // iteration tails
    // ---- compactness: labels stay those assigned against oldc (+ repairs); distances to the final centres
    {
        float oc[9], nc[9];
#pragma unroll
        for (int k = 0; k < 9; k++) { oc[k] = sh.oldc[k]; nc[k] = sh.cen[k]; }
        double acc = 0.0;
        const int n_fix = sh.n_fix;
#pragma unroll 4
        for (int i = tid; i < N; i += NT) {
            const float3 x = load_px<F32>(px, i);
            int lab = F32 ? argmin3(x, oc) : (int)lab_cache[i];
            for (int q = 0; q < n_fix; q++) if (sh.fix_idx[q] == i) lab = sh.fix_k[q];
            const float cx = lab == 0 ? nc[0] : (lab == 1 ? nc[3] : nc[6]);
            const float cy = lab == 0 ? nc[1] : (lab == 1 ? nc[4] : nc[7]);
            const float cz = lab == 0 ? nc[2] : (lab == 1 ? nc[5] : nc[8]);
            acc += (double)dist3(x.x, x.y, x.z, cx, cy, cz);
        }
        acc = warp_sum_d(acc);
        if (lane == 0) sh.red_d[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < NW; w++) t += sh.red_d[w];
            KmAttempt &r = results[unit * 3 + attempt];
            r.compactness = t;
            for (int k = 0; k < 9; k++) { r.centers[k] = sh.cen[k]; r.old_centers[k] = sh.oldc[k]; }
            r.n_fix = sh.n_fix;
            for (int q = 0; q < 2; q++) { r.fix_idx[q] = sh.fix_idx[q]; r.fix_k[q] = sh.fix_k[q]; }
            r.iters = iter;
        }
    }
    KM_TICK(5);   // compactness
#ifdef KM_TIMING
    if (tid == 0 && frame == 0)
        printf("attempt %d iters %d: pp %lld  pass1 %lld  prefix %lld  serial %lld  tails %lld  compact %lld cycles\n", attempt,
               iter, tk[0], tk[1], tk[2], tk[3], tk[4], tk[5]);
#endif
}

// ------------------------------------------------------------------------------------------------ zone classification
__device__ __forceinline__ int center_grey(const float *c)
{
    // int(sum(x) / 3) over numpy float32 scalars (sf_clustering.py:107,148): float32 adds, float32 divide, truncate
    float s = __fadd_rn(0.f, c[0]);
    s = __fadd_rn(s, c[1]);
    s = __fadd_rn(s, c[2]);
    return (int)__fdiv_rn(s, 3.0f);
}

#define ZC_THREADS 256
#define ZC_SPLIT 12                    // CTAs per frame: 12 x 8 warps over the (at most) 361 zones

// pixel (row i, column j of the canonical image) as float32 BGR. uint8 images are read in place; float32 images from
// the packed scratch of ckb_pack_region (region-linear float4).
template <bool F32>
__device__ __forceinline__ float3 zc_pixel(const void *base, int S, const Region &rg, int i, int j)
{
    if (F32) {
        return load_px<true>(base, (i - rg.x0) * rg.w + (j - rg.y0));
    } else {
        const uint8_t *s = (const uint8_t *)base + ((size_t)i * S + j) * 3;
        return make_float3((float)__ldg(s), (float)__ldg(s + 1), (float)__ldg(s + 2));
    }
}

// grid (ZC_SPLIT, n): each CTA votes a share of the zones (one warp per zone: labels of the zone's disk against the
// best attempt's centres, ballot-free integer histogram, ratios, colour); the last CTA of a frame to finish (ticket in
// `tickets`, which the launcher zeroes) runs check_density over the frame's 361 stones.
template <bool F32>
__global__ void __launch_bounds__(ZC_THREADS) ckb_zone_classify(const void *__restrict__ pixels, size_t frame_stride,
                                                                const __grid_constant__ RegionSet regs, int gsize,
                                                                const KmAttempt *__restrict__ results,
                                                                const int32_t *__restrict__ rects,
                                                                const uint8_t *__restrict__ mask, int S,
                                                                uint8_t *__restrict__ stones_ws, unsigned *__restrict__ tickets,
                                                                uint8_t *__restrict__ stones_out,
                                                                uint8_t *__restrict__ trusted_out,
                                                                uint8_t *__restrict__ ratios_out,
                                                                float *__restrict__ centers_out,
                                                                double *__restrict__ compact_out,
                                                                int32_t *__restrict__ labels_out)
{
    __shared__ int s_hist[3];
    __shared__ int s_last;
    // one unit per (frame, region); every output is indexed by unit (= frame for the single-region calls)
    const int unit = blockIdx.y, fr = unit / regs.n;
    const Region rg = regs.r[unit - fr * regs.n];
    const int rs = rg.rs, re = rg.re, cs = rg.cs, ce = rg.ce;
    const int frame = unit;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const void *px = (const char *)pixels + (size_t)fr * frame_stride;

    // best attempt: `if (compactness < best_compactness)` in attempt order, starting from DBL_MAX
    int best = 0;
    double bc = DBL_MAX;
    for (int a = 0; a < 3; a++) {
        const double c = results[frame * 3 + a].compactness;
        if (c < bc) { bc = c; best = a; }
    }
    const KmAttempt &R = results[frame * 3 + best];
    float oc[9];
#pragma unroll
    for (int k = 0; k < 9; k++) oc[k] = R.old_centers[k];
    const int n_fix = R.n_fix;
    const int f0 = R.fix_idx[0], f1 = R.fix_idx[1], fk0 = R.fix_k[0], fk1 = R.fix_k[1];

    int cv[3];
    for (int k = 0; k < 3; k++) cv[k] = center_grey(R.centers + 3 * k);
    const int mn = min(cv[0], min(cv[1], cv[2])), mx = max(cv[0], max(cv[1], cv[2]));
    const int med = cv[0] + cv[1] + cv[2] - mn - mx;
    const int mid = cv[0] == med ? 0 : (cv[1] == med ? 1 : 2);  // centers_val.index(sorted(centers_val)[1])
    uint8_t col[3];
    for (int k = 0; k < 3; k++) col[k] = cv[k] == mn ? CKB_B : (cv[k] == mx ? CKB_W : CKB_E);

    if (blockIdx.x == 0 && tid == 0) {
        if (centers_out) for (int k = 0; k < 9; k++) centers_out[frame * 9 + k] = R.centers[k];
        if (compact_out) compact_out[frame] = bc;
    }
    const int nz = gsize * gsize;
    for (int z = blockIdx.x * (ZC_THREADS / 32) + warp; z < nz; z += ZC_SPLIT * (ZC_THREADS / 32)) {
        const int zr = z / gsize, zc = z - zr * gsize;
        uint8_t r3[3] = {0, 0, 0};
        r3[mid] = 1;
        uint8_t stone = CKB_E;
        if (zr >= rs && zr < re && zc >= cs && zc < ce) {
            const int a0 = rects[z * 4 + 0], b0 = rects[z * 4 + 1], a1 = rects[z * 4 + 2], b1 = rects[z * 4 + 3];
            const int zw = b1 - b0, total = (a1 - a0) * zw;
            int c0n = 0, c1n = 0, c2n = 0;
            for (int t = lane; t < total; t += 32) {
                const int i = a0 + t / zw, j = b0 + t % zw;
                if (mask[(size_t)i * S + j]) {
                    const int li = (i - rg.x0) * rg.w + (j - rg.y0);
                    int lab = argmin3(zc_pixel<F32>(px, S, rg, i, j), oc);
                    if (n_fix > 0 && li == f0) lab = fk0;
                    if (n_fix > 1 && li == f1) lab = fk1;
                    c0n += lab == 0;
                    c1n += lab == 1;
                    c2n += lab == 2;
                }
            }
            c0n = __reduce_add_sync(0xffffffffu, c0n);
            c1n = __reduce_add_sync(0xffffffffu, c1n);
            c2n = __reduce_add_sync(0xffffffffu, c2n);
            // uint8(100 * count / zone_pixels), written only for labels that are present
            if (c0n) r3[0] = (uint8_t)((100 * c0n) / total);
            if (c1n) r3[1] = (uint8_t)((100 * c1n) / total);
            if (c2n) r3[2] = (uint8_t)((100 * c2n) / total);
            int b = 0;
            if (r3[1] > r3[b]) b = 1;
            if (r3[2] > r3[b]) b = 2;
            stone = col[b];
        }
        if (lane == 0) {
            stones_ws[(size_t)frame * CKB_MAX_ZONES + z] = stone;
            if (stones_out) stones_out[(size_t)frame * nz + z] = stone;
            if (ratios_out) {
                uint8_t *o = ratios_out + ((size_t)frame * nz + z) * 3;
                o[0] = r3[0]; o[1] = r3[1]; o[2] = r3[2];
            }
        }
    }
    if (labels_out) {
        int32_t *lo = labels_out + (size_t)frame * rg.N;
        for (int i = blockIdx.x * ZC_THREADS + tid; i < rg.N; i += ZC_SPLIT * ZC_THREADS) {
            const int row = i / rg.w, col2 = i - row * rg.w;
            int lab = argmin3(zc_pixel<F32>(px, S, rg, rg.x0 + row, rg.y0 + col2), oc);
            if (n_fix > 0 && i == f0) lab = fk0;
            if (n_fix > 1 && i == f1) lab = fk1;
            lo[i] = lab;
        }
    }
    // check_density: three distinct values, each seen at least twice (sf_clustering.py:170-178) — by the last CTA
    if (!trusted_out) return;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        s_last = atomicAdd(&tickets[frame], 1u) == ZC_SPLIT - 1;
        s_hist[0] = s_hist[1] = s_hist[2] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int z = tid; z < nz; z += ZC_THREADS) atomicAdd(&s_hist[__ldcg(stones_ws + (size_t)frame * CKB_MAX_ZONES + z)], 1);
    __syncthreads();
    if (tid == 0) trusted_out[frame] = (s_hist[0] >= 2 && s_hist[1] >= 2 && s_hist[2] >= 2) ? 1 : 0;
}

// ----------------------------------------------------------------------------------------------------------- launcher
int ckb_kmeans_init_tables(ckb_ctx *ctx)
{
    (void)ctx;
    return CKB_OK;
}

static Region make_region(const ckb_ctx *ctx, int rs, int re, int cs, int ce)
{
    // cluster_colors: x0, y0 = getrect(rs, cs)[:2]; x1, y1 = getrect(re-1, ce-1)[2:]   (sf_clustering.py:99-101)
    const int g = ctx->gsize;
    Region r;
    r.x0 = ctx->h_rects[(rs * g + cs) * 4 + 0];
    r.y0 = ctx->h_rects[(rs * g + cs) * 4 + 1];
    r.h = ctx->h_rects[((re - 1) * g + (ce - 1)) * 4 + 2] - r.x0;
    r.w = ctx->h_rects[((re - 1) * g + (ce - 1)) * 4 + 3] - r.y0;
    r.N = r.h * r.w;
    r.rs = rs; r.re = re; r.cs = cs; r.ce = ce;
    return r;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int ckb_launch_kmeans_cluster(ckb_ctx *ctx, const uint8_t *d_imgs, int n, const RegionSet &regs, const uint64_t *d_rng_states,
                              KmAttempt *d_results, cudaStream_t st);

// workspace: [n] per-frame scratch (packed pixels + label caches of the one-CTA kernel) | [3 units] KmAttempt |
// [units][361] stones of the zone vote | [units] tickets          (units = n frames x n_regions)
static size_t find_stones_workspace(const ckb_ctx *ctx, int n, int nreg)
{
    const size_t per_frame = align_up((size_t)ctx->S * ctx->S * 16, 256);
    const size_t units = (size_t)n * nreg;
    return (size_t)n * per_frame + align_up(units * 3 * sizeof(KmAttempt), 256) + align_up(units * CKB_MAX_ZONES, 256) +
           align_up(units * sizeof(unsigned), 256) + 256;
}

extern "C" size_t ckb_find_stones_workspace(const ckb_ctx *ctx, int n)
{
    if (!ctx || n < 0) return 0;
    return find_stones_workspace(ctx, n, 1);
}

extern "C" size_t ckb_find_stones_regions_workspace(const ckb_ctx *ctx, int n, int n_regions)
{
    if (!ctx || n < 0 || n_regions < 1 || n_regions > CKB_MAX_REGIONS) return 0;
    return find_stones_workspace(ctx, n, n_regions);
}

static int find_stones_impl(ckb_ctx *ctx, const void *d_imgs, int is_f32, int n, const RegionSet &regs,
                            const uint64_t *d_rng_states, void *d_work, size_t work_bytes, uint8_t *d_stones,
                            uint8_t *d_trusted, uint8_t *d_ratios, float *d_centers, double *d_compactness,
                            int32_t *d_labels, void *stream)
{
    const int g = ctx->gsize, nreg = regs.n;
    if (work_bytes < find_stones_workspace(ctx, n, nreg)) CKB_FAIL(ctx, CKB_E_NOMEM, "ckb_find_stones: workspace too small");
    if (((uintptr_t)d_work & 255) != 0) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_find_stones: workspace must be 256-byte aligned");
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    CKB_ENTER(ctx, stream);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t stride = align_up((size_t)ctx->S * ctx->S * 16, 256);
    const size_t units = (size_t)n * nreg;
    char *w = (char *)d_work + (size_t)n * stride;
    KmAttempt *res = (KmAttempt *)w;
    w += align_up(units * 3 * sizeof(KmAttempt), 256);
    uint8_t *stones_ws = (uint8_t *)w;
    w += align_up(units * CKB_MAX_ZONES, 256);
    unsigned *tickets = (unsigned *)w;
    CKB_CUDA(ctx, cudaMemsetAsync(tickets, 0, units * sizeof(unsigned), st));
    if (is_f32) {
        const Region &rg = regs.r[0];          // float32 images: single-region calls only
        dim3 pgrid((rg.N + 255) / 256, n);
        ckb_pack_region<true><<<pgrid, 256, 0, st>>>(d_imgs, ctx->S, rg, d_work, stride, nullptr, 1, 0);
        CKB_LAUNCH_CHECK(ctx, "ckb_pack_region");
        ckb_kmeans_attempt<true, KM_THREADS_F32><<<dim3(3, n), KM_THREADS_F32, 0, st>>>(d_work, stride, rg.N, d_rng_states, res, 0, 1, 0);
        CKB_LAUNCH_CHECK(ctx, "ckb_kmeans_attempt");
        ckb_zone_classify<true><<<dim3(ZC_SPLIT, n), ZC_THREADS, 0, st>>>(d_work, stride, regs, g, res, ctx->d_rects, ctx->d_mask,
                                                                         ctx->S, stones_ws, tickets, d_stones, d_trusted,
                                                                         d_ratios, d_centers, d_compactness, d_labels);
        CKB_LAUNCH_CHECK(ctx, "ckb_zone_classify");
    } else {
        // uint8 images: one thread-block cluster per (frame, region), pixels resident in distributed shared memory, the
        // three attempts in lock step (kmeans_cluster.cu). The attempts it declines (an empty cluster; sums within reach
        // of 2^25: pathological images) are marked in `res` and re-run by the one-CTA kernel, which otherwise exits at once.
        const int rc = ckb_launch_kmeans_cluster(ctx, (const uint8_t *)d_imgs, n, regs, d_rng_states, res, st);
        if (rc != CKB_OK) return rc;
        for (int r = 0; r < nreg; r++) {
            ckb_pack_region<false><<<dim3(8, n), 256, 0, st>>>(d_imgs, ctx->S, regs.r[r], d_work, stride, res, nreg, r);
            CKB_LAUNCH_CHECK(ctx, "ckb_pack_region");
            ckb_kmeans_attempt<false, 1024><<<dim3(3, n), 1024, 0, st>>>(d_work, stride, regs.r[r].N, d_rng_states, res, 1, nreg, r);
            CKB_LAUNCH_CHECK(ctx, "ckb_kmeans_attempt");
        }
        ckb_zone_classify<false><<<dim3(ZC_SPLIT, (unsigned)units), ZC_THREADS, 0, st>>>(
            d_imgs, (size_t)ctx->S * ctx->S * 3, regs, g, res, ctx->d_rects, ctx->d_mask, ctx->S, stones_ws, tickets, d_stones,
            d_trusted, d_ratios, d_centers, d_compactness, d_labels);
        CKB_LAUNCH_CHECK(ctx, "ckb_zone_classify");
    }
    return CKB_OK;
}

extern "C" int ckb_find_stones(ckb_ctx *ctx, const void *d_imgs, int is_f32, int n, int rs, int re, int cs, int ce,
                               const uint64_t *d_rng_states, void *d_work, size_t work_bytes, uint8_t *d_stones,
                               uint8_t *d_trusted, uint8_t *d_ratios, float *d_centers, double *d_compactness,
                               int32_t *d_labels, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    const int g = ctx->gsize;
    if (n == 0) return CKB_OK;   // an empty batch is a no-op, whatever the pointers
    if (!d_imgs || !d_rng_states || !d_work || n < 0) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_find_stones: bad argument");
    if (rs < 0 || cs < 0 || re > g || ce > g || rs >= re || cs >= ce)
        CKB_FAIL(ctx, CKB_E_INVALID, "ckb_find_stones: region [%d,%d)x[%d,%d) outside the %dx%d goban", rs, re, cs, ce, g, g);
    RegionSet regs;
    memset(&regs, 0, sizeof regs);
    regs.n = 1;
    regs.r[0] = make_region(ctx, rs, re, cs, ce);
    return find_stones_impl(ctx, d_imgs, is_f32, n, regs, d_rng_states, d_work, work_bytes, d_stones, d_trusted, d_ratios,
                            d_centers, d_compactness, d_labels, stream);
}

extern "C" int ckb_find_stones_regions(ckb_ctx *ctx, const uint8_t *d_imgs, int n, int n_regions, const int *h_regions4,
                                       const uint64_t *d_rng_states, void *d_work, size_t work_bytes, uint8_t *d_stones,
                                       uint8_t *d_trusted, uint8_t *d_ratios, float *d_centers, double *d_compactness,
                                       void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    const int g = ctx->gsize;
    if (n == 0) return CKB_OK;
    if (!d_imgs || !d_rng_states || !d_work || !h_regions4 || n < 0 || n_regions < 1 || n_regions > CKB_MAX_REGIONS)
        CKB_FAIL(ctx, CKB_E_INVALID, "ckb_find_stones_regions: bad argument (1 .. %d regions)", CKB_MAX_REGIONS);
    RegionSet regs;
    memset(&regs, 0, sizeof regs);
    regs.n = n_regions;
    for (int r = 0; r < n_regions; r++) {
        const int rs = h_regions4[4 * r], re = h_regions4[4 * r + 1], cs = h_regions4[4 * r + 2], ce = h_regions4[4 * r + 3];
        if (rs < 0 || cs < 0 || re > g || ce > g || rs >= re || cs >= ce)
            CKB_FAIL(ctx, CKB_E_INVALID, "ckb_find_stones_regions: region %d [%d,%d)x[%d,%d) outside the %dx%d goban", r, rs,
                     re, cs, ce, g, g);
        regs.r[r] = make_region(ctx, rs, re, cs, ce);
    }
    return find_stones_impl(ctx, d_imgs, 0, n, regs, d_rng_states, d_work, work_bytes, d_stones, d_trusted, d_ratios, d_centers,
                            d_compactness, nullptr, stream);
}
