// K4 front end: patch gather + conv1 + ReLU + conv2 + ReLU + 2x2 max-pool of the SfNeural CNN in ONE tcgen05 kernel.
//
// Replaces NNManager._get_x (nn_manager.py:216-225) and the first four layers of NNManager.create_net
// (nn_manager.py:280-285: Convolution2D(32,5,5) relu, Convolution2D(32,5,5) relu, MaxPooling2D(2,2)). conv2 is half of
// the network's MACs and conv1's 36x36x32 output is its largest activation (166 KB per patch as bf16 hi/lo planes,
// 1.06 GB per 64 frames): computing conv1 per conv2 tile inside the same kernel keeps that activation in shared memory
// and removes the separate gather/expand and conv1 kernels, which were HBM-bound.
//
// conv2 formulation: PIXEL-PAIR rows. With C_out = 32 a plain implicit GEMM has N = 64 / 32 per MMA, and an M = 128,
// K = 16 tcgen05.mma retires every max(N/2, 32 + N/4) cycles (tools/mma_probe.cu): operand-bandwidth bound, 54 % of
// the tensor rate. Here one MMA row is a PAIR of horizontally adjacent outputs (x = 2g, 2g+1) and the B operand is
// the block-Toeplitz pair [W(dy, p) | W(dy, p-1)]: for window position p = 0..5 the A operand is input pixel 2g + p
// and column block j = 0 / 1 accumulates tap dx = p - j of output 2g + j. N doubles (128 for A_hi [W_hi | W_lo],
// 64 for A_lo W_hi) at 6/5 of the MACs: 6240 instead of 8820 tensor cycles per 256 outputs. The A tile keeps even
// and odd input columns in separate sub-planes so that "pixel 2g + p" is again a plain start-address offset
// (row pitch = SBO), and the weights are stored once per (dy, chunk) as the chain T4 T3 T2 T1 T0 of 64-row tap blocks
// with W_hi / W_lo interleaved in 8-row groups: [T_p T_(p-1)] is a contiguous 128-row B operand (SBO = 128 B), and
// its W_hi rows alone are the same address with SBO = 256 B.
//
// One tile = 16 x 16 conv2 outputs of one patch. Per tile:
//   workers  : read the 24 x 24 raw uint8 window of the canonical image and expand it to conv1's A operand X
//              (per pixel the 5-pixel row window, k = dx*3 + c, 16 bf16 = two 16-byte chunks)          -> smem X
//   tensor   : conv1 = 5 vertical taps x 4 M-tiles of 128 window pixels, B = [W1_hi | W1_lo] (N = 64)   -> TMEM D1
//   workers  : D1 -> +bias, ReLU, bf16 hi/lo split -> conv2's A tile [plane][row 20][parity][10][16 B] -> smem A
//   tensor   : conv2 = 5 dy x 6 positions x 2 K-steps x (A_hi, A_lo) as above                         -> TMEM D2
//   workers  : D2 -> hi + lo + correction columns, +bias, ReLU, 2x2 max (in-thread across the pair, one warp shuffle
//              across rows), hi/lo split -> pooled planes in HBM (conv3's input)
// Warp 0 issues all MMAs (one elected thread); warps 1-8 are the workers (two per TMEM lane quarter). The tensor pipe
// executes conv1(0) | conv1(1) conv2(0) | conv1(2) conv2(1) | ...: conv1 of tile i+1 sits between conv2 of tiles i-1
// and i, which is when the workers drain D2 of tile i-1 (they release it as soon as their tcgen05.ld have completed) -
// so one D2 accumulator suffices and TMEM holds D2 = 192 columns + D1 = 4 x 64. The A tile is double buffered
// (epilogue 1 of tile i+1 writes while conv2 of tile i reads); X and D1 are single buffers, rebuilt / drained while a
// conv2 runs.
// Precision: bf16 hi/lo operand split, fp32 accumulation; conv1's uint8 input is exact in bf16.
#include <cuda_bf16.h>

#include "cnn_common.cuh"
#include "tc_ptx.cuh"

#define FR_WORKERS 8
#define FR_THREADS (32 + 32 * FR_WORKERS)

#define FR_XROWS 480                          // 24 x 20 window pixels; the last M-tile's discarded rows read on into
#define FR_XPLANE (FR_XROWS * 16)             //   whatever follows in shared memory (their outputs are never used)
#define FR_XTILE (2 * FR_XPLANE)              // 15360 B
#define FR_ROWPITCH 320                       // conv2 A tile row: [parity 2][10 pairs][16 B]
#define FR_PARITY 160
#define FR_APLANE (20 * FR_ROWPITCH)          // 6400 B
#define FR_ATILE (8 * FR_APLANE)              // 51200 B: 4 chunks x (hi, lo)
#define FR_W1 (5 * 2 * 64 * 16)               // 10240 B  [dy][chunk][W_hi 32 rows | W_lo 32 rows][8]
#define FR_W2 (5 * 4 * 5 * 1024)              // 102400 B [dy][chunk][tap 4..0][H0 L0 H1 L1 H2 L2 H3 L3 groups of 8 rows][8]
#define FR_W2_CHUNK (5 * 1024)
#define FR_W2_DY (4 * FR_W2_CHUNK)
#define FR_SMEM (FR_XTILE + 2 * FR_ATILE + FR_W1 + FR_W2 + 256 + 128)
static_assert(FR_SMEM <= 227 * 1024, "shared memory budget");

struct FrontArgs {
    const uint8_t *goban;   // canonical images [frames][380][380][3]
    const uint4 *w1, *w2;   // packed operand planes (cnn_tc.cu: pack_layer<Conv1Cfg>, pack_conv2_pairs)
    const float *b1, *b2;
    uint4 *out;             // pooled planes [2][4][out_plane], pixel = patch*256 + y*16 + x
    long long out_plane;
    uint4 *dbg_a1;          // optional: conv1 activations [2][4][a1_plane], pixel = patch*1296 + y*36 + x (tests)
    long long a1_plane;
    int n_tiles;            // patches * 4
};

__global__ void __launch_bounds__(FR_THREADS, 1) cnn_tc_front(const __grid_constant__ FrontArgs args)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint8_t *sX = smem;
    uint8_t *sA = sX + FR_XTILE;
    uint8_t *sW1 = sA + 2 * FR_ATILE;
    uint8_t *sW2 = sW1 + FR_W1;
    uint64_t *bars = (uint64_t *)(sW2 + FR_W2);
    const uint32_t b_xfull = smem_u32(bars + 0), b_xempty = smem_u32(bars + 1), b_d1full = smem_u32(bars + 2),
                   b_d1empty = smem_u32(bars + 3), b_afull = smem_u32(bars + 4), b_aempty = smem_u32(bars + 6),
                   b_d2full = smem_u32(bars + 8), b_d2empty = smem_u32(bars + 10), b_wfull = smem_u32(bars + 12);   // (one D2)
    uint32_t *tmem_slot = (uint32_t *)(bars + 13);
    const int warp = warp_index(), lane = threadIdx.x & 31;
    const int n_my = (args.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        mbar_init(b_xfull, FR_WORKERS);   mbar_init(b_xempty, 1);
        mbar_init(b_d1full, 1);           mbar_init(b_d1empty, FR_WORKERS);
        for (int i = 0; i < 2; i++) {
            mbar_init(b_afull + 8 * i, FR_WORKERS);  mbar_init(b_aempty + 8 * i, 1);
        }
        mbar_init(b_d2full, 1);           mbar_init(b_d2empty, FR_WORKERS);
        mbar_init(b_wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t D1_COL = 192;     // D2 accumulator: columns [0, 192); D1: [192, 448) = 4 M-tiles x (hi 32 | lo 32)

    if (warp == 0) {
        // ================================================================= MMA issuer (+ one-off weight load by TMA)
        if (elect_one()) {
            mbar_expect_tx(b_wfull, FR_W1 + FR_W2);
            bulk_g2s(smem_u32(sW1), args.w1, FR_W1, b_wfull);
            for (int off = 0; off < FR_W2; off += 25600) bulk_g2s(smem_u32(sW2 + off), (const uint8_t *)args.w2 + off, 25600, b_wfull);
            constexpr uint32_t IDESC128 = umma_idesc(128), IDESC64 = umma_idesc(64), IDESC32 = umma_idesc(32);
            constexpr uint32_t HI128 = desc_hi(128), HI256 = desc_hi(256), HI_A = desc_hi(FR_ROWPITCH);
            const uint32_t x_lo = desc_lo(smem_u32(sX), FR_XPLANE);
            const uint32_t w1_lo = desc_lo(smem_u32(sW1), 64 * 16), w2_lo = desc_lo(smem_u32(sW2), FR_W2_CHUNK);
            mbar_wait(b_wfull, 0);

            auto conv1 = [&](int i) {
                mbar_wait(b_xfull, i & 1);
                mbar_wait(b_d1empty, (i & 1) ^ 1);
                tc_fence_after();
#pragma unroll
                for (int mt = 0; mt < 4; mt++)
#pragma unroll
                    for (int dy = 0; dy < 5; dy++)
                        tc_mma_bf16(tmem_base + D1_COL + mt * 64, desc64(x_lo + (uint32_t)(mt * 128 + dy * 20), HI128),
                                    desc64(w1_lo + (uint32_t)((dy * 2 * 64 * 16) >> 4), HI128), IDESC64, dy != 0);
                tc_commit(b_xempty);
                tc_commit(b_d1full);
            };

            conv1(0);
            for (int i = 0; i < n_my; i++) {
                const int b = i & 1, k = i >> 1;
                mbar_wait(b_afull + 8 * b, k & 1);          // epilogue 1 of tile i is done: D1 and X are free again
                if (i + 1 < n_my) conv1(i + 1);
                mbar_wait(b_d2empty, (i & 1) ^ 1);          // the workers have read D2 of tile i-1
                tc_fence_after();
                const uint32_t d_tmem = tmem_base;
                const uint32_t a_lo = desc_lo(smem_u32(sA + b * FR_ATILE), FR_APLANE);
#pragma unroll
                for (int dy = 0; dy < 5; dy++) {
#pragma unroll
                    for (int pi = 0; pi < 6; pi++) {
                        const int p = pi == 0 ? 1 : (pi == 1 ? 0 : pi);   // position 1 first: it initialises all 192 columns
#pragma unroll
                        for (int k2 = 0; k2 < 2; k2++) {
                            const uint32_t ao = (uint32_t)((2 * k2 * FR_APLANE + dy * FR_ROWPITCH + (p & 1) * FR_PARITY + (p >> 1) * 16) >> 4);
                            const int tap = p < 5 ? p : 4;               // first tap block of the B operand
                            const uint32_t wo = (uint32_t)((dy * FR_W2_DY + 2 * k2 * FR_W2_CHUNK + (4 - tap) * 1024) >> 4);
                            const uint32_t acc = (dy | pi | k2) != 0;
                            const uint64_t da_hi = desc64(a_lo + ao, HI_A), da_lo = desc64(a_lo + ao + (uint32_t)((4 * FR_APLANE) >> 4), HI_A);
                            if (p >= 1 && p <= 4) {
                                tc_mma_bf16(d_tmem, da_hi, desc64(w2_lo + wo, HI128), IDESC128, acc);
                                tc_mma_bf16(d_tmem + 128, da_lo, desc64(w2_lo + wo, HI256), IDESC64, acc);
                            } else if (p == 0) {        // only output j = 0 has a tap (dx = 0) at this position
                                tc_mma_bf16(d_tmem, da_hi, desc64(w2_lo + wo, HI128), IDESC64, 1);
                                tc_mma_bf16(d_tmem + 128, da_lo, desc64(w2_lo + wo, HI256), IDESC32, 1);
                            } else {                    // p == 5: only output j = 1 (dx = 4)
                                tc_mma_bf16(d_tmem + 64, da_hi, desc64(w2_lo + wo, HI128), IDESC64, 1);
                                tc_mma_bf16(d_tmem + 160, da_lo, desc64(w2_lo + wo, HI256), IDESC32, 1);
                            }
                        }
                    }
                }
                tc_commit(b_aempty + 8 * b);
                tc_commit(b_d2full);
            }
        }
    } else {
        // ============================================================================================== workers
        const int wt = threadIdx.x - 32;                     // 0..255
        const int q = warp & 3, half = (warp - 1) >> 2;      // TMEM lane quarter of this warp; which channel chunks
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);

        // ---- X: raw uint8 window -> conv1 operand rows. Window pixel (ry, rx), ry < 24, rx < 20 -> row ry*20 + rx
        auto build_x = [&](int i) {
            const int tile = (int)blockIdx.x + i * (int)gridDim.x;
            const int patch = tile >> 2, ty = (tile >> 1) & 1, tx = tile & 1;
            const int frame = patch / 100, r = patch % 100;
            const int row0 = cnn_patch_origin(r / 10) + 16 * ty, col0 = cnn_patch_origin(r % 10) + 16 * tx;
            const uint8_t *base = args.goban + (size_t)frame * (380 * 380 * 3);
            if (i > 0) mbar_wait(b_xempty, (i - 1) & 1);     // conv1 of the previous tile has consumed X
            for (int idx = wt; idx < FR_XROWS; idx += 32 * FR_WORKERS) {
                const int ry = idx / 20, rx = idx - ry * 20;
                const uint8_t *src = base + ((size_t)(row0 + ry) * 380 + col0 + rx) * 3;   // 15 bytes: 5 px x BGR
                const uintptr_t a = (uintptr_t)src;
                const uint32_t *wp = (const uint32_t *)(a & ~(uintptr_t)3);
                const uint32_t sh = (uint32_t)(a & 3) * 8;
                uint32_t w[5];
#pragma unroll
                for (int j = 0; j < 4; j++) w[j] = __ldg(wp + j);
                w[4] = sh >= 16 ? __ldg(wp + 4) : 0u;      // only misalignments 2, 3 need a 5th word
                uint32_t v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = __funnelshift_r(w[j], w[j + 1], sh);
                v[3] &= 0x00ffffffu;                        // k = 15 is padding
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    o[2 * j] = pack_bf16x2((float)(v[j] & 0xff), (float)((v[j] >> 8) & 0xff));
                    o[2 * j + 1] = pack_bf16x2((float)((v[j] >> 16) & 0xff), (float)(v[j] >> 24));
                }
                *(uint4 *)(sX + idx * 16) = make_uint4(o[0], o[1], o[2], o[3]);
                *(uint4 *)(sX + FR_XPLANE + idx * 16) = make_uint4(o[4], o[5], o[6], o[7]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_xfull);
        };

        // ---- epilogue 1: D1 (conv1 accumulators) -> bias, ReLU, hi/lo split -> conv2's A tile in shared memory
        auto epilogue1 = [&](int i) {
            const int b = i & 1, k = i >> 1;
            const int tile = (int)blockIdx.x + i * (int)gridDim.x;
            mbar_wait(b_d1full, i & 1);
            mbar_wait(b_aempty + 8 * b, (k & 1) ^ 1);
            tc_fence_after();
            uint8_t *at = sA + b * FR_ATILE;
#pragma unroll
            for (int mt = 0; mt < 4; mt++) {
                const int m = mt * 128 + q * 32 + lane;      // window pixel (wy, wx) = (m / 20, m % 20), 400 valid
                const int wy = m / 20, wx = m - wy * 20;
                const uint32_t aoff = wy * FR_ROWPITCH + (wx & 1) * FR_PARITY + (wx >> 1) * 16;
                const uint32_t taddr = lane_base + D1_COL + mt * 64;
#pragma unroll
                for (int c2 = 0; c2 < 2; c2++) {
                    const int cc = 2 * half + c2;
                    float v[8], u[8];
                    tc_ld8(taddr + 8 * cc, v);
                    tc_ld8(taddr + 32 + 8 * cc, u);
                    tc_ld_wait();
                    const float4 b0 = __ldg((const float4 *)args.b1 + 2 * cc), b1 = __ldg((const float4 *)args.b1 + 2 * cc + 1);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int j = 0; j < 8; j++) v[j] = fmaxf(v[j] + u[j] + bb[j], 0.f);
                    uint4 hi, lo;
                    split8(v, hi, lo);
                    if (m < 400) {
                        *(uint4 *)(at + cc * FR_APLANE + aoff) = hi;
                        *(uint4 *)(at + (4 + cc) * FR_APLANE + aoff) = lo;
                        if (args.dbg_a1) {
                            const int patch = tile >> 2, ty = (tile >> 1) & 1, tx = tile & 1;
                            const long long p = (long long)patch * 1296 + (16 * ty + wy) * 36 + 16 * tx + wx;
                            args.dbg_a1[(long long)cc * args.a1_plane + p] = hi;
                            args.dbg_a1[(long long)(4 + cc) * args.a1_plane + p] = lo;
                        }
                    }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(b_afull + 8 * b);
                mbar_arrive(b_d1empty);
            }
        };

        // ---- epilogue 2: D2 (conv2 accumulators) -> bias, ReLU, 2x2 max-pool, hi/lo split -> HBM
        // MMA row m = 8 * (tile row r) + (pair g): columns [64 j + 16 cc, +16) = W_hi | W_lo parts of output (r, 2g + j),
        // channels 8 cc .. 8 cc + 7; columns [128 + 32 j + 8 cc, +8) = the A_lo W_hi correction.
        auto epilogue2 = [&](int i) {
            const int tile = (int)blockIdx.x + i * (int)gridDim.x;
            const int patch = tile >> 2, ty = (tile >> 1) & 1, tx = tile & 1;
            const int r = q * 4 + (lane >> 3), g = lane & 7;
            const bool writer = (lane & 8) == 0;             // even tile row: owns the pooled pixel (r / 2, g)
            const long long opix = (long long)patch * 256 + (8 * ty + (r >> 1)) * 16 + 8 * tx + g;
            mbar_wait(b_d2full, i & 1);
            tc_fence_after();
            const uint32_t taddr = lane_base;
            float h0[2][16], h1[2][16], l0[2][8], l1[2][8];
#pragma unroll
            for (int c2 = 0; c2 < 2; c2++) {
                const int cc = 2 * half + c2;
                tc_ld16(taddr + 16 * cc, h0[c2]);
                tc_ld16(taddr + 64 + 16 * cc, h1[c2]);
                tc_ld8(taddr + 128 + 8 * cc, l0[c2]);
                tc_ld8(taddr + 160 + 8 * cc, l1[c2]);
            }
            tc_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_d2empty);            // D2 is in registers: conv2 of the next tile may start
#pragma unroll
            for (int c2 = 0; c2 < 2; c2++) {
                const int cc = 2 * half + c2;
                const float4 b0 = __ldg((const float4 *)args.b2 + 2 * cc), b1 = __ldg((const float4 *)args.b2 + 2 * cc + 1);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float x0 = h0[c2][j] + h0[c2][8 + j] + l0[c2][j] + bb[j];
                    const float x1 = h1[c2][j] + h1[c2][8 + j] + l1[c2][j] + bb[j];
                    float x = fmaxf(fmaxf(x0, x1), 0.f);
                    x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 8));
                    v[j] = x;
                }
                if (writer) {
                    uint4 hi, lo;
                    split8(v, hi, lo);
                    args.out[(long long)cc * args.out_plane + opix] = hi;
                    args.out[(long long)(4 + cc) * args.out_plane + opix] = lo;
                }
            }
        };

        build_x(0);
        for (int i = 0; i < n_my; i++) {
            epilogue1(i);
            if (i + 1 < n_my) build_x(i + 1);
            if (i >= 1) epilogue2(i - 1);
        }
        epilogue2(n_my - 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

int ckb_cnn_front_init(ckb_ctx *ctx)
{
    CKB_CUDA(ctx, cudaFuncSetAttribute(cnn_tc_front, cudaFuncAttributeMaxDynamicSharedMemorySize, FR_SMEM));
    return CKB_OK;
}

int ckb_launch_cnn_front(ckb_ctx *ctx, const uint8_t *d_goban, int n_patches, const void *w1, const float *b1, const void *w2,
                         const float *b2, void *p2, long long p2_plane, void *dbg_a1, long long a1_plane, cudaStream_t st)
{
    FrontArgs a;
    a.goban = d_goban;
    a.w1 = (const uint4 *)w1;
    a.w2 = (const uint4 *)w2;
    a.b1 = b1;
    a.b2 = b2;
    a.out = (uint4 *)p2;
    a.out_plane = p2_plane;
    a.dbg_a1 = (uint4 *)dbg_a1;
    a.a1_plane = a1_plane;
    a.n_tiles = n_patches * 4;
    const int grid = a.n_tiles < ctx->num_sms ? a.n_tiles : ctx->num_sms;
    cnn_tc_front<<<grid, FR_THREADS, FR_SMEM, st>>>(a);
    CKB_LAUNCH_CHECK(ctx, "cnn_tc_front");
    return CKB_OK;
}
