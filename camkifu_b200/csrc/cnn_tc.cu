// K4 — forward pass of the SfNeural CNN on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Replaces NNCache.predict_all_stones = 100 x (NNManager._get_x + net.predict) (src/camkifu/stone/nn_cache.py:25-52)
// for the network of NNManager.create_net (nn_manager.py:277-298). This file holds the host side (weight packing,
// workspace, the layer sequence) and the generic layer kernel that runs conv3, conv4 and fc1; the patch gather, conv1,
// conv2 and the first max-pool are one fused kernel of their own (cnn_tc_front.cu), the tail (fc2, softmax, decode) is
// in cnn_simt.cu.
//
// Formulation. Every layer is a "shifted GEMM" over a flat list of pixels. Activations live in HBM as channel-chunk
// planes [plane hi|lo][chunk of 8 channels][pixel][8 x bf16]: a pixel's 8 channels are one 16-byte unit, and consecutive
// pixels of a plane are consecutive units — exactly the un-swizzled K-major core-matrix layout of the tcgen05 shared
// memory descriptors (8 rows x 16 bytes contiguous, SBO = 128 B between row groups, LBO = plane stride between the two
// K chunks). A 128-pixel tile plus its halo is therefore staged with one 1-D bulk copy (TMA, cp.async.bulk) per plane,
// and the A operand of filter tap (dy, dx) is the SAME shared-memory tile with its descriptor start address advanced
// by (dy * row_width + dx) pixels: implicit GEMM without im2col and without any shared-memory re-layout. Outputs are
// computed for every pixel of the input grid; the epilogue keeps the valid ones and compacts them into the next
// layer's grid. The dense layer fc1 runs through the same kernel with the 36 pooled pixels as "taps" whose A tiles
// are streamed next to their weights.
//
// Precision. Inputs are raw 0..255 and the parity bar on the softmax is 1e-3 relative, which single-pass bf16 or tf32
// operands miss by 1-2 orders of magnitude. Each float32 operand is therefore split into bf16 hi + lo
// and three tensor-core products are accumulated in float32:  A_hi W_hi + A_hi W_lo + A_lo W_hi  (the uint8 network
// input is exact in bf16, so conv1 needs two). W_hi | W_lo sit side by side in the B operand so that A_hi is read from
// shared memory once for both (an M = 128, K = 16 MMA retires every max(N/2, 32 + N/4) cycles: tools/mma_probe.cu).
//
// Kernel structure (one persistent CTA per SM, 320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA
// issuer (single thread), warps 2-9 = epilogue (tcgen05.ld -> bias, ReLU, hi/lo split -> coalesced 16-byte stores).
// mbarrier pipelines: A tile full/empty (double buffered), weight/A stage ring full/empty, accumulator full/empty
// (two TMEM accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1).
#include <cuda_bf16.h>

#include <new>

#include "cnn_common.cuh"
#include "tc_ptx.cuh"

size_t ckb_cnn_simt_workspace(int n);
int ckb_launch_fc2_decode(ckb_ctx *ctx, const float *d_f5, int n, void *d_tmp, float *d_softmax, uint8_t *d_stones,
                          float *d_conf, uint8_t *d_keep, cudaStream_t st);

// ------------------------------------------------------------------------------------------------------ layer configs
struct Conv1Cfg {   // geometry only (weights size), as Conv2Cfg: input = 16-channel "row window" expansion (k = dx*3 + c), taps = dy
    static constexpr int NTAPS = 5, GW = 40, HW_IN = 1600, OH = 36, OW = 36, KC = 2, N = 32, A_PLANES = 1;
    static constexpr bool CONCAT = true, A_RES = true, W_RES = true, OUT_F32 = false, POOL_X = false, POOL_XY = false;
    static constexpr int KCS = 2, NSTAGE = 1, NABUF = 2, NACC = 2;
    __host__ __device__ static constexpr int tapoff(int t) { return t * 40; }
};
struct Conv2Cfg {   // geometry only (weights size): conv1 and conv2 run in cnn_tc_front.cu, not through cnn_tc_layer
    static constexpr int NTAPS = 25, GW = 36, HW_IN = 1296, OH = 32, OW = 32, KC = 4, N = 32, A_PLANES = 2;
    static constexpr bool CONCAT = true, A_RES = true, W_RES = true, OUT_F32 = false, POOL_X = false, POOL_XY = false;
    static constexpr int KCS = 4, NSTAGE = 1, NABUF = 2, NACC = 2;
    __host__ __device__ static constexpr int tapoff(int t) { return (t / 5) * 36 + t % 5; }
};
struct Conv3Cfg {
    static constexpr int NTAPS = 9, GW = 16, HW_IN = 256, OH = 14, OW = 14, KC = 4, N = 96, A_PLANES = 2;
    static constexpr bool CONCAT = true, A_RES = true, W_RES = true, OUT_F32 = false, POOL_X = false, POOL_XY = false;
    static constexpr int KCS = 4, NSTAGE = 1, NABUF = 3, NACC = 2;
    __host__ __device__ static constexpr int tapoff(int t) { return (t / 3) * 16 + t % 3; }
};
struct Conv4Cfg {   // POOL_X: the epilogue takes the horizontal half of the 2x2 max-pool that follows (lanes m, m + 1)
    static constexpr int NTAPS = 9, GW = 14, HW_IN = 196, OH = 12, OW = 12, KC = 12, N = 96, A_PLANES = 2;
    static constexpr bool CONCAT = true, A_RES = true, W_RES = false, OUT_F32 = false, POOL_X = true, POOL_XY = false;
    static constexpr int KCS = 4, NSTAGE = 8, NABUF = 2, NACC = 2;
    __host__ __device__ static constexpr int tapoff(int t) { return (t / 3) * 14 + t % 3; }
};
// POOL_XY: the whole 2x2 max-pool in conv4's epilogue, written straight into fc1's per-tap planes (no a4 round trip through
// HBM, no pooling kernel). The vertical partner of a pixel is 14 rows up: 14 lanes away in the same warp, or in the
// previous lane quarter / the previous tile, which is why the epilogue warps exchange their last rows through shared
// memory and every CTA takes a CONTIGUOUS range of tiles (plus the tile before it, recomputed only for its last rows).
struct Conv4PoolCfg : Conv4Cfg {
    static constexpr bool POOL_XY = true;
};
struct Fc1Cfg {     // "pixels" are patches; tap q = pooled pixel, its A tile is streamed with its weights
    static constexpr int NTAPS = 36, GW = 1, HW_IN = 1, OH = 1, OW = 1, KC = 12, N = 160, A_PLANES = 2;
    static constexpr bool CONCAT = false, A_RES = false, W_RES = false, OUT_F32 = true, POOL_X = false, POOL_XY = false;
    static constexpr int KCS = 4, NSTAGE = 5, NABUF = 0, NACC = 2;
    __host__ __device__ static constexpr int tapoff(int) { return 0; }
};

template <class L>
struct Derived {
    static constexpr int HALO = L::A_RES ? L::tapoff(L::NTAPS - 1) : 0;
    static constexpr int APLANE = ((128 + HALO) * 16 + 127) / 128 * 128;         // bytes of one A plane in smem
    static constexpr int A_TILE = L::A_RES ? L::A_PLANES * L::KC * APLANE : 0;   // one resident A tile
    static constexpr int NABUF = L::A_RES ? L::NABUF : 0;
    static constexpr int NACC = L::NACC;
    static constexpr int NB = 2 * L::N;                                          // B rows: W_hi | W_lo
    static constexpr int W_TAP = L::KC * NB * 16;                                // bytes of one tap's weights
    static constexpr int W_ALL = L::W_RES ? L::NTAPS * W_TAP : 0;
    static constexpr int SPT = L::KC / L::KCS;                                   // stages per tap
    static constexpr int W_STAGE = L::W_RES ? 0 : L::KCS * NB * 16;
    static constexpr int A_STAGE = L::A_RES ? 0 : L::A_PLANES * L::KCS * 128 * 16;
    static constexpr int STAGE = W_STAGE + A_STAGE;
    static constexpr int NSTAGE = (L::W_RES && L::A_RES) ? 0 : L::NSTAGE;
    static constexpr int ACC_COLS = L::CONCAT ? 2 * L::N : L::N;
    static constexpr int TMEM_COLS = NACC * ACC_COLS <= 32 ? 32 : NACC * ACC_COLS <= 64 ? 64 : NACC * ACC_COLS <= 128 ? 128
                                     : NACC * ACC_COLS <= 256 ? 256 : 512;
    static constexpr int XCH_PITCH = 96;                                         // floats per exchanged row
    static constexpr int XCH = L::POOL_XY ? 4 * 7 * XCH_PITCH * 4 : 0;           // 7 rows of each lane quarter (what is left of smem)
    static constexpr int XCH_OFF = NABUF * A_TILE + W_ALL + NSTAGE * STAGE + 320;
    static constexpr int SMEM = XCH_OFF + XCH + 128 /*align*/;
    static_assert(L::KC % L::KCS == 0 && L::KCS % 2 == 0, "stages hold whole K=16 steps");
    static_assert(NACC * ACC_COLS <= 512 && NACC <= 4 && NABUF <= 6, "accumulators must fit in TMEM; barrier map");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    static_assert(L::N % 16 == 0 && NB <= 512, "UMMA N constraints (M = 128)");
};

struct LayerArgs {
    const uint4 *in;        // activation planes (bf16 x 8 units)
    long long in_plane;     // plane stride in units (pixels)
    const uint4 *w;         // packed weights [tap][chunk][row 0..2N)[8 bf16]
    const float *bias;      // N floats (zero padded)
    uint4 *out;             // next layer's planes [2][N/8][out_plane]
    long long out_plane;
    float *out_f32;         // OUT_F32: dense [pixel][N] instead
    int n_tiles;            // 128-pixel tiles of the input grid
    int n_patches;
};

// -------------------------------------------------------------------------------------------------- the layer kernel
#define TC_EPI_WARPS 8                        // two epilogue warps per TMEM lane quarter, each takes every other column chunk
#define TC_THREADS (64 + 32 * TC_EPI_WARPS)
template <class L>
__global__ void __launch_bounds__(TC_THREADS, 1) cnn_tc_layer(const __grid_constant__ LayerArgs args)
{
    using D = Derived<L>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint8_t *sA = smem;                                  // NABUF x A_TILE
    uint8_t *sW = sA + D::NABUF * D::A_TILE;             // W_ALL
    uint8_t *sS = sW + D::W_ALL;                         // NSTAGE x STAGE   (W part first, then A part)
    uint64_t *bars = (uint64_t *)(sS + D::NSTAGE * D::STAGE);
    // barrier map
    const uint32_t b_afull = smem_u32(bars + 0);         // [NABUF <= 6]
    const uint32_t b_aempty = smem_u32(bars + 6);        // [NABUF]
    const uint32_t b_tfull = smem_u32(bars + 12);        // [NACC <= 4] accumulator ready
    const uint32_t b_tempty = smem_u32(bars + 16);       // [NACC] accumulator drained
    const uint32_t b_wfull = smem_u32(bars + 20);        // resident weights landed
    const uint32_t b_sfull = smem_u32(bars + 21);        // [NSTAGE <= 8]
    const uint32_t b_sempty = smem_u32(bars + 29);       // [NSTAGE]
    uint32_t *tmem_slot = (uint32_t *)(bars + 37);
    static_assert(D::NSTAGE <= 8, "barrier map");

    const int warp = warp_index(), lane = threadIdx.x & 31;
    // tiles of this CTA: t_first + i * t_step, i < n_my. Strided over the grid, or (POOL_XY) a contiguous range preceded
    // by the tile before it ("lead": computed for the rows the first real tile pools with, nothing of it is stored)
    int t_first = (int)blockIdx.x, t_step = (int)gridDim.x;
    int n_my = ((int)args.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    bool lead = false;
    if constexpr (L::POOL_XY) {
        const int base = args.n_tiles / (int)gridDim.x, extra = args.n_tiles % (int)gridDim.x;
        const int start = (int)blockIdx.x * base + min((int)blockIdx.x, extra);
        lead = start > 0;
        t_first = start - (lead ? 1 : 0);
        t_step = 1;
        n_my = base + ((int)blockIdx.x < extra ? 1 : 0) + (lead ? 1 : 0);
    }

    if (threadIdx.x == 0) {
        for (int i = 0; i < 6; i++) {
            mbar_init(b_afull + 8 * i, 1);
            mbar_init(b_aempty + 8 * i, 1);
        }
        for (int i = 0; i < 4; i++) {
            mbar_init(b_tfull + 8 * i, 1);
            mbar_init(b_tempty + 8 * i, TC_EPI_WARPS);
        }
        mbar_init(b_wfull, 1);
        for (int i = 0; i < 8; i++) {
            mbar_init(b_sfull + 8 * i, 1);
            mbar_init(b_sempty + 8 * i, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)D::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================================================ producer: TMA bulk copies, one elected thread
        if (elect_one()) {
            if (L::W_RES) {
                mbar_expect_tx(b_wfull, (uint32_t)D::W_ALL);
                constexpr int PIECE = 32768;
                for (int off = 0; off < D::W_ALL; off += PIECE)
                    bulk_g2s(smem_u32(sW + off), (const uint8_t *)args.w + off,
                             (uint32_t)(D::W_ALL - off < PIECE ? D::W_ALL - off : PIECE), b_wfull);
            }
            uint32_t sit = 0;  // stage counter
            for (int i = 0; i < n_my; i++) {
                const long long tile = (long long)t_first + (long long)i * t_step;
                if constexpr (L::A_RES) {
                    const int ab = i % D::NABUF;
                    mbar_wait(b_aempty + 8 * ab, ((i / D::NABUF) & 1) ^ 1);
                    mbar_expect_tx(b_afull + 8 * ab, (uint32_t)(L::A_PLANES * L::KC * (128 + D::HALO) * 16));
                    for (int pl = 0; pl < L::A_PLANES * L::KC; pl++)
                        bulk_g2s(smem_u32(sA + ab * D::A_TILE + pl * D::APLANE),
                                 args.in + (long long)pl * args.in_plane + tile * 128, (128 + D::HALO) * 16,
                                 b_afull + 8 * ab);
                }
                if constexpr (D::NSTAGE > 0) {
                    for (int t = 0; t < L::NTAPS; t++)
                        for (int g = 0; g < D::SPT; g++, sit++) {
                            const int s = sit % D::NSTAGE;
                            mbar_wait(b_sempty + 8 * s, ((sit / D::NSTAGE) & 1) ^ 1);
                            mbar_expect_tx(b_sfull + 8 * s, (uint32_t)D::STAGE);
                            uint8_t *dst = sS + s * D::STAGE;
                            bulk_g2s(smem_u32(dst), (const uint8_t *)args.w + (size_t)t * D::W_TAP + (size_t)g * D::W_STAGE,
                                     D::W_STAGE, b_sfull + 8 * s);
                            if (!L::A_RES) {
                                // A of tap t: planes [pl][t][chunk] of the input, 128 rows each
                                for (int pl = 0; pl < L::A_PLANES; pl++)
                                    for (int c = 0; c < L::KCS; c++)
                                        bulk_g2s(smem_u32(dst + D::W_STAGE + (pl * L::KCS + c) * 2048),
                                                 args.in + ((long long)(pl * L::NTAPS + t) * L::KC + g * L::KCS + c) *
                                                               args.in_plane + tile * 128,
                                                 2048, b_sfull + 8 * s);
                            }
                        }
                }
            }
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer: one thread drives the tensor core
        if (elect_one()) {
            constexpr uint32_t IDESC_WIDE = umma_idesc(L::CONCAT ? 2 * L::N : L::N);
            constexpr uint32_t IDESC_N = umma_idesc(L::N);
            constexpr uint32_t HI = desc_hi(128);
            constexpr uint32_t A_LBO = L::A_RES ? D::APLANE : 2048;
            constexpr uint32_t A_LO_OFF = (L::A_RES ? L::KC * D::APLANE : L::KCS * 2048) >> 4;   // hi plane -> lo plane
            constexpr uint32_t A_STEP = (2 * A_LBO) >> 4, W_STEP = (2 * D::NB * 16) >> 4;       // one K = 16 step
            if (L::W_RES) mbar_wait(b_wfull, 0);
            uint32_t sit = 0;
            for (int i = 0; i < n_my; i++) {
                const int acc = i % D::NACC;
                const uint32_t d_tmem = tmem_base + acc * D::ACC_COLS;
                mbar_wait(b_tempty + 8 * acc, ((i / D::NACC) & 1) ^ 1);
                const int ab = L::A_RES ? i % (L::A_RES ? D::NABUF : 1) : 0;
                if (L::A_RES) mbar_wait(b_afull + 8 * ab, (i / (L::A_RES ? D::NABUF : 1)) & 1);
                tc_fence_after();
                const uint32_t a_tile_lo = desc_lo(smem_u32(sA + ab * D::A_TILE), A_LBO);
                const uint32_t w_res_lo = desc_lo(smem_u32(sW), D::NB * 16);
                uint32_t first = 1;
#pragma unroll(L::A_RES ? L::NTAPS : 1)
                for (int t = 0; t < L::NTAPS; t++)
#pragma unroll
                    for (int g = 0; g < D::SPT; g++) {
                        uint32_t a_lo, w_lo;
                        if constexpr (D::NSTAGE > 0) {
                            const int s = sit % D::NSTAGE;
                            mbar_wait(b_sfull + 8 * s, (sit / D::NSTAGE) & 1);
                            tc_fence_after();
                            const uint32_t stage = smem_u32(sS + s * D::STAGE);
                            w_lo = desc_lo(stage, D::NB * 16);
                            a_lo = L::A_RES ? 0u : desc_lo(stage + D::W_STAGE, A_LBO);
                        } else {
                            w_lo = w_res_lo + (uint32_t)((t * D::W_TAP + g * L::KCS * D::NB * 16) >> 4);
                        }
                        if (L::A_RES) a_lo = a_tile_lo + (uint32_t)((g * L::KCS * D::APLANE + L::tapoff(t) * 16) >> 4);
#pragma unroll
                        for (int k2 = 0; k2 < L::KCS / 2; k2++) {
                            const uint64_t da = desc64(a_lo + k2 * A_STEP, HI);
                            const uint64_t db = desc64(w_lo + k2 * W_STEP, HI);
                            if (L::CONCAT) {
                                tc_mma_bf16(d_tmem, da, db, IDESC_WIDE, first ^ 1);           // A_hi [W_hi | W_lo]
                                if (L::A_PLANES == 2)
                                    tc_mma_bf16(d_tmem, desc64(a_lo + A_LO_OFF + k2 * A_STEP, HI), db, IDESC_N, 1);
                            } else {
                                tc_mma_bf16(d_tmem, da, db, IDESC_N, first ^ 1);              // A_hi W_hi
                                tc_mma_bf16(d_tmem, da, desc64(w_lo + k2 * W_STEP + ((L::N * 16) >> 4), HI), IDESC_N, 1);   // A_hi W_lo
                                if (L::A_PLANES == 2)
                                    tc_mma_bf16(d_tmem, desc64(a_lo + A_LO_OFF + k2 * A_STEP, HI), db, IDESC_N, 1);
                            }
                            first = 0;
                        }
                        if constexpr (D::NSTAGE > 0) {
                            tc_commit(b_sempty + 8 * (sit % D::NSTAGE));   // stage reusable once these MMAs retire
                            sit++;
                        }
                    }
                if (L::A_RES) tc_commit(b_aempty + 8 * ab);
                tc_commit(b_tfull + 8 * acc);
            }
        }
    } else {
        // ================================================= epilogue: TMEM -> registers -> bias / ReLU / split -> HBM
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                // which column chunks (j = half, half + 2, ...) it handles
        const int row = q * 32 + lane;
        if constexpr (L::POOL_XY) {
            // conv4 + ReLU + 2x2 max-pool -> fc1's tap planes. Grid rows are GW = 14 pixels: the pixel above row r of a
            // tile is row r - 14. Odd-y pixels ("lower") fetch their upper neighbour's values (already maxed with x + 1) by
            // a shuffle from lane - 14, or for lanes < 14 from the rows the previous quarter's warp (for quarter 0: quarter 3
            // of the previous tile) left in shared memory, and store the pooled pixel.
            static_assert(L::GW == 14 && L::OH == 12 && L::OW == 12 && L::N == 96 && TC_EPI_WARPS == 8, "conv4 geometry");
            constexpr int NJ = 6, XP = D::XCH_PITCH;
            float *xch = (float *)(smem + D::XCH_OFF);
            const uint32_t bar_id = 1 + half;            // the four warps (quarters) that share this half's column chunks
            for (int i = 0; i < n_my; i++) {
                const int tile = t_first + i;
                const int acc = i % D::NACC;
                const int p = tile * 128 + row;
                const int patch = p / L::HW_IN;
                const int rem = p - patch * L::HW_IN;
                const int y = rem / L::GW, x = rem - y * L::GW;
                const bool lower = patch < args.n_patches && (y & 1) && y < L::OH && x < L::OW && !(x & 1) && !(lead && i == 0);
                const int qtap = (y >> 1) * (L::OW / 2) + (x >> 1);
                mbar_wait(b_tfull + 8 * acc, (i / D::NACC) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * D::ACC_COLS;
                float vv[NJ][8], uu[NJ][8];
#pragma unroll
                for (int jj = 0; jj < NJ; jj++) {
                    const int j = half + jj * 2;
                    tc_ld8(taddr + 8 * j, vv[jj]);
                    tc_ld8(taddr + L::N + 8 * j, uu[jj]);
                }
                tc_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_tempty + 8 * acc);
#pragma unroll
                for (int jj = 0; jj < NJ; jj++) {
                    const int j = half + jj * 2;
                    float *v = vv[jj];
                    const float4 b0 = __ldg((const float4 *)args.bias + 2 * j), b1 = __ldg((const float4 *)args.bias + 2 * j + 1);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        v[k] = fmaxf(v[k] + uu[jj][k] + bb[k], 0.f);
                        v[k] = fmaxf(v[k], __shfl_xor_sync(0xffffffffu, v[k], 1));
                    }
                }
                // the last 14 rows of the quarter (their even-x lanes: 18, 20, ... 30) for the next quarter's lanes < 14.
                // Quarter 3's rows are for quarter 0 of the NEXT tile, which still reads the previous ones in this
                // iteration: they are published after the second barrier instead.
                float *dst = xch + (q * 7 + ((lane - 18) >> 1)) * XP;
                const bool pub = lane >= 18 && !(lane & 1);
                if (pub && q < 3) {
#pragma unroll
                    for (int jj = 0; jj < NJ; jj++) {
                        const int j = half + jj * 2;
                        *(float4 *)(dst + 8 * j) = make_float4(vv[jj][0], vv[jj][1], vv[jj][2], vv[jj][3]);
                        *(float4 *)(dst + 8 * j + 4) = make_float4(vv[jj][4], vv[jj][5], vv[jj][6], vv[jj][7]);
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                const float *src = xch + (((q + 3) & 3) * 7 + (lane >> 1)) * XP;
#pragma unroll
                for (int jj = 0; jj < NJ; jj++) {
                    const int j = half + jj * 2;
                    float u[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) u[k] = __shfl_up_sync(0xffffffffu, vv[jj][k], 14);
                    if (lower) {
                        if (lane < 14) {
                            const float4 s0 = *(const float4 *)(src + 8 * j), s1 = *(const float4 *)(src + 8 * j + 4);
                            u[0] = s0.x; u[1] = s0.y; u[2] = s0.z; u[3] = s0.w;
                            u[4] = s1.x; u[5] = s1.y; u[6] = s1.z; u[7] = s1.w;
                        }
#pragma unroll
                        for (int k = 0; k < 8; k++) u[k] = fmaxf(u[k], vv[jj][k]);
                        uint4 hi, lo;
                        split8(u, hi, lo);
                        args.out[(long long)(qtap * L::KC + j) * args.out_plane + patch] = hi;
                        args.out[(long long)((L::OH / 2 * (L::OW / 2) + qtap) * L::KC + j) * args.out_plane + patch] = lo;
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // the exchange rows may be overwritten
                if (pub && q == 3) {
#pragma unroll
                    for (int jj = 0; jj < NJ; jj++) {
                        const int j = half + jj * 2;
                        *(float4 *)(dst + 8 * j) = make_float4(vv[jj][0], vv[jj][1], vv[jj][2], vv[jj][3]);
                        *(float4 *)(dst + 8 * j + 4) = make_float4(vv[jj][4], vv[jj][5], vv[jj][6], vv[jj][7]);
                    }
                }
            }
        } else
        for (int i = 0; i < n_my; i++) {
            const int tile = t_first + i * t_step;
            const int acc = i % D::NACC;
            const int p = tile * 128 + row;              // flat pixel of the input grid (< 2^31 for 64 frames)
            const int patch = p / L::HW_IN;
            const int rem = p - patch * L::HW_IN;
            const int y = rem / L::GW, x = rem - y * L::GW;
            const bool valid = patch < args.n_patches && y < L::OH && x < L::OW && !(L::POOL_X && (x & 1));
            // POOL_X: rows m, m+1 of a tile are horizontally adjacent pixels (tile starts and grid widths are even);
            // the even one keeps max(x, x+1) and the output grid is OH x OW/2
            const long long opix = L::POOL_X ? (long long)patch * (L::OH * L::OW / 2) + y * (L::OW / 2) + (x >> 1)
                                             : (long long)patch * (L::OH * L::OW) + y * L::OW + x;
            mbar_wait(b_tfull + 8 * acc, (i / D::NACC) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * D::ACC_COLS;
            // all TMEM loads of this warp's column chunks first, one wait, then the arithmetic and the stores: the
            // accumulator is released as soon as it is in registers
            constexpr int NJ = L::N / 8 / (TC_EPI_WARPS / 4);     // chunks of 8 channels per warp
            static_assert(L::N / 8 % (TC_EPI_WARPS / 4) == 0, "column chunks split evenly over the epilogue warps");
            float vv[NJ][8], uu[L::CONCAT ? NJ : 1][8];
#pragma unroll
            for (int jj = 0; jj < NJ; jj++) {
                const int j = half + jj * (TC_EPI_WARPS / 4);
                tc_ld8(taddr + 8 * j, vv[jj]);
                if (L::CONCAT) tc_ld8(taddr + L::N + 8 * j, uu[jj]);
            }
            tc_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_tempty + 8 * acc);
#pragma unroll
            for (int jj = 0; jj < NJ; jj++) {
                const int j = half + jj * (TC_EPI_WARPS / 4);
                float *v = vv[jj];
                if (L::CONCAT) {
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] += uu[jj][k];
                }
                const float4 b0 = __ldg((const float4 *)args.bias + 2 * j), b1 = __ldg((const float4 *)args.bias + 2 * j + 1);
                v[0] = fmaxf(v[0] + b0.x, 0.f); v[1] = fmaxf(v[1] + b0.y, 0.f);
                v[2] = fmaxf(v[2] + b0.z, 0.f); v[3] = fmaxf(v[3] + b0.w, 0.f);
                v[4] = fmaxf(v[4] + b1.x, 0.f); v[5] = fmaxf(v[5] + b1.y, 0.f);
                v[6] = fmaxf(v[6] + b1.z, 0.f); v[7] = fmaxf(v[7] + b1.w, 0.f);
                if (L::POOL_X) {
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = fmaxf(v[k], __shfl_xor_sync(0xffffffffu, v[k], 1));
                }
                if (valid) {
                    if (L::OUT_F32) {
                        float4 *o = (float4 *)(args.out_f32 + opix * L::N + 8 * j);
                        o[0] = make_float4(v[0], v[1], v[2], v[3]);
                        o[1] = make_float4(v[4], v[5], v[6], v[7]);
                    } else {
                        uint4 hi, lo;
                        split8(v, hi, lo);
                        args.out[(long long)j * args.out_plane + opix] = hi;
                        args.out[(long long)(L::N / 8 + j) * args.out_plane + opix] = lo;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)D::TMEM_COLS)
                     : "memory");
    }
}


// ------------------------------------------------------------------------------------------- elementwise companions
__device__ __forceinline__ void unpack8(const uint4 hi, const uint4 lo, float *v)
{
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        v[2 * i] = __uint_as_float(h[i] << 16) + __uint_as_float(l[i] << 16);
        v[2 * i + 1] = __uint_as_float(h[i] & 0xffff0000u) + __uint_as_float(l[i] & 0xffff0000u);
    }
}

// test aid: planes [2][KC][plane] -> dense float32 [pixel][C]
__global__ void cnn_tc_unpack(const uint4 *__restrict__ in, long long in_plane, int KC, long long n_pix, int C,
                              float *__restrict__ out)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pix * KC) return;
    const long long p = idx % n_pix;
    const int c = (int)(idx / n_pix);
    float v[8];
    unpack8(in[(long long)c * in_plane + p], in[(long long)(KC + c) * in_plane + p], v);
    for (int k = 0; k < 8; k++)
        if (c * 8 + k < C) out[p * C + c * 8 + k] = v[k];
}

// test aid: fc1's per-tap planes (the pooled conv4 output: plane (pl * 36 + q) * 12 + chunk, row = patch) -> dense float32
// [patch][q = y * 6 + x][90]
__global__ void cnn_tc_unpack_taps(const uint4 *__restrict__ in, long long in_plane, int n_patches, float *__restrict__ out)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_patches * 36 * 12) return;
    const int patch = (int)(idx % n_patches);
    const int qc = (int)(idx / n_patches), q = qc / 12, c = qc - 12 * q;
    float v[8];
    unpack8(in[(long long)(q * 12 + c) * in_plane + patch], in[(long long)((36 + q) * 12 + c) * in_plane + patch], v);
    for (int k = 0; k < 8; k++)
        if (c * 8 + k < 90) out[((size_t)patch * 36 + q) * 90 + c * 8 + k] = v[k];
}

// ------------------------------------------------------------------------------------------------ host: weight packing
static inline uint16_t bf16_rn(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7fffu + ((u >> 16) & 1u);   // round to nearest even (parameters are finite, checked by the caller)
    return (uint16_t)(u >> 16);
}
static inline float bf16_f(uint16_t h)
{
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

struct TcBlob {
    size_t off_w[5], off_b[5], total;
};

// packed device blob: per layer the bf16 operand planes [tap][chunk][row 0..2N)[8] followed by N padded float32 biases
template <class L, class F>
static void pack_layer(uint8_t *blob, size_t off_w, size_t off_b, F weight /*(tap, k, n) -> float*/, const float *bias,
                       int n_real)
{
    uint16_t *w = (uint16_t *)(blob + off_w);
    for (int t = 0; t < L::NTAPS; t++)
        for (int c = 0; c < L::KC; c++)
            for (int r = 0; r < 2 * L::N; r++)
                for (int j = 0; j < 8; j++) {
                    const int n = r % L::N;
                    const float f = weight(t, c * 8 + j, n);
                    const uint16_t hi = bf16_rn(f);
                    const uint16_t v = r < L::N ? hi : bf16_rn(f - bf16_f(hi));
                    w[(((size_t)t * L::KC + c) * 2 * L::N + r) * 8 + j] = v;
                }
    float *b = (float *)(blob + off_b);
    for (int n = 0; n < L::N; n++) b[n] = n < n_real ? bias[n] : 0.f;
}

template <class L>
static size_t layer_w_bytes() { return (size_t)L::NTAPS * L::KC * 2 * L::N * 16; }

static TcBlob blob_layout()
{
    TcBlob b;
    size_t o = 0;
    const size_t wb[5] = {layer_w_bytes<Conv1Cfg>(), layer_w_bytes<Conv2Cfg>(), layer_w_bytes<Conv3Cfg>(),
                          layer_w_bytes<Conv4Cfg>(), layer_w_bytes<Fc1Cfg>()};
    const int nn[5] = {Conv1Cfg::N, Conv2Cfg::N, Conv3Cfg::N, Conv4Cfg::N, Fc1Cfg::N};
    for (int i = 0; i < 5; i++) {
        b.off_w[i] = o; o += (wb[i] + 255) / 256 * 256;
        b.off_b[i] = o; o += ((size_t)nn[i] * 4 + 255) / 256 * 256;
    }
    b.total = o;
    return b;
}

int ckb_cnn_tc_pack(ckb_ctx *ctx, const float *p)
{
    const TcBlob L = blob_layout();
    uint8_t *h = new (std::nothrow) uint8_t[L.total];
    if (!h) CKB_FAIL(ctx, CKB_E_NOMEM, "host allocation failed");
    memset(h, 0, L.total);
    const float *w1 = p + OFF_W1, *w2 = p + OFF_W2, *w3 = p + OFF_W3, *w4 = p + OFF_W4, *w5 = p + OFF_W5;
    // conv1: tap = dy, k = dx*3 + c
    pack_layer<Conv1Cfg>(h, L.off_w[0], L.off_b[0],
                         [&](int t, int k, int n) { return k < 15 ? w1[((t * 5 + k / 3) * 3 + k % 3) * 32 + n] : 0.f; },
                         p + OFF_B1, 32);
    // conv2 for cnn_tc_front's pixel-pair formulation: per (dy, 8-channel chunk) the chain of tap blocks dx = 4..0, each
    // 64 rows = W_hi / W_lo of the 32 output channels interleaved in groups of 8 rows: [H0-7 L0-7 H8-15 L8-15 ...]
    {
        uint16_t *w = (uint16_t *)(h + L.off_w[1]);
        for (int dy = 0; dy < 5; dy++)
            for (int c = 0; c < 4; c++)
                for (int dx = 0; dx < 5; dx++)
                    for (int co = 0; co < 32; co++)
                        for (int e = 0; e < 8; e++) {
                            const float f = w2[(((dy * 5 + dx) * 32) + c * 8 + e) * 32 + co];
                            const uint16_t hi = bf16_rn(f), lo = bf16_rn(f - bf16_f(hi));
                            const size_t base = ((size_t)(dy * 4 + c) * 5 + (4 - dx)) * 512;   // in bf16 elements (1 KB per tap block)
                            w[base + (2 * (co / 8)) * 64 + (co % 8) * 8 + e] = hi;
                            w[base + (2 * (co / 8) + 1) * 64 + (co % 8) * 8 + e] = lo;
                        }
        float *b = (float *)(h + L.off_b[1]);
        for (int n = 0; n < 32; n++) b[n] = p[OFF_B2 + n];
    }
    pack_layer<Conv3Cfg>(h, L.off_w[2], L.off_b[2],
                         [&](int t, int k, int n) { return n < 90 ? w3[(t * 32 + k) * 90 + n] : 0.f; }, p + OFF_B3, 90);
    pack_layer<Conv4Cfg>(h, L.off_w[3], L.off_b[3],
                         [&](int t, int k, int n) { return (k < 90 && n < 90) ? w4[(t * 90 + k) * 90 + n] : 0.f; },
                         p + OFF_B4, 90);
    // fc1: tap = pooled pixel q = y*6 + x, k = channel; Keras Flatten is (H, W, C): feature = q*90 + c
    pack_layer<Fc1Cfg>(h, L.off_w[4], L.off_b[4],
                       [&](int t, int k, int n) { return k < 90 ? w5[((size_t)t * 90 + k) * 160 + n] : 0.f; },
                       p + OFF_B5, 160);
    cudaError_t e = cudaMalloc(&ctx->cnn->d_tc, L.total);
    if (e == cudaSuccess) e = cudaMemcpy(ctx->cnn->d_tc, h, L.total, cudaMemcpyHostToDevice);
    delete[] h;
    if (e != cudaSuccess) CKB_FAIL(ctx, CKB_E_CUDA, "tensor-core weight upload failed: %s", cudaGetErrorString(e));
    ctx->cnn->tc_bytes = L.total;
    // opt in to the large dynamic shared memory footprints once
    CKB_CUDA(ctx, cudaFuncSetAttribute(cnn_tc_layer<Conv3Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Derived<Conv3Cfg>::SMEM));
    CKB_CUDA(ctx, cudaFuncSetAttribute(cnn_tc_layer<Conv4PoolCfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Derived<Conv4PoolCfg>::SMEM));
    CKB_CUDA(ctx, cudaFuncSetAttribute(cnn_tc_layer<Fc1Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Derived<Fc1Cfg>::SMEM));
    const int rc = ckb_cnn_tail_init(ctx);
    if (rc != CKB_OK) return rc;
    return ckb_cnn_front_init(ctx);
}

void ckb_cnn_tc_free(ckb_ctx *ctx)
{
    if (ctx->cnn && ctx->cnn->d_tc) { cudaFree(ctx->cnn->d_tc); ctx->cnn->d_tc = nullptr; }
}

// ---------------------------------------------------------------------------------------------------- workspace layout
#define TC_MAX_FRAMES 64   // frames per internal pass (bounds the workspace: ~53 MB per frame)

struct TcWork {
    // plane strides in 16-byte units (pixels), each with a tail so that the last tile's halo read stays inside
    long long a1_plane, p2_plane, a3_plane, p4_plane;
    size_t a1, p2, a3, p4, f5, tmp, total;
};

static long long plane_units(long long pixels, int halo) { return ((pixels + 127) / 128 * 128 + halo + 7) / 8 * 8; }

// dump_a1: also reserve room for conv1's activations, which normally never leave shared memory (test aid)
static TcWork tc_work_layout(int nf, bool dump_a1)
{
    const long long P = (long long)nf * 100;
    TcWork w;
    w.a1_plane = dump_a1 ? plane_units(P * 1296, 0) : 0;
    w.p2_plane = plane_units(P * 256, Derived<Conv3Cfg>::HALO);
    w.a3_plane = plane_units(P * 196, Derived<Conv4Cfg>::HALO);
    w.p4_plane = plane_units(P, 0);
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255) / 256 * 256; return at; };
    w.p2 = take((size_t)w.p2_plane * 16 * 8);
    w.a3 = take((size_t)w.a3_plane * 16 * 24);
    w.p4 = take((size_t)w.p4_plane * 16 * 2 * 36 * 12);
    w.f5 = take((size_t)P * 160 * 4);
    w.tmp = take((size_t)P * (81 + 81 + 2) * 4);
    w.a1 = take((size_t)w.a1_plane * 16 * 8);
    w.total = o;
    return w;
}

extern "C" size_t ckb_cnn_workspace(const ckb_ctx *ctx, int n)
{
    if (!ctx || n < 0) return 0;
    return tc_work_layout(n < TC_MAX_FRAMES ? n : TC_MAX_FRAMES, ctx->cnn && ctx->cnn->dump_a1).total + 256;
}

template <class L>
static int launch_layer(ckb_ctx *ctx, const char *name, const void *in, long long in_plane, size_t off_w, size_t off_b,
                        void *out, long long out_plane, float *out_f32, long long n_pixels, int n_patches, cudaStream_t st)
{
    LayerArgs a;
    a.in = (const uint4 *)in;
    a.in_plane = in_plane;
    a.w = (const uint4 *)((const uint8_t *)ctx->cnn->d_tc + off_w);
    a.bias = (const float *)((const uint8_t *)ctx->cnn->d_tc + off_b);
    a.out = (uint4 *)out;
    a.out_plane = out_plane;
    a.out_f32 = out_f32;
    a.n_tiles = (int)((n_pixels + 127) / 128);
    a.n_patches = n_patches;
    const int grid = a.n_tiles < ctx->num_sms ? a.n_tiles : ctx->num_sms;
    cnn_tc_layer<L><<<grid, TC_THREADS, Derived<L>::SMEM, st>>>(a);
    CKB_LAUNCH_CHECK(ctx, name);
    return CKB_OK;
}

#define TC_TRY(call)                 \
    do {                             \
        const int rc__ = (call);     \
        if (rc__ != CKB_OK) return rc__; \
    } while (0)

static int tc_forward_pass(ckb_ctx *ctx, const uint8_t *d_goban, int nf, uint8_t *work, float *d_softmax,
                           uint8_t *d_stones, float *d_conf, uint8_t *d_keep, cudaStream_t st)
{
    const bool dump = ctx->cnn->dump_a1 != 0;
    const TcWork W = tc_work_layout(nf, dump);
    const TcBlob B = blob_layout();
    const int P = nf * 100;
    uint4 *p2 = (uint4 *)(work + W.p2), *a3 = (uint4 *)(work + W.a3);
    uint4 *p4 = (uint4 *)(work + W.p4);
    float *f5 = (float *)(work + W.f5);
    const uint8_t *tc = (const uint8_t *)ctx->cnn->d_tc;
    // gather + conv1 + conv2 + pool in one kernel (cnn_tc_front.cu)
    TC_TRY(ckb_launch_cnn_front(ctx, d_goban, P, tc + B.off_w[0], (const float *)(tc + B.off_b[0]), tc + B.off_w[1],
                                (const float *)(tc + B.off_b[1]), p2, W.p2_plane, dump ? work + W.a1 : nullptr, W.a1_plane, st));
    TC_TRY(launch_layer<Conv3Cfg>(ctx, "cnn_tc_conv3", p2, W.p2_plane, B.off_w[2], B.off_b[2], a3, W.a3_plane, nullptr,
                                  (long long)P * 256, P, st));
    // conv4 + ReLU + the second 2x2 max-pool, written as fc1's per-tap planes
    TC_TRY(launch_layer<Conv4PoolCfg>(ctx, "cnn_tc_conv4", a3, W.a3_plane, B.off_w[3], B.off_b[3], p4, W.p4_plane, nullptr,
                                      (long long)P * 196, P, st));
    TC_TRY(launch_layer<Fc1Cfg>(ctx, "cnn_tc_fc1", p4, W.p4_plane, B.off_w[4], B.off_b[4], nullptr, 0, f5, P, P, st));
    return ckb_launch_fc2_decode(ctx, f5, nf, work + W.tmp, d_softmax, d_stones, d_conf, d_keep, st);
}

extern "C" int ckb_cnn_forward(ckb_ctx *ctx, const uint8_t *d_goban, int n, void *d_work, size_t work_bytes,
                               float *d_softmax, uint8_t *d_stones, float *d_conf, uint8_t *d_keep, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    if (!ctx->cnn || !ctx->cnn->d_tc) CKB_FAIL(ctx, CKB_E_STATE, "ckb_cnn_forward: call ckb_set_cnn_weights first");
    if (n == 0) return CKB_OK;   // an empty batch is a no-op, whatever the pointers
    if (!d_goban || !d_work || n < 0) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_cnn_forward: bad argument");
    if (work_bytes < ckb_cnn_workspace(ctx, n)) CKB_FAIL(ctx, CKB_E_NOMEM, "ckb_cnn_forward: workspace too small");
    if (((uintptr_t)d_work & 255) != 0) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_cnn_forward: workspace must be 256-byte aligned");
    if (n == 0) return CKB_OK;
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    CKB_ENTER(ctx, stream);
    for (int f0 = 0; f0 < n; f0 += TC_MAX_FRAMES) {
        const int nf = n - f0 < TC_MAX_FRAMES ? n - f0 : TC_MAX_FRAMES;
        TC_TRY(tc_forward_pass(ctx, d_goban + (size_t)f0 * 380 * 380 * 3, nf, (uint8_t *)d_work,
                               d_softmax ? d_softmax + (size_t)f0 * 100 * 81 : nullptr,
                               d_stones ? d_stones + (size_t)f0 * 361 : nullptr, d_conf ? d_conf + (size_t)f0 * 361 : nullptr,
                               d_keep ? d_keep + (size_t)f0 * 361 : nullptr, (cudaStream_t)stream));
    }
    return CKB_OK;
}

// Test aid: keep conv1's activations (which normally live only in shared memory) in the workspace during the next
// forward passes, so that ckb_cnn_debug_activation(layer 1) can return them. Changes ckb_cnn_workspace().
extern "C" int ckb_cnn_set_debug(ckb_ctx *ctx, int on)
{
    if (!ctx || !ctx->cnn) return CKB_E_INVALID;
    ctx->cnn->dump_a1 = on != 0;
    return CKB_OK;
}

// Test aid: after ckb_cnn_forward on n <= 64 frames, unpack one intermediate activation of the tensor-core path from
// the workspace into dense float32 [patch][H][W][C]: layer 1 = conv1 (36,36,32), 2 = pooled conv2 (16,16,32),
// 3 = conv3 (14,14,90), 4 = pooled conv4 (6,6,90; from fc1's tap planes, written by conv4's epilogue), 5 = fc1 (160).
extern "C" int ckb_cnn_debug_activation(ckb_ctx *ctx, const void *d_work, int n, int layer, float *d_out, void *stream)
{
    if (!ctx || !d_work || !d_out || n < 1 || n > TC_MAX_FRAMES) return CKB_E_INVALID;
    const bool dump = ctx->cnn && ctx->cnn->dump_a1;
    const TcWork W = tc_work_layout(n, dump);
    const uint8_t *work = (const uint8_t *)d_work;
    cudaStream_t st = (cudaStream_t)stream;
    CKB_ENTER(ctx, stream);
    const long long P = (long long)n * 100;
    auto go = [&](size_t off, long long plane, int KC, long long npix, int C) {
        cnn_tc_unpack<<<(unsigned)((npix * KC + 255) / 256), 256, 0, st>>>((const uint4 *)(work + off), plane, KC, npix, C, d_out);
    };
    switch (layer) {
    case 1:
        if (!dump) CKB_FAIL(ctx, CKB_E_STATE, "ckb_cnn_debug_activation: layer 1 needs ckb_cnn_set_debug(ctx, 1) before the forward pass");
        go(W.a1, W.a1_plane, 4, P * 1296, 32);
        break;
    case 2: go(W.p2, W.p2_plane, 4, P * 256, 32); break;
    case 3: go(W.a3, W.a3_plane, 12, P * 196, 90); break;
    case 4:
        cnn_tc_unpack_taps<<<(unsigned)((P * 36 * 12 + 255) / 256), 256, 0, st>>>((const uint4 *)(work + W.p4), W.p4_plane, (int)P, d_out);
        break;
    case 5: CKB_CUDA(ctx, cudaMemcpyAsync(d_out, work + W.f5, (size_t)P * 160 * 4, cudaMemcpyDeviceToDevice, st)); return CKB_OK;
    default: CKB_FAIL(ctx, CKB_E_INVALID, "ckb_cnn_debug_activation: layer must be 1 .. 5");
    }
    CKB_LAUNCH_CHECK(ctx, "cnn_tc_unpack");
    return CKB_OK;
}
