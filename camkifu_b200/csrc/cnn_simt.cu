// K4, verification variant — the SfNeural forward pass on plain fp32 CUDA cores (FMA), plus the softmax / decode
// kernels shared with the tensor-core path.
//
// Replaces NNCache.predict_all_stones (src/camkifu/stone/nn_cache.py:25-52): 100 x (NNManager._get_x + net.predict) and
// the base-3 decode of NNManager.compute_stones (nn_manager.py:246-254) with confidence max(y)/sum(y), and the
// MIN_CONFIDENCE = 0.6 rule of SfNeural.predict_all (sf_neural.py:18,57-70).
//
// This direct-convolution path exists so that every tensor-core layer can be checked on the device against an
// independent fp32 implementation (tests/test_gpu_cnn.py); ckb_cnn_forward (cnn_tc.cu) is the product path.
#include "cnn_common.cuh"

// out[p][oy][ox][co] = relu(b[co] + sum_{dy,dx,c} in[p][oy+dy][ox+dx][c] * w[dy][dx][c][co]); thread <-> (p, oy, ox, co)
// U8IN: the input is the canonical image itself and patch p = (frame, i, j) is read in place (NNManager._get_x).
template <int KH, int KW, int CIN, int COUT, int IH, int IW, bool U8IN>
__global__ void __launch_bounds__(256) cnn_conv_relu_simt(const void *__restrict__ in, const float *__restrict__ w,
                                                          const float *__restrict__ b, float *__restrict__ out,
                                                          int n_patches)
{
    constexpr int OH = IH - KH + 1, OW = IW - KW + 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n_patches * OH * OW * COUT;
    if (idx >= total) return;
    const int co = (int)(idx % COUT);
    long long t = idx / COUT;
    const int ox = (int)(t % OW);
    t /= OW;
    const int oy = (int)(t % OH);
    const int p = (int)(t / OH);
    float acc = 0.f;
    if (U8IN) {
        const int frame = p / 100, r = p % 100;
        const int x0 = cnn_patch_origin(r / 10), y0 = cnn_patch_origin(r % 10);
        const uint8_t *img = (const uint8_t *)in + (size_t)frame * 380 * 380 * 3;
        for (int dy = 0; dy < KH; dy++)
            for (int dx = 0; dx < KW; dx++) {
                const uint8_t *ip = img + ((size_t)(x0 + oy + dy) * 380 + (y0 + ox + dx)) * 3;
                const float *wp = w + (size_t)((dy * KW + dx) * CIN) * COUT + co;
#pragma unroll
                for (int c = 0; c < CIN; c++) acc = fmaf((float)__ldg(ip + c), __ldg(wp + c * COUT), acc);
            }
    } else {
        const float *base = (const float *)in + (size_t)p * IH * IW * CIN;
        for (int dy = 0; dy < KH; dy++)
            for (int dx = 0; dx < KW; dx++) {
                const float *ip = base + ((size_t)(oy + dy) * IW + (ox + dx)) * CIN;
                const float *wp = w + (size_t)((dy * KW + dx) * CIN) * COUT + co;
#pragma unroll 8
                for (int c = 0; c < CIN; c++) acc = fmaf(__ldg(ip + c), __ldg(wp + c * COUT), acc);
            }
    }
    acc += __ldg(b + co);
    out[idx] = acc > 0.f ? acc : 0.f;
}

template <int IH, int IW, int C>
__global__ void __launch_bounds__(256) cnn_maxpool2_simt(const float *__restrict__ in, float *__restrict__ out,
                                                         int n_patches)
{
    constexpr int OH = IH / 2, OW = IW / 2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_patches * OH * OW * C) return;
    const int c = (int)(idx % C);
    long long t = idx / C;
    const int ox = (int)(t % OW);
    t /= OW;
    const int oy = (int)(t % OH);
    const int p = (int)(t / OH);
    const float *q = in + ((size_t)p * IH * IW + (size_t)(2 * oy) * IW + 2 * ox) * C + c;
    out[idx] = fmaxf(fmaxf(q[0], q[C]), fmaxf(q[(size_t)IW * C], q[(size_t)IW * C + C]));
}

template <int NIN, int NOUT, bool RELU>
__global__ void __launch_bounds__(256) cnn_dense_simt(const float *__restrict__ in, const float *__restrict__ w,
                                                      const float *__restrict__ b, float *__restrict__ out,
                                                      int n_patches)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_patches * NOUT) return;
    const int o = (int)(idx % NOUT);
    const int p = (int)(idx / NOUT);
    const float *ip = in + (size_t)p * NIN;
    float acc = 0.f;
#pragma unroll 8
    for (int i = 0; i < NIN; i++) acc = fmaf(__ldg(ip + i), __ldg(w + (size_t)i * NOUT + o), acc);
    acc += __ldg(b + o);
    out[idx] = RELU ? (acc > 0.f ? acc : 0.f) : acc;
}

// ---------------------------------------------------------------------------------------------------- softmax + decode
// One warp per region: softmax over 81 logits, then label = argmax (first maximum, as np.argmax), confidence =
// max(y) / sum(y) with Python's sequential float32 sum (nn_cache.py:28-30). One thread per intersection then applies
// the write order of predict_all_stones (regions in i, j order: region 9 overwrites row / column 17 of region 8).
__global__ void __launch_bounds__(128) cnn_softmax_label(const float *__restrict__ logits, int n_regions,
                                                         float *__restrict__ softmax, int *__restrict__ label,
                                                         float *__restrict__ conf)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_regions) return;
    const float *z = logits + (size_t)warp * CNN_F6;
    float v[3];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int o = lane + 32 * k;
        v[k] = o < CNN_F6 ? z[o] : -INFINITY;
        m = fmaxf(m, v[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        v[k] = (lane + 32 * k) < CNN_F6 ? expf(v[k] - m) : 0.f;
        s += v[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float *y = softmax + (size_t)warp * CNN_F6;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int o = lane + 32 * k;
        if (o < CNN_F6) y[o] = __fdiv_rn(v[k], s);
    }
    __syncwarp();
    if (lane == 0) {
        int best = 0;
        float bv = y[0], tot = 0.f;
        for (int o = 0; o < CNN_F6; o++) {
            const float t = y[o];
            if (t > bv) { bv = t; best = o; }
            tot = __fadd_rn(tot, t);
        }
        label[warp] = best;
        conf[warp] = __fdiv_rn(bv, tot);
    }
}

__global__ void __launch_bounds__(128) cnn_decode_board(const int *__restrict__ label, const float *__restrict__ conf,
                                                        int n_frames, uint8_t *__restrict__ stones,
                                                        float *__restrict__ conf_out, uint8_t *__restrict__ keep)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_frames * 361) return;
    const int f = idx / 361, q = idx % 361, r = q / 19, c = q % 19;
    const int i = r < 17 ? r >> 1 : 9, j = c < 17 ? c >> 1 : 9;   // last region that writes (r, c)
    const int a = r - (i < 9 ? 2 * i : 17), b = c - (j < 9 ? 2 * j : 17);
    const int reg = f * 100 + i * 10 + j;
    int k = label[reg];
    const int pw = a * 2 + b;  // base-3 digit index (nn_manager.py:236-254)
    for (int t = 0; t < pw; t++) k /= 3;
    const uint8_t s = (uint8_t)(k % 3);
    const float cf = conf[reg];
    if (stones) stones[idx] = s;
    if (conf_out) conf_out[idx] = cf;
    if (keep) keep[idx] = (s != CKB_E && cf > 0.6f) ? 1 : 0;
}

// d_softmax_tmp: n*100*81 floats used when the caller does not want the softmax; d_lab / d_cf live behind it
int ckb_launch_decode(ckb_ctx *ctx, const float *d_logits, int n, float *d_softmax_or_null, float *d_softmax_tmp,
                      uint8_t *d_stones, float *d_conf, uint8_t *d_keep, cudaStream_t st)
{
    const int nreg = n * 100;
    float *sm = d_softmax_or_null ? d_softmax_or_null : d_softmax_tmp;
    int *lab = (int *)(d_softmax_tmp + (size_t)nreg * CNN_F6);
    float *cf = (float *)(lab + nreg);
    cnn_softmax_label<<<(nreg * 32 + 127) / 128, 128, 0, st>>>(d_logits, nreg, sm, lab, cf);
    CKB_LAUNCH_CHECK(ctx, "cnn_softmax_label");
    cnn_decode_board<<<(n * 361 + 127) / 128, 128, 0, st>>>(lab, cf, n, d_stones, d_conf, d_keep);
    CKB_LAUNCH_CHECK(ctx, "cnn_decode_board");
    return CKB_OK;
}

// Tail of the tensor-core path (cnn_tc.cu): fc2 (Dense(81), nn_manager.py:295; 13 k MAC per patch, float32 FMA in the
// same sequential order as cnn_dense_simt) + softmax + label / confidence in one kernel. The 160 x 81 weights are
// staged once per block in shared memory; a warp takes FC2_R regions at a time so that every weight it reads from shared
// memory feeds FC2_R FMAs (one region per warp made the kernel LSU-bound: 3 LDS + 1 SHFL per 3 FMA). The regions' 160
// inputs sit in shared memory too and are read as broadcast 16-byte loads; each lane owns outputs lane, lane + 32,
// lane + 64 of every region. The label / confidence walk (sequential float32 sum, first maximum) runs on FC2_R lanes at
// once, one region each.
#define FC2_THREADS 256
#define FC2_R 4
#define FC2_WFLOATS ((CNN_F5 * CNN_F6 + CNN_F6 + 3) / 4 * 4)
#define FC2_SMEM ((FC2_WFLOATS + (FC2_THREADS / 32) * FC2_R * CNN_F5) * 4)
static_assert((OFF_W6 % 4) == 0 && (CNN_F5 * CNN_F6) % 4 == 0 && CNN_F5 % 4 == 0 && CNN_F6 <= FC2_THREADS, "16-byte loads of fc2's weights / inputs");
__global__ void __launch_bounds__(FC2_THREADS, 2) cnn_fc2_softmax_label(const float *__restrict__ f5, const float *__restrict__ w,
                                                                     const float *__restrict__ b, int n_regions,
                                                                     float *__restrict__ softmax, int *__restrict__ label,
                                                                     float *__restrict__ conf)
{
    extern __shared__ __align__(16) float s_w[];         // [160][81] weights, [81] biases, then [warp][FC2_R][160] inputs
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const bool has2 = lane + 64 < CNN_F6;
    float *s_x = s_w + FC2_WFLOATS + wib * (FC2_R * CNN_F5);
    const int n_groups = (n_regions + FC2_R - 1) / FC2_R;
    constexpr int XV = FC2_R * CNN_F5 / 4 / 32;          // 16-byte units of a group's inputs per lane
    static_assert(FC2_R * CNN_F5 / 4 % 32 == 0, "a group's inputs split evenly over the lanes");
    // a group's inputs: FC2_R x 160 consecutive floats of f5
    auto load_x = [&](int grp, float4 *xr) {
        const int reg0 = grp * FC2_R, nr = min(FC2_R, n_regions - reg0);
        const float4 *x4 = (const float4 *)(f5 + (size_t)reg0 * CNN_F5);
#pragma unroll
        for (int u = 0; u < XV; u++) {
            const int i = lane + 32 * u;
            xr[u] = (grp < n_groups && i < nr * (CNN_F5 / 4)) ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    const int grp0 = blockIdx.x * (FC2_THREADS / 32) + wib;
    float4 xr[XV];
    load_x(grp0, xr);                                    // in flight together with the weights
    {
        // all of a thread's loads are issued before the first store: a fixed trip count with predicates, not a loop whose
        // remainder runs one L2 round trip per iteration
        constexpr int NW4 = CNN_F5 * CNN_F6 / 4, PER = (NW4 + FC2_THREADS - 1) / FC2_THREADS;
        const float4 *w4 = (const float4 *)w;
        float4 *s4 = (float4 *)s_w;
        float4 wv[PER];
#pragma unroll
        for (int it = 0; it < PER; it++) {
            const int i = threadIdx.x + it * FC2_THREADS;
            wv[it] = i < NW4 ? __ldg(w4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float bv = threadIdx.x < CNN_F6 ? __ldg(b + threadIdx.x) : 0.f;
#pragma unroll
        for (int it = 0; it < PER; it++) {
            const int i = threadIdx.x + it * FC2_THREADS;
            if (i < NW4) s4[i] = wv[it];
        }
        if (threadIdx.x < CNN_F6) s_w[CNN_F5 * CNN_F6 + threadIdx.x] = bv;
    }
    __syncthreads();
    for (int grp = grp0; grp < n_groups; grp += gridDim.x * (FC2_THREADS / 32)) {
        const int reg0 = grp * FC2_R;
        const int nr = min(FC2_R, n_regions - reg0);
        if (grp != grp0) load_x(grp, xr);
        __syncwarp();                                    // the previous group's walk has finished with s_x
#pragma unroll
        for (int u = 0; u < XV; u++) ((float4 *)s_x)[lane + 32 * u] = xr[u];
        __syncwarp();
        float a0[FC2_R], a1[FC2_R], a2[FC2_R];
#pragma unroll
        for (int r = 0; r < FC2_R; r++) a0[r] = a1[r] = a2[r] = 0.f;
#pragma unroll 2
        for (int k4 = 0; k4 < CNN_F5 / 4; k4++) {
            float4 xv[FC2_R];
#pragma unroll
            for (int r = 0; r < FC2_R; r++) xv[r] = *(const float4 *)(s_x + r * CNN_F5 + 4 * k4);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const float *wr = s_w + (4 * k4 + kk) * CNN_F6;
                // lanes without a third output read into the next weight row (inside the array): the value is unused
                const float w0 = wr[lane], w1 = wr[lane + 32], w2 = wr[lane + 64];
#pragma unroll
                for (int r = 0; r < FC2_R; r++) {
                    const float xi = kk == 0 ? xv[r].x : kk == 1 ? xv[r].y : kk == 2 ? xv[r].z : xv[r].w;
                    a0[r] = fmaf(xi, w0, a0[r]);
                    a1[r] = fmaf(xi, w1, a1[r]);
                    a2[r] = fmaf(xi, w2, a2[r]);
                }
            }
        }
        __syncwarp();                                    // all lanes are done with the inputs: s_x now takes the softmax
        const float bias0 = s_w[CNN_F5 * CNN_F6 + lane], bias1 = s_w[CNN_F5 * CNN_F6 + lane + 32];
        const float bias2 = has2 ? s_w[CNN_F5 * CNN_F6 + lane + 64] : 0.f;
#pragma unroll
        for (int r = 0; r < FC2_R; r++) {
            float v[3];
            v[0] = a0[r] + bias0;
            v[1] = a1[r] + bias1;
            v[2] = has2 ? a2[r] + bias2 : -INFINITY;
            float m = fmaxf(fmaxf(v[0], v[1]), v[2]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                v[k] = (lane + 32 * k) < CNN_F6 ? expf(v[k] - m) : 0.f;
                s += v[k];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            float *y = softmax + (size_t)(reg0 + r) * CNN_F6;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                v[k] = __fdiv_rn(v[k], s);
                if (lane + 32 * k < CNN_F6) {
                    s_x[r * CNN_F5 + lane + 32 * k] = v[k];
                    if (r < nr) y[lane + 32 * k] = v[k];
                }
            }
        }
        __syncwarp();
        // label = first maximum (np.argmax); confidence = max(y) / sum(y) with Python's sequential float32 sum
        // (nn_cache.py:28-30): lane r walks the 81 values of region r in order
        if (lane < nr) {
            const float *y = s_x + lane * CNN_F5;
            int best = 0;
            float bv = y[0], tot = 0.f;
#pragma unroll 9
            for (int o = 0; o < CNN_F6; o++) {
                const float t = y[o];
                if (t > bv) { bv = t; best = o; }
                tot = __fadd_rn(tot, t);
            }
            label[reg0 + lane] = best;
            conf[reg0 + lane] = __fdiv_rn(bv, tot);
        }
    }
}

int ckb_cnn_tail_init(ckb_ctx *ctx)   // per context (= per device), called when the weights are installed
{
    CKB_CUDA(ctx, cudaFuncSetAttribute(cnn_fc2_softmax_label, cudaFuncAttributeMaxDynamicSharedMemorySize, FC2_SMEM));
    return CKB_OK;
}

// d_tmp: n*100*(81 + 2) floats (softmax when the caller does not want it, labels, confidences)
int ckb_launch_fc2_decode(ckb_ctx *ctx, const float *d_f5, int n, void *d_tmp, float *d_softmax, uint8_t *d_stones,
                          float *d_conf, uint8_t *d_keep, cudaStream_t st)
{
    const int P = n * 100;
    const float *w = ctx->cnn->d_params;
    float *sm = d_softmax ? d_softmax : (float *)d_tmp;
    int *lab = (int *)((float *)d_tmp + (size_t)P * CNN_F6);
    float *cf = (float *)(lab + P);
    const int per_block = FC2_THREADS / 32 * FC2_R;
    int grid = (P + per_block - 1) / per_block;
    if (grid > 3 * ctx->num_sms) grid = 3 * ctx->num_sms;
    cnn_fc2_softmax_label<<<grid, FC2_THREADS, FC2_SMEM, st>>>(d_f5, w + OFF_W6, w + OFF_B6, P, sm, lab, cf);
    CKB_LAUNCH_CHECK(ctx, "cnn_fc2_softmax_label");
    cnn_decode_board<<<(n * 361 + 127) / 128, 128, 0, st>>>(lab, cf, n, d_stones, d_conf, d_keep);
    CKB_LAUNCH_CHECK(ctx, "cnn_decode_board");
    return CKB_OK;
}

// workspace of the SIMT path, floats per patch
#define SIMT_PER_PATCH (CNN_A1 + CNN_A2 + CNN_P2 + CNN_A3 + CNN_A4 + CNN_P4 + CNN_F5 + CNN_F6 + CNN_F6 + 2)

size_t ckb_cnn_simt_workspace(int n) { return (size_t)n * 100 * SIMT_PER_PATCH * sizeof(float) + 256; }

extern "C" size_t ckb_cnn_workspace_simt(const ckb_ctx *ctx, int n) { return (!ctx || n < 0) ? 0 : ckb_cnn_simt_workspace(n); }

static inline unsigned blocks_for(long long total) { return (unsigned)((total + 255) / 256); }

extern "C" int ckb_cnn_forward_simt(ckb_ctx *ctx, const uint8_t *d_goban, int n, void *d_work, size_t work_bytes,
                                    float *d_softmax, uint8_t *d_stones, float *d_conf, uint8_t *d_keep, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    if (!ctx->cnn) CKB_FAIL(ctx, CKB_E_STATE, "ckb_cnn_forward_simt: call ckb_set_cnn_weights first");
    if (n == 0) return CKB_OK;   // an empty batch is a no-op, whatever the pointers
    if (!d_goban || !d_work || n < 0) CKB_FAIL(ctx, CKB_E_INVALID, "ckb_cnn_forward_simt: bad argument");
    if (work_bytes < ckb_cnn_simt_workspace(n)) CKB_FAIL(ctx, CKB_E_NOMEM, "ckb_cnn_forward_simt: workspace too small");
    if (n == 0) return CKB_OK;
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    CKB_ENTER(ctx, stream);
    cudaStream_t st = (cudaStream_t)stream;
    const int P = n * 100;
    const float *w = ctx->cnn->d_params;
    float *a1 = (float *)d_work;
    float *a2 = a1 + (size_t)P * CNN_A1;
    float *p2 = a2 + (size_t)P * CNN_A2;
    float *a3 = p2 + (size_t)P * CNN_P2;
    float *a4 = a3 + (size_t)P * CNN_A3;
    float *p4 = a4 + (size_t)P * CNN_A4;
    float *f5 = p4 + (size_t)P * CNN_P4;
    float *f6 = f5 + (size_t)P * CNN_F5;
    float *tmp = f6 + (size_t)P * CNN_F6;
    cnn_conv_relu_simt<5, 5, 3, 32, 40, 40, true><<<blocks_for((long long)P * CNN_A1), 256, 0, st>>>(d_goban, w + OFF_W1, w + OFF_B1, a1, P);
    CKB_LAUNCH_CHECK(ctx, "cnn_conv1_simt");
    cnn_conv_relu_simt<5, 5, 32, 32, 36, 36, false><<<blocks_for((long long)P * CNN_A2), 256, 0, st>>>(a1, w + OFF_W2, w + OFF_B2, a2, P);
    CKB_LAUNCH_CHECK(ctx, "cnn_conv2_simt");
    cnn_maxpool2_simt<32, 32, 32><<<blocks_for((long long)P * CNN_P2), 256, 0, st>>>(a2, p2, P);
    CKB_LAUNCH_CHECK(ctx, "cnn_pool2_simt");
    cnn_conv_relu_simt<3, 3, 32, 90, 16, 16, false><<<blocks_for((long long)P * CNN_A3), 256, 0, st>>>(p2, w + OFF_W3, w + OFF_B3, a3, P);
    CKB_LAUNCH_CHECK(ctx, "cnn_conv3_simt");
    cnn_conv_relu_simt<3, 3, 90, 90, 14, 14, false><<<blocks_for((long long)P * CNN_A4), 256, 0, st>>>(a3, w + OFF_W4, w + OFF_B4, a4, P);
    CKB_LAUNCH_CHECK(ctx, "cnn_conv4_simt");
    cnn_maxpool2_simt<12, 12, 90><<<blocks_for((long long)P * CNN_P4), 256, 0, st>>>(a4, p4, P);
    CKB_LAUNCH_CHECK(ctx, "cnn_pool4_simt");
    cnn_dense_simt<3240, 160, true><<<blocks_for((long long)P * CNN_F5), 256, 0, st>>>(p4, w + OFF_W5, w + OFF_B5, f5, P);
    CKB_LAUNCH_CHECK(ctx, "cnn_fc1_simt");
    cnn_dense_simt<160, 81, false><<<blocks_for((long long)P * CNN_F6), 256, 0, st>>>(f5, w + OFF_W6, w + OFF_B6, f6, P);
    CKB_LAUNCH_CHECK(ctx, "cnn_fc2_simt");
    return ckb_launch_decode(ctx, f6, n, d_softmax, tmp, d_stones, d_conf, d_keep, st);
}
