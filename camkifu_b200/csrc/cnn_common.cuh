// SfNeural CNN (NNManager.create_net, src/camkifu/stone/nn_manager.py:277-298): shared shapes and device weight store.
#pragma once
#include "ckb_common.cuh"

// layer geometry: 40 -conv5-> 36 -conv5-> 32 -pool-> 16 -conv3-> 14 -conv3-> 12 -pool-> 6 ; 3240 -> 160 -> 81
#define CNN_IN 40
#define CNN_C1 32
#define CNN_C2 32
#define CNN_C3 90
#define CNN_C4 90
#define CNN_F5 160
#define CNN_F6 81
#define CNN_A1 (36 * 36 * CNN_C1)
#define CNN_A2 (32 * 32 * CNN_C2)
#define CNN_P2 (16 * 16 * CNN_C2)
#define CNN_A3 (14 * 14 * CNN_C3)
#define CNN_A4 (12 * 12 * CNN_C4)
#define CNN_P4 (6 * 6 * CNN_C4)

// offsets into the flat Keras-order parameter blob
#define OFF_W1 0
#define OFF_B1 (OFF_W1 + 5 * 5 * 3 * 32)
#define OFF_W2 (OFF_B1 + 32)
#define OFF_B2 (OFF_W2 + 5 * 5 * 32 * 32)
#define OFF_W3 (OFF_B2 + 32)
#define OFF_B3 (OFF_W3 + 3 * 3 * 32 * 90)
#define OFF_W4 (OFF_B3 + 90)
#define OFF_B4 (OFF_W4 + 3 * 3 * 90 * 90)
#define OFF_W5 (OFF_B4 + 90)
#define OFF_B5 (OFF_W5 + 3240 * 160)
#define OFF_W6 (OFF_B5 + 160)
#define OFF_B6 (OFF_W6 + 160 * 81)
static_assert(OFF_B6 + 81 == CKB_CNN_NPARAM, "parameter count");

struct ckb_cnn_weights {
    float *d_params;   // the flat fp32 blob (biases, fc2 and the SIMT verification path read it)
    void *d_tc;        // tensor-core operand planes (cnn_tc.cu)
    size_t tc_bytes;
    int dump_a1;       // test aid: the front kernel also writes conv1 activations to the workspace
};

// patch origin of region (i, j): NNManager._get_rect_nn(*_subregion(i, j)) (nn_manager.py:92-126,256-275): 40 i, except
// the last region which is shifted back to end at 380 (340).
__host__ __device__ __forceinline__ int cnn_patch_origin(int i) { return i < 9 ? 40 * i : 340; }

int ckb_cnn_tc_pack(ckb_ctx *ctx, const float *h_params);   // cnn_tc.cu
void ckb_cnn_tc_free(ckb_ctx *ctx);
int ckb_cnn_front_init(ckb_ctx *ctx);                       // cnn_tc_front.cu
int ckb_cnn_tail_init(ckb_ctx *ctx);                        // cnn_simt.cu
int ckb_launch_cnn_front(ckb_ctx *ctx, const uint8_t *d_goban, int n_patches, const void *w1, const float *b1, const void *w2,
                         const float *b2, void *p2, long long p2_plane, void *dbg_a1, long long a1_plane, cudaStream_t st);
int ckb_launch_decode(ckb_ctx *ctx, const float *d_logits, int n, float *d_softmax_or_null, float *d_softmax_tmp,
                      uint8_t *d_stones, float *d_conf, uint8_t *d_keep, cudaStream_t st);
