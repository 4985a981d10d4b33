// Shared declarations of the camkifu_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/camkifu_b200.h"

#define CKB_MAX_G 19
#define CKB_MAX_ZONES (CKB_MAX_G * CKB_MAX_G)
#define CKB_JPEG_LANES 8   // independent nvJPEG decoders per context (ckb_jpeg_decode's `lane`)

struct ckb_cnn_weights;  // cnn_pack.cu
struct ckb_jpeg_state;   // jpeg_ingest.cu

struct ckb_prof_entry {
    const char *name;   // NULL = boundary marker (entry of an API call)
    cudaEvent_t ev;
};

struct ckb_ctx {
    int device;
    int gsize;
    int S;  // canonical side = 20 * gsize
    int num_sms;
    char err[512];
    uint64_t launches;
    // constant tables on the device (geometry of stonesfinder.py:412-493)
    int32_t *d_rects;   // [g*g][4] x0,y0,x1,y1
    uint8_t *d_mask;    // [S*S] disk mask
    int32_t h_rects[CKB_MAX_ZONES * 4];
    ckb_cnn_weights *cnn;
    struct ckb_jpeg_state *jpeg[CKB_JPEG_LANES];   // nvJPEG decoders of the Motion-JPEG ingest (jpeg_ingest.cu), created on first use
    // per-kernel timing (ckb_profile_begin / ckb_profile_end): one CUDA event after every launch on the caller's stream
    cudaStream_t cur_stream;
    int prof_on, prof_n, prof_cap;
    ckb_prof_entry *prof;
};

void ckb_prof_mark(ckb_ctx *ctx, const char *name);

// every kernel-launching entry point starts with this: remembers the stream and opens a timing interval
#define CKB_ENTER(ctx, st)                             \
    do {                                               \
        (ctx)->cur_stream = (cudaStream_t)(st);        \
        if ((ctx)->prof_on) ckb_prof_mark(ctx, nullptr); \
    } while (0)

#define CKB_FAIL(ctx, code, ...)                              \
    do {                                                      \
        snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
        return (code);                                        \
    } while (0)

#define CKB_CUDA(ctx, call)                                                                              \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) CKB_FAIL(ctx, CKB_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                                         __FILE__, __LINE__);                                            \
    } while (0)

#define CKB_LAUNCH_CHECK(ctx, name)                                                                       \
    do {                                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                                             \
        if (e__ != cudaSuccess) CKB_FAIL(ctx, CKB_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
        (ctx)->launches++;                                                                                \
        if ((ctx)->prof_on) ckb_prof_mark(ctx, name);                                                     \
    } while (0)

// geometry (host side, geometry.cu)
void ckb_host_zone_rects(int gsize, int32_t *rects);
void ckb_host_zone_mask(int gsize, const int32_t *rects, uint8_t *mask);

// kernel launchers (one per .cu)
int ckb_launch_warp(ckb_ctx *ctx, const uint8_t *d_frames, int n, int H, int W, size_t row_pitch, size_t frame_pitch,
                    const double *h_minv, int n_mtx, uint8_t *d_goban, cudaStream_t st);
int ckb_launch_accumulate(ckb_ctx *ctx, const uint8_t *d_goban, int n, float *d_accu, float alpha, int first,
                          float *d_snap, int snap_every, int snap_phase, cudaStream_t st);
