/*
 * camkifu_b200.h — C ABI of the B200-native (sm_100a) stone-detection hot path of CamKifu.
 *
 * The reference (ArnaudPel/CamKifu) has NO native interface: its plugin boundary is the Python class contract of
 * camkifu.stone.StonesFinder (src/camkifu/stone/stonesfinder.py:18) and the arithmetic is done by cv2 / Keras calls.
 * Each entry point below replaces one of those third-party calls (or the Python loop around it); the reference call
 * site is cited on every function. The Python mirror of the plugin API that sits on top is camkifu_b200/plugins.py;
 * INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success, a negative CKB_E_* code on failure and never throws or aborts;
 *     ckb_last_error(ctx) gives the message of the last failure on that context;
 *   - the caller owns every device buffer (raw device pointers) and the CUDA stream (cudaStream_t passed as void*);
 *     the library owns only the opaque context (constant tables, packed CNN weights);
 *   - all work is enqueued on `stream` and is asynchronous with respect to the host; no function synchronises;
 *   - one context per finder instance / thread; contexts share no mutable state. Every kernel-launching entry point
 *     makes the context's device current on the calling thread (cudaSetDevice) and leaves it current: a thread that
 *     drives contexts on several GPUs must not rely on the current device across calls;
 *   - images are BGR uint8, interleaved, row-major (what cv2.VideoCapture hands the reference, vmanager.py:584);
 *   - board colours are uint8 codes CKB_E / CKB_B / CKB_W (the reference's Golib constants 'E','B','W').
 */
#ifndef CAMKIFU_B200_H
#define CAMKIFU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CKB_VERSION 100 /* 0.1.0 */

#define CKB_E 0
#define CKB_B 1
#define CKB_W 2

#define CKB_OK 0
#define CKB_E_INVALID (-1)  /* bad argument */
#define CKB_E_CUDA (-2)     /* a CUDA runtime call failed */
#define CKB_E_STATE (-3)    /* e.g. CNN weights not set */
#define CKB_E_NOMEM (-4)    /* workspace too small */

#define CKB_CNN_NPARAM 658665 /* NNManager.create_net, nn_manager.py:277-298 */
#define CKB_CNN_PATCHES 100   /* NNManager.split ** 2, nn_manager.py:46 */
#define CKB_CNN_CLASSES 81    /* 3 ** 4, nn_manager.py:48 */

typedef struct ckb_ctx ckb_ctx;

int ckb_version(void);

/* gsize: 9, 13 or 19 (golib_conf.gsize; canonical image side = 20 * gsize, cvconf.py:10). The CNN entry points need 19. */
int ckb_create(ckb_ctx **out, int device, int gsize);
int ckb_destroy(ckb_ctx *ctx);
const char *ckb_last_error(const ckb_ctx *ctx);

/* cv::RNG helpers (cv2.kmeans draws from the process-global cv::theRNG(); here the caller owns the state).
 * ckb_rng_seed(s) is the state cv2.setRNGSeed(s) installs; one cv2.kmeans(K=3, attempts=3, PP) call consumes
 * CKB_RNG_DRAWS_PER_KMEANS 32-bit draws whatever the data. */
#define CKB_RNG_DRAWS_PER_KMEANS 39
uint64_t ckb_rng_seed(uint32_t seed);
uint64_t ckb_rng_advance(uint64_t state, uint64_t n_draws);

/* Geometry tables of the canonical image (host helpers; the kernels hold the same tables on the device).
 * Replaces: StonesFinder.getrect(r, c)  stonesfinder.py:412-450  -> rects4: gsize*gsize x {x0, y0, x1, y1}, row-major (x = row)
 *           StonesFinder.getmask()      stonesfinder.py:452-493  -> mask: (20 gsize)^2 uint8, 1 inside each zone's disk */
int ckb_zone_rects(int gsize, int32_t *rects4);
int ckb_zone_mask(int gsize, uint8_t *mask);

/* 3x3 inverse exactly as cv::invert computes it for the matrix warpPerspective receives (host helper). */
int ckb_invert_homography(const double *m9, double *minv9);

/* ---- K1 -----------------------------------------------------------------------------------------------------------
 * Replaces: cv2.warpPerspective(frame, transform, self.canonical_shape)          stonesfinder.py:140
 * n frames of H x W x 3 uint8 (row_pitch / frame_pitch in bytes) -> n canonical images of S x S x 3 uint8, densely
 * packed, S = 20 * gsize. h_mtx: n_mtx (1 or n) row-major 3x3 float64 frame->canonical homographies, i.e. exactly the
 * `BoardFinder.mtx` the reference passes (boardfinder.py:43-45); HOST memory, consumed before the call returns.
 * Result is bit-identical to OpenCV's INTER_LINEAR / BORDER_CONSTANT(0) fixed-point remap. */
int ckb_warp(ckb_ctx *ctx, const uint8_t *d_frames, int n, int H, int W, size_t row_pitch, size_t frame_pitch,
             const double *h_mtx, int n_mtx, uint8_t *d_goban, void *stream);

/* Replaces: self.accu = gframe.astype(np.float32) / cv2.accumulateWeighted(gframe, self.accu, 0.2)
 *                                                                                  sf_clustering.py:33-36
 * Applies the n canonical images IN ORDER to the running average d_accu (S*S*3 float32): frame 0 initialises it when
 * first != 0, otherwise accu = fma(alpha, src - accu, accu). If d_snapshots != NULL, the state after frame i is also
 * stored to d_snapshots[j] for the j-th frame with (i + snap_phase) % snap_every == 0 (the reference runs detection
 * when total_f_processed % 3 == 0, sf_clustering.py:37). */
int ckb_accumulate(ckb_ctx *ctx, const uint8_t *d_goban, int n, float *d_accu, float alpha, int first,
                   float *d_snapshots, int snap_every, int snap_phase, void *stream);

/* ---- background model ----------------------------------------------------------------------------
 * Replaces: self.bg_model = cv2.createBackgroundSubtractorMOG2(detectShadows=False)        stonesfinder.py:113-115
 *           self._fg = self.bg_model.apply(self.goban_img, learningRate=learning)          stonesfinder.py:171-176
 * d_state: ckb_mog2_state_bytes(ctx) bytes of DEVICE memory owned by the caller, one per finder (the per-pixel Gaussian
 * mixture: 5 modes x (weight, variance, BGR mean) + mode count); ckb_mog2_reset = a freshly created subtractor.
 * ckb_mog2_apply feeds n canonical images IN ORDER; frames_before = how many frames the model has already seen (OpenCV's
 * nframes: the very first frame always learns at rate 1/2), h_learning_rates = the n `learningRate` arguments (HOST;
 * negative = OpenCV's automatic 1/min(2 nframes, 500)). d_fgmask: n x S x S uint8, 0 or 255, bit-identical to cv2's.
 * ckb_zone_fg_counts: per zone rectangle (getrect) the number of foreground pixels = np.sum(fg[a0:a1, b0:b1]) / 255, the
 * quantity SfNeural.is_agitated thresholds (sf_neural.py:178-180); d_counts n x gsize^2 int32. */
size_t ckb_mog2_state_bytes(const ckb_ctx *ctx);
int ckb_mog2_reset(ckb_ctx *ctx, void *d_state, void *stream);
int ckb_mog2_apply(ckb_ctx *ctx, const uint8_t *d_goban, int n, void *d_state, long long frames_before,
                   const double *h_learning_rates, uint8_t *d_fgmask, void *stream);
int ckb_zone_fg_counts(ckb_ctx *ctx, const uint8_t *d_fgmask, int n, int32_t *d_counts, void *stream);

/* ---- K3 + K2 ------------------------------------------------------------------------------------------------------
 * Replaces: SfClustering.find_stones(img, rs, re, cs, ce)                          sf_clustering.py:48-178
 *   = cv2.kmeans(pixels, 3, None, (TERM_CRITERIA_EPS, 15, 3), 3, KMEANS_PP_CENTERS)   :103-104
 *   + masked per-zone label histogram -> ratios                                     :105-129
 *   + interpret_ratios / check_density                                              :131-178
 * d_imgs: n canonical images, uint8 (is_f32 == 0) or float32 (is_f32 != 0, e.g. accu snapshots), S x S x 3 dense.
 * d_rng_states: n cv::RNG states, one per image (DEVICE memory) — the state cv::theRNG() would hold when the
 *   reference reaches that image's cv2.kmeans call.
 * d_work: scratch of at least ckb_find_stones_workspace(ctx, n) bytes.
 * Outputs (DEVICE; any optional pointer may be NULL):
 *   d_stones  n x g x g uint8 codes (E outside the region)       d_trusted n uint8 (0 = the reference returns None)
 *   d_ratios  n x g x g x 3 uint8    d_centers n x 9 float32     d_compactness n float64
 *   d_labels  n x (x1-x0)*(y1-y0) int32 k-means labels of the region's bounding box, as cv2.kmeans returns them */
size_t ckb_find_stones_workspace(const ckb_ctx *ctx, int n);
int ckb_find_stones(ckb_ctx *ctx, const void *d_imgs, int is_f32, int n, int rs, int re, int cs, int ce,
                    const uint64_t *d_rng_states, void *d_work, size_t work_bytes, uint8_t *d_stones,
                    uint8_t *d_trusted, uint8_t *d_ratios, float *d_centers, double *d_compactness, int32_t *d_labels,
                    void *stream);

/* Batched form for SfMeta, whose 3 x 3 Regions each call cluster.find_stones(img, rs=, re=, cs=, ce=) on the same image
 * (sf_meta.py:245-262 try_clustering, :232-243 routine): n_regions (<= 16) regions x n uint8 images in one set of launches.
 * h_regions4: n_regions x {rs, re, cs, ce} (HOST). d_rng_states: n x n_regions states (image-major), the state of the
 * call (image, region). Outputs are indexed the same way: d_stones n x n_regions x g*g (E outside each region), d_trusted
 * n x n_regions, d_ratios n x n_regions x g*g*3, d_centers n x n_regions x 9, d_compactness n x n_regions. */
size_t ckb_find_stones_regions_workspace(const ckb_ctx *ctx, int n, int n_regions);
int ckb_find_stones_regions(ckb_ctx *ctx, const uint8_t *d_imgs, int n, int n_regions, const int *h_regions4,
                            const uint64_t *d_rng_states, void *d_work, size_t work_bytes, uint8_t *d_stones,
                            uint8_t *d_trusted, uint8_t *d_ratios, float *d_centers, double *d_compactness, void *stream);

/* ---- per-zone statistics of SfMeta / SfContours ------------------------------------------------------
 * ckb_zone_means replaces the zone loop of SfContours.find_stones and _norm_channels   sf_contours.py:87-102,113-126
 *   d_imgs n canonical images; d_masks n x S x S uint8, non-zero where the reference's `mask` (filled convex hulls of the
 *   contours, drawn on the host: Canny / findContours / convexHull stay on the CPU) is 1, in canonical image coordinates.
 *   d_zones n x (re-rs) x (ce-cs) x 4 int16: [0] = 1 if more than 40 % of the zone's pixels are under the mask, [1..3] =
 *   int16(sum of the B, G, R values of the visible pixels / their number) if so, else the same over the masked-out pixels.
 * ckb_history_vote replaces the per-intersection vote of Region.commit                  sf_meta.py:305-340
 *   d_history n_items x histo uint8 colour codes (the CyclicBuffer entries of an intersection, any order), d_is_empty
 *   n_items uint8 (StonesFinder.is_empty); d_moves n_items uint8: 0 = nothing to submit, CKB_B / CKB_W = the move. */
int ckb_zone_means(ckb_ctx *ctx, const uint8_t *d_imgs, const uint8_t *d_masks, int n, int rs, int re, int cs, int ce,
                   int16_t *d_zones, void *stream);
int ckb_history_vote(ckb_ctx *ctx, const uint8_t *d_history, const uint8_t *d_is_empty, int n_items, int histo,
                     uint8_t *d_moves, void *stream);

/* ---- K4 -----------------------------------------------------------------------------------------------------------
 * Replaces: NNManager.get_net() / create_net() weights                              nn_manager.py:58-74,277-298
 * h_params: CKB_CNN_NPARAM float32 in Keras channels-last order (w1 b1 w2 b2 w3 b3 w4 b4 w5 b5 w6 b6, conv kernels
 * (kh,kw,cin,cout), dense (in,out)); HOST memory. Packs them into the tensor-core operand layout on the device
 * (synchronous; call once per model). */
int ckb_set_cnn_weights(ckb_ctx *ctx, const float *h_params, size_t n_params);

/* Replaces: NNCache.predict_all_stones() = 100 x (NNManager._get_x + net.predict) + decode   nn_cache.py:25-52
 *           and the MIN_CONFIDENCE rule of SfNeural.predict_all                      sf_neural.py:57-70
 * d_goban: n canonical 380 x 380 x 3 uint8 images. d_work: at least ckb_cnn_workspace(ctx, n) bytes.
 * Outputs (DEVICE; optional ones may be NULL):
 *   d_softmax n x 100 x 81 float32 (region (i, j) at index i*10+j)
 *   d_stones  n x 361 uint8 codes, d_conf n x 361 float32 (max(y)/sum(y) of the region that wrote the intersection
 *             last, i.e. region 9 overwrites region 8 on row / column 17), d_keep n x 361 uint8 (stone != E and
 *             conf > 0.6) */
size_t ckb_cnn_workspace(const ckb_ctx *ctx, int n);
int ckb_cnn_forward(ckb_ctx *ctx, const uint8_t *d_goban, int n, void *d_work, size_t work_bytes, float *d_softmax,
                    uint8_t *d_stones, float *d_conf, uint8_t *d_keep, void *stream);

/* Verification aids (tests only). ckb_cnn_forward_simt: the same forward pass on plain fp32 CUDA cores, no tensor cores;
 * same arguments, but d_work must hold ckb_cnn_workspace_simt(ctx, n) bytes. ckb_cnn_debug_activation: after
 * ckb_cnn_forward on n <= 64 frames, unpacks one intermediate activation of the tensor-core path from its workspace into
 * dense float32 [patch][H][W][C]: layer 1 = conv1 (36,36,32), 2 = pooled conv2 (16,16,32), 3 = conv3 (14,14,90),
 * 4 = pooled conv4 (6,6,90), 5 = fc1 (160). conv1's activations normally never leave shared memory (gather + conv1 + conv2 + pool are one kernel):
 * ckb_cnn_set_debug(ctx, 1) makes the following forward passes also write them to the workspace (which grows:
 * query ckb_cnn_workspace again) so that layer 1 can be inspected. */
int ckb_cnn_set_debug(ckb_ctx *ctx, int on);
size_t ckb_cnn_workspace_simt(const ckb_ctx *ctx, int n);
int ckb_cnn_debug_activation(ckb_ctx *ctx, const void *d_work, int n, int layer, float *d_out, void *stream);
int ckb_cnn_forward_simt(ckb_ctx *ctx, const uint8_t *d_goban, int n, void *d_work, size_t work_bytes,
                         float *d_softmax, uint8_t *d_stones, float *d_conf, uint8_t *d_keep, void *stream);

/* ---- host <-> device staging ----------------------------------------------------------------------------------------
 * ckb_frame_roi: the part of an H x W frame that ckb_warp can read for homography m9 (frame -> canonical S x S):
 * roi4 = {y0, y1, x0, x1}, half-open (host helper; the whole frame when the homography gives no bound).
 * ckb_upload_frames: asynchronous copy of that region of n HOST frames (pinned memory for real overlap) into the device
 * frame buffer ckb_warp reads; roi4 == NULL copies whole frames. Replaces nothing in the reference (which has no
 * device); it is the H2D leg of StonesFinder._doframe's `frame` argument (stonesfinder.py:123). */
int ckb_frame_roi(const double *m9, int H, int W, int S, int *roi4);
int ckb_upload_frames(ckb_ctx *ctx, const uint8_t *h_frames, int n, int H, int W, size_t h_row_pitch,
                      size_t h_frame_pitch, const int *roi4, uint8_t *d_frames, size_t d_row_pitch,
                      size_t d_frame_pitch, void *stream);

/* ---- Motion-JPEG ingest on the device (optional) ------------------------------------------------------
 * Replaces, for MJPG files: CaptureReader.read_file -> cv2.VideoCapture.read                 vmanager.py:563-586
 * h_jpeg / h_sizes: n compressed frames in HOST memory (consumed before the call returns); every frame must decode to
 * H x W. d_frames: n x H x W x 3 uint8 BGR interleaved on the device (pitches in bytes), decoded by nvJPEG (a CUDA-toolkit
 * library: the hardware JPEG engines when available, else its CUDA decoder) on `stream`. cpu_threads: host threads the
 * library may use for its CPU stages. The pixels may differ from FFmpeg's decode of the same frame by a level or two
 * (different IDCT / upsampling), which is why this is an option of the batch API, not the default ingest.
 * lane (0..7): which of the context's independent decoders to use. One batch keeps the GPU busy for a fraction of its
 * decode time only (a chain of small kernels), so calls on different lanes, from different host threads on different
 * streams, overlap: two lanes give 1.7 x one. A lane must not be used by two threads at once.
 * ckb_jpeg_backend: which nvJPEG backend the context got. */
int ckb_jpeg_decode(ckb_ctx *ctx, const uint8_t *const *h_jpeg, const size_t *h_sizes, int n, int H, int W, uint8_t *d_frames,
                    size_t row_pitch, size_t frame_pitch, int cpu_threads, int lane, void *stream);
const char *ckb_jpeg_backend(ckb_ctx *ctx);

/* ---- per-kernel timing (bench.py's roofline) ---------------------------------------------------------------------------
 * Between ckb_profile_begin and ckb_profile_end the library records a CUDA event on the caller's stream after every
 * kernel it launches (up to `capacity` events). ckb_profile_end waits for the last one and returns, per launch in issue
 * order, the kernel's name (32-byte slots) and its device time in ms. */
int ckb_profile_begin(ckb_ctx *ctx, int capacity);
int ckb_profile_end(ckb_ctx *ctx, int max_entries, char *names, float *ms, int *n_out);

/* Number of kernels the library has launched on this context since creation (bench.py's gpu_launches). */
uint64_t ckb_launch_count(const ckb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* CAMKIFU_B200_H */
