// Motion-JPEG ingest on the GPU ("NVDEC / batched decode").
//
// Replaces, for MJPG sources, CaptureReader.read_file -> cv2.VideoCapture.read (src/camkifu/core/vmanager.py:563-586):
// the reference decodes one frame per processed frame on the finder's thread; at GPU speeds host decode (0.9 k frames/s
// on 16 cores) and the PCIe upload of decoded frames (6.2 MB each at 1080p) are what bounds "fast video file processing"
// (README.md:35). Here the COMPRESSED frames cross PCIe (~0.3 MB each) and a batch is decoded on the device by nvJPEG — a
// CUDA-toolkit library, used the way cuBLAS would be: it is not one of this repo's kernels — straight into the frame
// buffer ckb_warp reads (interleaved BGR). NVDEC itself is not reachable from this image (no header).
//
// The decoder is a different JPEG implementation than FFmpeg's (IDCT and chroma upsampling round differently by a level
// or two), so this path is an OPTION of the batch API (process_video(ingest="nvjpeg")), not the default: the default
// ingest hands the detection path the very pixels cv2.VideoCapture gives the reference.
#include <dlfcn.h>
#include <stdlib.h>
#include <nvjpeg.h>

#include "ckb_common.cuh"

// nvJPEG is bound at first use (dlopen), not at link time: a box without libnvjpeg must still load this library; only
// the two entry points below then fail, with a message.
struct NvjpegApi {
    void *lib;
    decltype(&nvjpegCreateEx) CreateEx;
    decltype(&nvjpegCreateSimple) CreateSimple;
    decltype(&nvjpegDestroy) Destroy;
    decltype(&nvjpegJpegStateCreate) JpegStateCreate;
    decltype(&nvjpegJpegStateDestroy) JpegStateDestroy;
    decltype(&nvjpegGetImageInfo) GetImageInfo;
    decltype(&nvjpegDecodeBatchedInitialize) DecodeBatchedInitialize;
    decltype(&nvjpegDecodeBatched) DecodeBatched;
    decltype(&nvjpegDecodeBatchedPreAllocate) DecodeBatchedPreAllocate;   // optional
};

static bool nvjpeg_load(NvjpegApi &api)
{
    const char *names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"};
    api.lib = nullptr;
    for (const char *n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (api.lib) break;
    }
    if (!api.lib) return false;
#define CKB_SYM(field, name) api.field = (decltype(api.field))dlsym(api.lib, name)
    CKB_SYM(CreateEx, "nvjpegCreateEx");
    CKB_SYM(CreateSimple, "nvjpegCreateSimple");
    CKB_SYM(Destroy, "nvjpegDestroy");
    CKB_SYM(JpegStateCreate, "nvjpegJpegStateCreate");
    CKB_SYM(JpegStateDestroy, "nvjpegJpegStateDestroy");
    CKB_SYM(GetImageInfo, "nvjpegGetImageInfo");
    CKB_SYM(DecodeBatchedInitialize, "nvjpegDecodeBatchedInitialize");
    CKB_SYM(DecodeBatched, "nvjpegDecodeBatched");
    CKB_SYM(DecodeBatchedPreAllocate, "nvjpegDecodeBatchedPreAllocate");
#undef CKB_SYM
    return api.CreateEx && api.CreateSimple && api.Destroy && api.JpegStateCreate && api.JpegStateDestroy &&
           api.GetImageInfo && api.DecodeBatchedInitialize && api.DecodeBatched;
}

static NvjpegApi *nvjpeg_api()
{
    static NvjpegApi api;
    static const bool ok = nvjpeg_load(api);   // once, thread-safe: the lanes may make their first call concurrently
    return ok ? &api : nullptr;
}

struct ckb_jpeg_state {
    nvjpegHandle_t handle;
    nvjpegJpegState_t state;
    int batch;              // batch size of the last nvjpegDecodeBatchedInitialize
    int backend;            // nvjpegBackend_t in use
    nvjpegImage_t *dst;     // [batch] destinations
    int dst_cap;
};

static const char *jpeg_status(nvjpegStatus_t s)
{
    switch (s) {
    case NVJPEG_STATUS_SUCCESS: return "success";
    case NVJPEG_STATUS_NOT_INITIALIZED: return "not initialized";
    case NVJPEG_STATUS_INVALID_PARAMETER: return "invalid parameter";
    case NVJPEG_STATUS_BAD_JPEG: return "bad jpeg";
    case NVJPEG_STATUS_JPEG_NOT_SUPPORTED: return "jpeg not supported";
    case NVJPEG_STATUS_ALLOCATOR_FAILURE: return "allocator failure";
    case NVJPEG_STATUS_EXECUTION_FAILED: return "execution failed";
    case NVJPEG_STATUS_ARCH_MISMATCH: return "arch mismatch";
    case NVJPEG_STATUS_INTERNAL_ERROR: return "internal error";
    case NVJPEG_STATUS_IMPLEMENTATION_NOT_SUPPORTED: return "implementation not supported";
    case NVJPEG_STATUS_INCOMPLETE_BITSTREAM: return "incomplete bitstream";
    default: return "unknown";
    }
}

#define CKB_JPEG(ctx, call)                                                                             \
    do {                                                                                                \
        nvjpegStatus_t s__ = (call);                                                                    \
        if (s__ != NVJPEG_STATUS_SUCCESS)                                                               \
            CKB_FAIL(ctx, CKB_E_CUDA, "%s: nvjpeg %s (%s:%d)", #call, jpeg_status(s__), __FILE__, __LINE__); \
    } while (0)

void ckb_jpeg_free(ckb_ctx *ctx)
{
    NvjpegApi *nj = nvjpeg_api();
    for (int l = 0; l < CKB_JPEG_LANES; l++) {
        ckb_jpeg_state *j = ctx->jpeg[l];
        if (!j) continue;
        if (nj && j->state) nj->JpegStateDestroy(j->state);
        if (nj && j->handle) nj->Destroy(j->handle);
        delete[] j->dst;
        delete j;
        ctx->jpeg[l] = nullptr;
    }
}

static int jpeg_init(ckb_ctx *ctx, int lane)
{
    if (ctx->jpeg[lane]) return CKB_OK;
    NvjpegApi *nj = nvjpeg_api();
    if (!nj) CKB_FAIL(ctx, CKB_E_STATE, "libnvjpeg is not available on this machine (dlopen failed)");
    ckb_jpeg_state *j = new ckb_jpeg_state();
    memset(j, 0, sizeof(*j));
    // the hardware JPEG engines when the library offers them on this GPU, else the CUDA decoder
    // (CKB_JPEG_BACKEND = 0 / 1 / 2: force nvJPEG's default / hybrid / GPU-hybrid backend, for measurements)
    const char *forced = getenv("CKB_JPEG_BACKEND");
    j->backend = NVJPEG_BACKEND_HARDWARE;
    if (forced && forced[0] >= '0' && forced[0] <= '2' &&
        nj->CreateEx((nvjpegBackend_t)(forced[0] - '0'), nullptr, nullptr, 0, &j->handle) == NVJPEG_STATUS_SUCCESS) {
        j->backend = forced[0] - '0';
    } else if (nj->CreateEx(NVJPEG_BACKEND_HARDWARE, nullptr, nullptr, 0, &j->handle) != NVJPEG_STATUS_SUCCESS) {
        j->backend = NVJPEG_BACKEND_GPU_HYBRID;
        if (nj->CreateEx(NVJPEG_BACKEND_GPU_HYBRID, nullptr, nullptr, 0, &j->handle) != NVJPEG_STATUS_SUCCESS) {
            j->backend = NVJPEG_BACKEND_DEFAULT;
            nvjpegStatus_t s = nj->CreateSimple(&j->handle);
            if (s != NVJPEG_STATUS_SUCCESS) {
                delete j;
                CKB_FAIL(ctx, CKB_E_CUDA, "nvjpegCreate: %s", jpeg_status(s));
            }
        }
    }
    nvjpegStatus_t s = nj->JpegStateCreate(j->handle, &j->state);
    if (s != NVJPEG_STATUS_SUCCESS) {
        nj->Destroy(j->handle);
        delete j;
        CKB_FAIL(ctx, CKB_E_CUDA, "nvjpegJpegStateCreate: %s", jpeg_status(s));
    }
    ctx->jpeg[lane] = j;
    return CKB_OK;
}

extern "C" const char *ckb_jpeg_backend(ckb_ctx *ctx)
{
    if (!ctx || jpeg_init(ctx, 0) != CKB_OK) return "unavailable";
    switch (ctx->jpeg[0]->backend) {
    case NVJPEG_BACKEND_HARDWARE: return "nvjpeg hardware engine";
    case NVJPEG_BACKEND_GPU_HYBRID: return "nvjpeg GPU hybrid (CUDA Huffman + IDCT)";
    case NVJPEG_BACKEND_HYBRID: return "nvjpeg hybrid (CPU Huffman, CUDA IDCT)";
    default: return "nvjpeg default (hybrid)";
    }
}

extern "C" int ckb_jpeg_decode(ckb_ctx *ctx, const uint8_t *const *h_jpeg, const size_t *h_sizes, int n, int H, int W,
                               uint8_t *d_frames, size_t row_pitch, size_t frame_pitch, int cpu_threads, int lane, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    if (n == 0) return CKB_OK;
    if (!h_jpeg || !h_sizes || !d_frames || n < 0 || H < 1 || W < 1 || lane < 0 || lane >= CKB_JPEG_LANES)
        CKB_FAIL(ctx, CKB_E_INVALID, "ckb_jpeg_decode: bad argument");
    if (row_pitch < (size_t)W * 3 || (n > 1 && frame_pitch < row_pitch * (size_t)H))
        CKB_FAIL(ctx, CKB_E_INVALID, "ckb_jpeg_decode: pitches smaller than the image");
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int rc = jpeg_init(ctx, lane);
    if (rc != CKB_OK) return rc;
    ckb_jpeg_state *j = ctx->jpeg[lane];
    NvjpegApi *nj = nvjpeg_api();
    // every frame must be the H x W the destination was sized for
    nvjpegChromaSubsampling_t sub0 = NVJPEG_CSS_420;
    for (int i = 0; i < n; i++) {
        int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
        nvjpegChromaSubsampling_t sub;
        CKB_JPEG(ctx, nj->GetImageInfo(j->handle, h_jpeg[i], h_sizes[i], &nc, &sub, ws, hs));
        if (ws[0] != W || hs[0] != H)
            CKB_FAIL(ctx, CKB_E_INVALID, "ckb_jpeg_decode: frame %d is %d x %d, expected %d x %d", i, ws[0], hs[0], W, H);
        if (i == 0) sub0 = sub;
    }
    if (j->batch != n) {
        CKB_JPEG(ctx, nj->DecodeBatchedInitialize(j->handle, j->state, n, cpu_threads > 0 ? cpu_threads : 1, NVJPEG_OUTPUT_BGRI));
        // size the decoder's internal buffers once: without this every call allocates and frees device memory
        if (nj->DecodeBatchedPreAllocate) nj->DecodeBatchedPreAllocate(j->handle, j->state, n, W, H, sub0, NVJPEG_OUTPUT_BGRI);
        j->batch = n;
    }
    if (j->dst_cap < n) {
        delete[] j->dst;
        j->dst = new nvjpegImage_t[n];
        j->dst_cap = n;
    }
    for (int i = 0; i < n; i++) {
        memset(&j->dst[i], 0, sizeof(nvjpegImage_t));
        j->dst[i].channel[0] = d_frames + (size_t)i * frame_pitch;
        j->dst[i].pitch[0] = row_pitch;
    }
    CKB_JPEG(ctx, nj->DecodeBatched(j->handle, j->state, h_jpeg, h_sizes, j->dst, (cudaStream_t)stream));
    return CKB_OK;
}
