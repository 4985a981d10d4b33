// K4 — tensor-core (tcgen05) forward pass of the SfNeural CNN.  (placeholder: filled in next)
#include "cnn_common.cuh"

size_t ckb_cnn_simt_workspace(int n);

int ckb_cnn_tc_pack(ckb_ctx *ctx, const float *h_params)
{
    (void)ctx; (void)h_params;
    return CKB_OK;
}

void ckb_cnn_tc_free(ckb_ctx *ctx)
{
    if (ctx->cnn && ctx->cnn->d_tc) { cudaFree(ctx->cnn->d_tc); ctx->cnn->d_tc = nullptr; }
}

extern "C" size_t ckb_cnn_workspace(const ckb_ctx *ctx, int n)
{
    if (!ctx || n < 0) return 0;
    return ckb_cnn_simt_workspace(n);
}

extern "C" int ckb_cnn_forward(ckb_ctx *ctx, const uint8_t *d_goban, int n, void *d_work, size_t work_bytes,
                               float *d_softmax, uint8_t *d_stones, float *d_conf, uint8_t *d_keep, void *stream)
{
    if (!ctx) return CKB_E_INVALID;
    CKB_FAIL(ctx, CKB_E_STATE, "ckb_cnn_forward: tensor-core path not built yet");
}
