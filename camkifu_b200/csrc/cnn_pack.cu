// CNN weight store: ckb_set_cnn_weights keeps the flat fp32 blob on the device and asks cnn_tc.cu to build the
// tensor-core operand planes. Replaces NNManager.get_net()/create_net() as the weight source (nn_manager.py:58-74).
#include <new>

#include "cnn_common.cuh"

void ckb_cnn_free(ckb_ctx *ctx)
{
    if (!ctx->cnn) return;
    ckb_cnn_tc_free(ctx);
    if (ctx->cnn->d_params) cudaFree(ctx->cnn->d_params);
    delete ctx->cnn;
    ctx->cnn = nullptr;
}

extern "C" int ckb_set_cnn_weights(ckb_ctx *ctx, const float *h_params, size_t n_params)
{
    if (!ctx) return CKB_E_INVALID;
    if (!h_params || n_params != CKB_CNN_NPARAM)
        CKB_FAIL(ctx, CKB_E_INVALID, "ckb_set_cnn_weights: expected %d parameters", CKB_CNN_NPARAM);
    if (ctx->gsize != 19) CKB_FAIL(ctx, CKB_E_STATE, "the SfNeural network is defined for 19x19 only (nn_manager.py:281,295)");
    for (size_t i = 0; i < n_params; i++)
        if (!(h_params[i] == h_params[i]) || h_params[i] > 3.0e38f || h_params[i] < -3.0e38f)
            CKB_FAIL(ctx, CKB_E_INVALID, "ckb_set_cnn_weights: parameter %zu is not finite", i);
    CKB_CUDA(ctx, cudaSetDevice(ctx->device));
    ckb_cnn_free(ctx);
    ctx->cnn = new (std::nothrow) ckb_cnn_weights();
    if (!ctx->cnn) CKB_FAIL(ctx, CKB_E_NOMEM, "host allocation failed");
    memset(ctx->cnn, 0, sizeof(*ctx->cnn));
    CKB_CUDA(ctx, cudaMalloc(&ctx->cnn->d_params, sizeof(float) * CKB_CNN_NPARAM));
    CKB_CUDA(ctx, cudaMemcpy(ctx->cnn->d_params, h_params, sizeof(float) * CKB_CNN_NPARAM, cudaMemcpyHostToDevice));
    return ckb_cnn_tc_pack(ctx, h_params);
}
