// out_layer.cu -- C-ABI entry points of the output layer (training step + inference scores); dispatches on precision
// to the CUDA-core fp32 path (dense_simt.cu) or the tcgen05/TMEM path (out_tc.cu).
#include <stdlib.h>

#include "common.cuh"

size_t ntf_out_train_fp32_workspace_bytes(int B, int h, int E, int flipout);
int ntf_out_train_fp32(ntf_ctx* ctx, cudaStream_t st, const ntf_out_train_args* a, void* workspace, size_t workspace_bytes);
int ntf_infer_scores_fp32(cudaStream_t st, const float* A, const float* W, const float* b, int B, int h, int E, const float* A_s,
                          const float* W_delta, const float* b_delta, const uint32_t* sign_out, int pitch_words, int accumulate,
                          float* P, float* T_ws);
size_t ntf_out_train_tc_workspace_bytes(const ntf_ctx* ctx, int B, int h, int E, int flipout);
int ntf_out_train_tc(ntf_ctx* ctx, cudaStream_t st, const ntf_out_train_args* a, void* workspace, size_t workspace_bytes);
int ntf_out_tc_supported(int B, int h, int E, int flipout);
size_t ntf_out_train_tc2_workspace_bytes(int B, int h, int E);
int ntf_out_train_tc2(ntf_ctx* ctx, cudaStream_t st, const ntf_out_train_args* a, void* workspace, size_t workspace_bytes);
int ntf_infer_scores_tc(ntf_ctx* ctx, cudaStream_t st, const float* A, const float* W, const float* b, int B, int h, int E, float* P,
                        void* workspace, size_t workspace_bytes);
size_t ntf_infer_scores_tc_workspace_bytes(int B, int h);
size_t ntf_infer_scores_tc_flip_workspace_bytes(int B, int h, int E);
int ntf_infer_scores_tc_flip(ntf_ctx* ctx, cudaStream_t st, const float* A, const float* W, const float* b, int B, int h, int E, const float* A_s,
                             const float* W_delta, const float* b_delta, const uint32_t* sign_out, int pitch_words, float* P, void* workspace,
                             size_t workspace_bytes);

extern "C" int ntf_tc_supported(int B, int h, int E, int flipout) { return ntf_out_tc_supported(B, h, E, flipout); }

extern "C" size_t ntf_out_train_workspace_bytes(const ntf_ctx* ctx, int precision, int B, int h, int E, int flipout) {
  if (precision == NTF_TF32) {
    const size_t v1 = ntf_out_train_tc_workspace_bytes(ctx, B, h, E, flipout), v2 = flipout ? 0 : ntf_out_train_tc2_workspace_bytes(B, h, E);
    return v1 > v2 ? v1 : v2;
  }
  return ntf_out_train_fp32_workspace_bytes(B, h, E, flipout);
}

extern "C" int ntf_out_train(ntf_ctx* ctx, void* stream, int precision, const ntf_out_train_args* a, void* workspace,
                             size_t workspace_bytes) {
  NTF_REQUIRE(ctx && a, NTF_ERR_BAD_ARG, "out_train: null ctx/args");
  NTF_REQUIRE(a->A && a->W && a->b && a->m_indptr && a->m_indices && a->loss_out, NTF_ERR_BAD_ARG, "out_train: null pointer");
  NTF_REQUIRE(a->B > 0 && a->h > 0 && a->E > 0, NTF_ERR_BAD_ARG, "out_train: B=%d h=%d E=%d", a->B, a->h, a->E);
  NTF_REQUIRE(!a->special || a->pitch_words * 32 >= a->E, NTF_ERR_BAD_ARG, "out_train: pitch_words=%d too small for E=%d", a->pitch_words, a->E);
  const bool train = a->dW != nullptr;
  NTF_REQUIRE(!train || (a->db && (a->dA || true)), NTF_ERR_BAD_ARG, "out_train: dW given without db");
  const bool flip = a->W_delta != nullptr;
  if (flip) {
    NTF_REQUIRE(a->A_s && a->b_delta && a->sign_out, NTF_ERR_BAD_ARG, "out_train: Flipout needs A_s, b_delta, sign_out");
    NTF_REQUIRE(!train || (a->dW_delta && a->db_delta), NTF_ERR_BAD_ARG, "out_train: Flipout training needs dW_delta, db_delta");
    NTF_REQUIRE(a->pitch_words * 32 >= a->E, NTF_ERR_BAD_ARG, "out_train: sign_out pitch too small");
  }
  NTF_REQUIRE(workspace || workspace_bytes == 0, NTF_ERR_BAD_ARG, "out_train: null workspace");
  if (precision == NTF_TF32) {
    NTF_REQUIRE(ntf_out_tc_supported(a->B, a->h, a->E, flip), NTF_ERR_UNSUPPORTED,
                "out_train(tf32): shape B=%d h=%d E=%d flipout=%d not supported by the tcgen05 kernel (use NTF_FP32)", a->B, a->h, a->E, (int)flip);
    // Fnn: the persistent kernel (out_tc2.cu); the Flipout layer and NTF_TC_V1=1 (round 1's one-CTA-per-expert-tile kernel, kept for A/B runs): out_tc.cu
    if (!flip && !a->special_t && getenv("NTF_TC_V1") == nullptr) return ntf_out_train_tc2(ctx, as_stream(stream), a, workspace, workspace_bytes);
    return ntf_out_train_tc(ctx, as_stream(stream), a, workspace, workspace_bytes);
  }
  NTF_REQUIRE(precision == NTF_FP32, NTF_ERR_BAD_ARG, "out_train: precision=%d", precision);
  return ntf_out_train_fp32(ctx, as_stream(stream), a, workspace, workspace_bytes);
}

extern "C" size_t ntf_infer_scores_workspace_bytes(int B, int E, int flipout) {
  // Flipout: the perturbation term T[B,E] (fp32 kernels) or the fp16 tiles of the tensor-core path, whichever is larger; otherwise the
  // fp16 copy of the activations for the tensor-core path (h <= 128 there)
  if (!flipout) return ntf_infer_scores_tc_workspace_bytes(B, 128);
  const size_t a = align_up((size_t)B * E * sizeof(float), 256), t = ntf_infer_scores_tc_flip_workspace_bytes(B, 128, E);
  return a > t ? a : t;
}

extern "C" int ntf_infer_scores(ntf_ctx* ctx, void* stream, int precision, const float* A, const float* W, const float* b, int B,
                                int h, int E, const float* A_s, const float* W_delta, const float* b_delta,
                                const uint32_t* sign_out, int pitch_words, int accumulate, float* P, void* workspace,
                                size_t workspace_bytes) {
  NTF_REQUIRE(ctx && A && W && b && P, NTF_ERR_BAD_ARG, "infer_scores: null pointer");
  NTF_REQUIRE(B > 0 && h > 0 && E > 0, NTF_ERR_BAD_ARG, "infer_scores: B=%d h=%d E=%d", B, h, E);
  const bool flip = W_delta != nullptr;
  if (flip) {
    NTF_REQUIRE(A_s && b_delta && sign_out && pitch_words * 32 >= E, NTF_ERR_BAD_ARG, "infer_scores: incomplete Flipout arguments");
    NTF_REQUIRE(workspace && workspace_bytes >= ntf_infer_scores_workspace_bytes(B, E, 1), NTF_ERR_WORKSPACE, "infer_scores: workspace too small");
  }
  if (precision == NTF_TF32 && !flip && !accumulate && ntf_out_tc_supported(B, h, E, 0))
    return ntf_infer_scores_tc(ctx, as_stream(stream), A, W, b, B, h, E, P, workspace, workspace_bytes);
  if (precision == NTF_TF32 && flip && !accumulate && ntf_out_tc_supported(B, h, E, 1))
    return ntf_infer_scores_tc_flip(ctx, as_stream(stream), A, W, b, B, h, E, A_s, W_delta, b_delta, sign_out, pitch_words, P, workspace, workspace_bytes);
  NTF_REQUIRE(precision == NTF_FP32 || precision == NTF_TF32, NTF_ERR_BAD_ARG, "infer_scores: precision=%d", precision);
  return ntf_infer_scores_fp32(as_stream(stream), A, W, b, B, h, E, A_s, W_delta, b_delta, sign_out, pitch_words, accumulate, P,
                               (float*)workspace);
}

namespace {
__global__ void axpy_kernel(size_t n, float a, const float* __restrict__ x, float* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] += a * x[i];
}
}  // namespace

extern "C" int ntf_axpy(ntf_ctx* ctx, void* stream, size_t n, float a, const float* x, float* y) {
  NTF_REQUIRE(ctx && x && y, NTF_ERR_BAD_ARG, "axpy: null pointer");
  if (!n) return NTF_OK;
  size_t b = (n + 255) / 256, cap = (size_t)ctx->sm_count * 16;
  NTF_COUNT_LAUNCH; axpy_kernel<<<(unsigned)(b < cap ? b : cap), 256, 0, as_stream(stream)>>>(n, a, x, y);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
