// out_tc.cu -- NTF_TF32 mode: tcgen05 / TMEM / TMA kernels of the output layer (placeholder until the kernel lands).
#include "common.cuh"

size_t ntf_out_train_tc_workspace_bytes(const ntf_ctx*, int, int, int, int) { return 0; }
int ntf_out_tc_supported(int, int, int, int) { return 0; }
int ntf_out_train_tc(ntf_ctx*, cudaStream_t, const ntf_out_train_args*, void*, size_t) {
  ntf_set_error("out_train(tf32): not built");
  return NTF_ERR_UNSUPPORTED;
}
int ntf_infer_scores_tc(ntf_ctx*, cudaStream_t, const float*, const float*, const float*, int, int, int, float*) {
  ntf_set_error("infer_scores(tf32): not built");
  return NTF_ERR_UNSUPPORTED;
}
