// step.cu -- ntf_fnn_step: one whole Fnn batch (fnn.py:118-151) enqueued by ONE call.
//
// The reference's step is ~40 ATen launches driven from Python with a host sync per batch (loss.item(), fnn.py:140).  Here a step
// is ~20 launches of this library; driven one by one through ctypes the host needs ~10 us per launch and becomes the bottleneck
// once the kernels are fast, so the sequence lives here: the host mirror (opentf_b200/engine.py) fills one struct per step.
// No kernels in this file -- it only calls the entry points of include/ntf_b200.h in the order of fnn.py's loop body.
#include "common.cuh"

static size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

extern "C" size_t ntf_fnn_step_workspace_bytes(const ntf_ctx* ctx, const ntf_fnn_step_args* a) {
  if (!a || a->n_layers < 2 || a->n_layers > NTF_MAX_LAYERS) return 0;
  const int L = a->n_layers, h_last = a->hidden[L - 2];
  size_t w = 0;
  w = max_sz(w, ntf_expert_cdf_workspace_bytes(a->E_total > 0 ? a->E_total : a->E));
  w = max_sz(w, ntf_out_train_workspace_bytes(ctx, a->precision, a->B, h_last, a->E, 0));
  w = max_sz(w, ntf_csr_bag_bwd_workspace_bytes(a->S, a->hidden[0]));
  for (int i = 0; i < L - 1; ++i) w = max_sz(w, ntf_act_bwd_workspace_bytes(a->B, a->hidden[i]));
  for (int i = 1; i < L - 1; ++i) w = max_sz(w, ntf_dense_bwd_workspace_bytes(a->B, a->hidden[i - 1], a->hidden[i]));
  return w;
}

extern "C" int ntf_fnn_step(ntf_ctx* ctx, void* stream, const ntf_fnn_step_args* a, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ctx && a, NTF_ERR_BAD_ARG, "fnn_step: null ctx/args");
  NTF_REQUIRE(a->n_layers >= 2 && a->n_layers <= NTF_MAX_LAYERS, NTF_ERR_UNSUPPORTED, "fnn_step: %d layers (2..%d)", a->n_layers, NTF_MAX_LAYERS);
  NTF_REQUIRE(a->B > 0 && a->S > 0 && a->E > 0, NTF_ERR_BAD_ARG, "fnn_step: B=%d S=%d E=%d", a->B, a->S, a->E);
  NTF_REQUIRE(workspace_bytes >= ntf_fnn_step_workspace_bytes(ctx, a), NTF_ERR_WORKSPACE, "fnn_step: workspace too small");
  const int L = a->n_layers, Lo = L - 1, B = a->B;
  const int* h = a->hidden;
  const int Etot = a->E_total > 0 ? a->E_total : a->E;  // the sampler works on the global expert axis
  const int phase = a->phase ? a->phase : 3;
  NTF_REQUIRE(a->e_lo >= 0 && a->e_lo + a->E <= Etot, NTF_ERR_BAD_ARG, "fnn_step: expert range [%d,%d) of %d", a->e_lo, a->e_lo + a->E, Etot);
  int rc;
#define STEP(call) do { if ((rc = (call)) != NTF_OK) return rc; } while (0)
  if (phase & 1) {
  // ---- forward through the hidden layers: fnn.py:25 (layer 0 = CSR bag, ntf.py:23 never densified) ----
  STEP(ntf_csr_bag_fwd(ctx, stream, B, a->s_indptr, a->s_indices, a->W[0], a->b[0], a->S, h[0], a->act[0]));
  for (int i = 1; i < Lo; ++i) STEP(ntf_dense_fwd(ctx, stream, a->act[i - 1], a->W[i], a->b[i], B, h[i - 1], h[i], 1, a->act[i]));
  // ---- negative sampling: fnn.py:34-36,48-76 (unless the caller supplied the indices) ----
  const int32_t* neg = nullptr;
  int ns = 0;
  if (a->neg_given) { neg = a->neg; ns = a->ns; }
  else if (a->nsd != NTF_NS_NONE) {
    if (a->nsd == NTF_NS_UNIGRAM_B) {
      const int gB = a->gB > 0 ? a->gB : B;
      STEP(ntf_expert_cdf(ctx, stream, gB, a->gB > 0 ? a->g_m_indptr : a->m_indptr, a->m_indices, Etot, a->counts, a->cdf, workspace, workspace_bytes));
    }
    STEP(ntf_neg_sample(ctx, stream, a->nsd, a->seed, a->step, a->row0, B, a->m_indptr, a->m_indices, Etot, a->ns,
                        a->nsd == NTF_NS_UNIFORM ? nullptr : a->cdf, a->neg));
    neg = a->neg; ns = a->ns;
  }
  // ---- output layer: forward + weighted BCE (+ backward): fnn.py:32-46,135,137 ----
  const bool tc = a->precision == NTF_TF32;
  ntf_out_train_args o;
  memset(&o, 0, sizeof(o));
  if (tc) {
    STEP(ntf_special_tiles(ctx, stream, 1, B, a->m_indptr, a->m_indices, neg, ns, a->E, a->e_lo, a->special_t, a->member_t));
    o.special_t = a->special_t; o.member_t = a->member_t;
  } else {
    STEP(ntf_special_bits(ctx, stream, 1, B, a->m_indptr, a->m_indices, neg, ns, a->E, a->e_lo, a->special, a->pitch_words));
    o.special = a->special; o.pitch_words = a->pitch_words;
  }
  o.A = a->act[Lo - 1]; o.W = a->W[Lo]; o.b = a->b[Lo];
  o.m_indptr = a->m_indptr; o.m_indices = a->m_indices;
  o.B = B; o.h = h[Lo - 1]; o.E = a->E; o.e_lo = a->e_lo;
  o.tpw = a->tpw; o.tnw = a->tnw; o.loss_scale = a->loss_scale; o.loss_out = a->loss_out;
  if (a->train) { o.dW = a->gW[Lo]; o.db = a->gb[Lo]; o.dA = a->dact[Lo - 1]; }
  if (a->prof_ev[0]) NTF_CUDA(cudaEventRecord((cudaEvent_t)a->prof_ev[0], as_stream(stream)));
  STEP(ntf_out_train(ctx, stream, a->precision, &o, workspace, workspace_bytes));
  if (a->prof_ev[1]) NTF_CUDA(cudaEventRecord((cudaEvent_t)a->prof_ev[1], as_stream(stream)));
  if (!tc) STEP(ntf_special_bits(ctx, stream, 0, B, a->m_indptr, a->m_indices, neg, ns, a->E, a->e_lo, a->special, a->pitch_words));
  }
  if (!a->train || !(phase & 2)) return NTF_OK;
  // ---- backward through the hidden layers ----
  for (int i = Lo - 1; i > 0; --i) {
    STEP(ntf_act_bwd(ctx, stream, a->dact[i], a->act[i], B, h[i], 1, a->dz[i], a->gb[i], workspace, workspace_bytes));
    STEP(ntf_dense_bwd(ctx, stream, a->act[i - 1], a->W[i], a->dz[i], B, h[i - 1], h[i], a->gW[i], a->dact[i - 1], workspace, workspace_bytes));
  }
  STEP(ntf_act_bwd(ctx, stream, a->dact[0], a->act[0], B, h[0], 1, a->dz[0], a->gb[0], workspace, workspace_bytes));
  STEP(ntf_csr_bag_bwd(ctx, stream, B, a->s_indptr, a->s_indices, a->s_ent_row, a->row_base, a->dz[0], a->S, h[0], a->gW[0], workspace, workspace_bytes));
  // ---- optimiser: fnn.py:139 (skipped when the caller all-reduces the gradients first: data-parallel ranks) ----
  if (a->run_adam) STEP(ntf_adam_step(ctx, stream, a->params, a->grads, a->adam_m, a->adam_v, a->n_params, a->lr, a->beta1, a->beta2, a->eps, a->adam_t));
#undef STEP
  return NTF_OK;
}
