// step.cu -- ntf_fnn_step: one whole Fnn batch (fnn.py:118-151) enqueued by ONE call.
//
// The reference's step is ~40 ATen launches driven from Python with a host sync per batch (loss.item(), fnn.py:140).  Here a step
// is ~12 launches of this library; driven one by one through ctypes the host needs ~10 us per launch and becomes the bottleneck
// once the kernels are fast, so the sequence lives here: the host mirror (opentf_b200/engine.py) fills one struct per step, and
// replays the whole call as a CUDA graph from the second epoch on (ntf_graph_*).
// The parts of a step that depend on the batch's CSR only -- negative sampling + the condition planes, and the slot fill of the
// input layer's backward pass -- run on the handle's side streams next to the input layer / the output layer (parallel branches
// of the captured graph) and are joined where their results are consumed.
// No kernels in this file -- it only calls the entry points of include/ntf_b200.h in the order of fnn.py's loop body.
#include <stdlib.h>

#include "common.cuh"

int ntf_neg_sample_impl(ntf_ctx* ctx, void* stream, int nsd, uint64_t seed, uint64_t step, int row0, int B, const int32_t* m_indptr,
                        const int32_t* m_indices, int E, int ns, const uint32_t* cdf, const int32_t* pool_indptr, int pool_rows,
                        int32_t* neg, const ntf_dyn* dyn);
int ntf_adam_step_impl(ntf_ctx* ctx, cudaStream_t st, float* p, const float* g, float* m, float* v, size_t n, double lr, double beta1,
                       double beta2, double eps, int64_t step, const ntf_dyn* dyn, void* shadow, size_t sh_off, size_t sh_n);
int ntf_csr_bag_fwd_impl(ntf_ctx* ctx, void* stream, int B, const int32_t* indptr, const int32_t* indices, const float* W0T,
                         const float* b0, int S, int h, float* A, void* A16, bool chained);
int ntf_csr_bag_bwd_fill_impl(ntf_ctx* ctx, cudaStream_t st, int B, const int32_t* indptr, const int32_t* indices, const int32_t* ent_row,
                              int row_base, int S, int h, void* workspace, size_t workspace_bytes, const uint32_t* ent_sign);
int ntf_csr_bag_bwd_reduce_impl(ntf_ctx* ctx, cudaStream_t st, int B, const int32_t* indptr, const int32_t* indices, const int32_t* ent_row,
                                int row_base, const float* dZ, int S, int h, float* dW0T, void* workspace, size_t workspace_bytes,
                                const uint32_t* ent_sign, cudaStream_t st_hot);

int ntf_peer_exchange_adam_impl(ntf_ctx* ctx, cudaStream_t st, const ntf_peers* pr, float* adam_m, float* adam_v, size_t off, size_t n, double lr,
                                double beta1, double beta2, double eps, int64_t step, const ntf_dyn* dyn, int channel);

int ntf_peer_allreduce_impl(ntf_ctx* ctx, cudaStream_t st, const ntf_peers* pr, size_t n, float* dst, int channel);

static size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

// workspace = [ input-layer backward (slots, lives from the fork to the end of the step) | everything that runs on the main stream ]
static size_t bag_region(const ntf_fnn_step_args* a) { return a->x_dense ? 0 : align_up(ntf_csr_bag_bwd_workspace_bytes(a->S, a->hidden[0]), 256); }

extern "C" size_t ntf_fnn_step_workspace_bytes(const ntf_ctx* ctx, const ntf_fnn_step_args* a) {
  if (!a || a->n_layers < 2 || a->n_layers > NTF_MAX_LAYERS) return 0;
  const int L = a->n_layers, h_last = a->hidden[L - 2];
  size_t w = 0;
  w = max_sz(w, ntf_out_train_workspace_bytes(ctx, a->precision, a->B, h_last, a->E, 0));
  for (int i = 0; i < L - 1; ++i) w = max_sz(w, ntf_act_bwd_workspace_bytes(a->B, a->hidden[i]));
  for (int i = 1; i < L - 1; ++i) w = max_sz(w, ntf_dense_bwd_workspace_bytes(a->B, a->hidden[i - 1], a->hidden[i]));
  if (a->x_dense) w = max_sz(w, ntf_dense_bwd_workspace_bytes(a->B, a->S, a->hidden[0]));
  return bag_region(a) + w;
}

extern "C" int ntf_fnn_step(ntf_ctx* ctx, void* stream, const ntf_fnn_step_args* a, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ctx && a, NTF_ERR_BAD_ARG, "fnn_step: null ctx/args");
  NTF_REQUIRE(a->n_layers >= 2 && a->n_layers <= NTF_MAX_LAYERS, NTF_ERR_UNSUPPORTED, "fnn_step: %d layers (2..%d)", a->n_layers, NTF_MAX_LAYERS);
  NTF_REQUIRE(a->B > 0 && a->S > 0 && a->E > 0, NTF_ERR_BAD_ARG, "fnn_step: B=%d S=%d E=%d", a->B, a->S, a->E);
  NTF_REQUIRE(workspace && workspace_bytes >= ntf_fnn_step_workspace_bytes(ctx, a), NTF_ERR_WORKSPACE, "fnn_step: workspace too small");
  const int L = a->n_layers, Lo = L - 1, B = a->B;
  const int* h = a->hidden;
  const int Etot = a->E_total > 0 ? a->E_total : a->E;  // the sampler works on the global expert axis
  const int phase = a->phase ? a->phase : 3;
  NTF_REQUIRE(a->e_lo >= 0 && a->e_lo + a->E <= Etot, NTF_ERR_BAD_ARG, "fnn_step: expert range [%d,%d) of %d", a->e_lo, a->e_lo + a->E, Etot);
  cudaStream_t st = as_stream(stream);
  void* ws_bag = workspace;
  const size_t ws_bag_bytes = bag_region(a);
  void* ws_main = (char*)workspace + ws_bag_bytes;
  const size_t ws_main_bytes = workspace_bytes - ws_bag_bytes;
  const bool bwd_here = a->train && (phase & 2);
  // tensor-core Fnn output layer = the persistent kernel + sparse correction pass of out_tc2.cu (NTF_TC_V1=1: round 1's kernel, bit planes)
  const bool tc2 = a->precision == NTF_TF32 && getenv("NTF_TC_V1") == nullptr;
  const bool shard_x = a->peers != nullptr && a->E < Etot;  // expert shard with a peer table: dA is exchanged inside the step
  float* dA_out = shard_x ? a->peers->grads[a->peers->rank] : a->dact[Lo - 1];
  // one hidden layer, nothing between the output layer's dA and layer 0's activation derivative: the correction pass finishes it (no ntf_act_bwd)
  const bool act_fused = tc2 && Lo == 1 && a->train && phase == 3 && !shard_x;
  ntf_out_train_args o;  // the output layer's call; its final reductions may run later, off the critical path (finish_pending)
  memset(&o, 0, sizeof(o));
  bool finish_pending = false, finish_side = false;
  int rc;
#define STEP(call) do { if ((rc = (call)) != NTF_OK) return rc; } while (0)
  // ---- fork: what needs the batch's CSR only ----
  if ((phase & 1) || bwd_here) NTF_CUDA(cudaEventRecord(ctx->ev_fork, st));
  const bool prep = bwd_here && tc2 && (phase & 1);
  if (bwd_here && (!a->x_dense || prep)) {
    NTF_CUDA(cudaStreamWaitEvent(ctx->side[1], ctx->ev_fork, 0));
    // side 1: what the output layer's kernel needs cleared (dA, the dW / db rows two CTAs share) -- off the critical path
    if (prep) STEP(ntf_out_train_prepare(ctx, (void*)ctx->side[1], B, h[Lo - 1], a->E, a->gW[Lo], a->gb[Lo], dA_out));
    // ... and every (team, skill) entry takes a slot of its skill (pass 1 of the input layer's backward)
    if (!a->x_dense) STEP(ntf_csr_bag_bwd_fill_impl(ctx, ctx->side[1], B, a->s_indptr, a->s_indices, a->s_ent_row, a->row_base, a->S, h[0], ws_bag, ws_bag_bytes, nullptr));
    NTF_CUDA(cudaEventRecord(ctx->ev_join[1], ctx->side[1]));
  }
  if (phase & 1) {
    const bool tc = a->precision == NTF_TF32;
    // ---- side 0: negative sampling fnn.py:34-36,48-76 (unless the caller supplied the indices) + the condition planes fnn.py:33-43 ----
    const int32_t* neg = nullptr;
    int ns = 0;
    NTF_CUDA(cudaStreamWaitEvent(ctx->side[0], ctx->ev_fork, 0));
    if (a->neg_given) { neg = a->neg; ns = a->ns; }
    else if (a->nsd != NTF_NS_NONE) {
      // unigram_b (fnn.py:74-76): the pool of a draw is the member CSR of the GLOBAL batch this call is a slice of
      STEP(ntf_neg_sample_impl(ctx, ctx->side[0], a->nsd, a->seed, a->step, a->row0, B, a->m_indptr, a->m_indices, Etot, a->ns,
                               a->nsd == NTF_NS_UNIGRAM ? a->cdf : nullptr, a->gB > 0 ? a->g_m_indptr : a->m_indptr, a->gB > 0 ? a->gB : B,
                               a->neg, a->dyn));
      neg = a->neg; ns = a->ns;
    }
    if (tc2) {
      o.neg = neg; o.ns = ns;  // persistent kernel + sparse correction pass (out_tc2.cu): the pairs of weight tpw are named by the lists themselves
      o.prepared = prep ? 1 : 0;
      if (act_fused) { o.act_prev = a->act[0]; o.dz_prev = a->dz[0]; o.db_prev = a->gb[0]; }
      // loss_out / db_prev gate nothing but the optimiser: their reductions run next to the input layer's backward pass (side stream).  Only in
      // the one-hidden-layer CSR configuration: there nothing else touches the main part of the workspace while they are pending
      o.defer_finish = (bwd_here && a->run_adam && act_fused && !a->x_dense) ? 1 : 0;
      finish_pending = o.defer_finish != 0;
    } else if (tc) {
      STEP(ntf_special_tiles(ctx, ctx->side[0], 1, B, a->m_indptr, a->m_indices, neg, ns, a->E, a->e_lo, a->special_t, a->member_t));
      o.special_t = a->special_t; o.member_t = a->member_t;
    } else {
      STEP(ntf_special_bits(ctx, ctx->side[0], 1, B, a->m_indptr, a->m_indices, neg, ns, a->E, a->e_lo, a->special, a->pitch_words));
      o.special = a->special; o.pitch_words = a->pitch_words;
    }
    NTF_CUDA(cudaEventRecord(ctx->ev_join[0], ctx->side[0]));
    // ---- main: forward through the hidden layers: fnn.py:25 (layer 0 = CSR bag, ntf.py:23 never densified) ----
    // (one hidden layer + tensor-core output layer: the bag kernel also writes the fp16 operand copy the output layer reads)
    void* A16 = (tc && Lo == 1 && (h[0] % 8) == 0 && !a->x_dense) ? ws_main : nullptr;
    if (a->x_dense) STEP(ntf_dense_fwd(ctx, stream, a->x_dense, a->W[0], a->b[0], B, a->S, h[0], 1, a->act[0]));  // ntf.py:24: embedded skills
    else STEP(ntf_csr_bag_fwd_impl(ctx, stream, B, a->s_indptr, a->s_indices, a->W[0], a->b[0], a->S, h[0], a->act[0], A16, false));
    for (int i = 1; i < Lo; ++i) STEP(ntf_dense_fwd(ctx, stream, a->act[i - 1], a->W[i], a->b[i], B, h[i - 1], h[i], 1, a->act[i]));
    // ---- output layer: forward + weighted BCE (+ backward): fnn.py:32-46,135,137 ----
    o.A = a->act[Lo - 1]; o.W = a->W[Lo]; o.b = a->b[Lo]; o.A16 = A16; o.W16 = a->W16;
    o.m_indptr = a->m_indptr; o.m_indices = a->m_indices;
    o.B = B; o.h = h[Lo - 1]; o.E = a->E; o.e_lo = a->e_lo;
    o.tpw = a->tpw; o.tnw = a->tnw; o.loss_scale = a->loss_scale; o.loss_out = a->loss_out;
    // (expert-sharded layer with a peer table: this shard's dA goes to its peer-visible exchange block, the sum over shards into dact below)
    if (a->train) { o.dW = a->gW[Lo]; o.db = a->gb[Lo]; o.dA = dA_out; }
    NTF_CUDA(cudaStreamWaitEvent(st, ctx->ev_join[0], 0));
    if (prep) NTF_CUDA(cudaStreamWaitEvent(st, ctx->ev_join[1], 0));
    // (inside a stream capture the pair becomes two external event-record nodes, re-recorded by every replay of the graph)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (a->prof_ev[0]) NTF_CUDA(cudaStreamIsCapturing(st, &cap));
    const unsigned evf = cap == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault;
    if (a->prof_ev[0]) NTF_CUDA(cudaEventRecordWithFlags((cudaEvent_t)a->prof_ev[0], st, evf));
    STEP(ntf_out_train(ctx, stream, a->precision, &o, ws_main, ws_main_bytes));
    if (a->prof_ev[1]) NTF_CUDA(cudaEventRecordWithFlags((cudaEvent_t)a->prof_ev[1], st, evf));
    if (!tc) STEP(ntf_special_bits(ctx, stream, 0, B, a->m_indptr, a->m_indices, neg, ns, a->E, a->e_lo, a->special, a->pitch_words));
    if (shard_x && a->train) {
      NTF_REQUIRE((((size_t)B * h[Lo - 1]) % 4) == 0, NTF_ERR_UNSUPPORTED, "fnn_step: sharded exchange needs B*h a multiple of 4");
      STEP(ntf_peer_allreduce_impl(ctx, st, a->peers, (size_t)B * h[Lo - 1], a->dact[Lo - 1], 0));
    }
  }
  if (!bwd_here) return NTF_OK;
  // ---- optimiser, part 1: the output layer's segment of the arena (its gradients are final once ntf_out_train has run) is
  // stepped on a side stream NEXT TO the backward pass through the hidden layers (a chain of small launches that leaves most
  // of the machine idle); the rest of the arena follows on the main stream.  Only when this call ran the output layer too.
  // the fp16 image of the last layer's weight (a->W16, read by the tensor-core output layer) follows the parameters: the Adam launch that
  // steps that weight rewrites it; after a peer-memory exchange (other ranks stepped most of it) a local conversion pass does
  const size_t w_off = (size_t)(a->W[Lo] - a->params), w_n = (size_t)a->E * h[Lo - 1];
  const bool shadow = a->W16 != nullptr && a->run_adam && a->W[Lo] >= a->params && w_off + w_n <= a->n_params && (w_off % 4) == 0 && (w_n % 4) == 0;
  NTF_REQUIRE(a->W16 == nullptr || !a->run_adam || shadow, NTF_ERR_BAD_ARG, "fnn_step: W16 given but the last layer's weight is not a 4-float aligned part of the arena");
  size_t opt_split = a->n_params;  // floats [opt_split, n_params) belong to the last layer (arena order: layer by layer)
  if (a->run_adam && phase == 3 && getenv("NTF_ADAM_SPLIT") == nullptr) {
    const float* lo = a->gW[Lo] < a->gb[Lo] ? a->gW[Lo] : a->gb[Lo];
    const size_t off = (size_t)(lo - a->grads);
    bool last = lo >= a->grads && off < a->n_params && (off % 32) == 0;
    for (int i = 0; i < Lo && last; ++i) last = a->gW[i] < lo && a->gb[i] < lo;
    if (last) opt_split = off;
  }
  // data-parallel ranks: sum the gradients over the ranks before they are stepped (ncclAllReduce, in place, fp32 sum) -- on the
  // communication stream, so that the exchange of the output layer's segment hides behind the hidden layers' backward pass and
  // its Adam step behind the exchange of the rest
  const bool peers = a->peers != nullptr && !(a->E < Etot);  // exchange + Adam fused over peer memory (peer.cu); takes precedence over comm
                                                             // (an expert-sharded layer's table is the dA exchange above: its Adam is local)
  NTF_REQUIRE(!peers || (a->run_adam && phase == 3 && a->peers->params[a->peers->rank] == a->params && a->peers->grads[a->peers->rank] == a->grads),
              NTF_ERR_BAD_ARG, "fnn_step: peers needs run_adam, phase 3 and this rank's own arenas in the table");
  const bool dp = a->comm != nullptr && !peers;
  NTF_REQUIRE(!dp || (a->allreduce && a->run_adam && phase == 3), NTF_ERR_BAD_ARG, "fnn_step: comm needs allreduce, run_adam and phase 3");
  const ntf_allreduce_fn allreduce = (ntf_allreduce_fn)a->allreduce;
#define ALLREDUCE(ptr, count)                                                                                               \
  do {                                                                                                                      \
    const int nr = allreduce((ptr), (ptr), (count), /*ncclFloat32*/ 7, /*ncclSum*/ 0, a->comm, (void*)ctx->comm_st);        \
    NTF_REQUIRE(nr == 0, NTF_ERR_CUDA, "fnn_step: ncclAllReduce failed with ncclResult %d", nr);                          \
  } while (0)
  if (opt_split < a->n_params) {
    NTF_CUDA(cudaEventRecord(ctx->ev_fork_opt, st));
    if (dp) {
      NTF_CUDA(cudaStreamWaitEvent(ctx->comm_st, ctx->ev_fork_opt, 0));
      ALLREDUCE(a->grads + opt_split, a->n_params - opt_split);
      NTF_CUDA(cudaEventRecord(ctx->ev_ar[0], ctx->comm_st));
      NTF_CUDA(cudaStreamWaitEvent(ctx->side[0], ctx->ev_ar[0], 0));
    } else
    NTF_CUDA(cudaStreamWaitEvent(ctx->side[0], ctx->ev_fork_opt, 0));
    if (finish_pending) {
      STEP(ntf_out_train_finish(ctx, (void*)ctx->side[0], &o, ws_main, ws_main_bytes));
      NTF_CUDA(cudaEventRecord(ctx->ev_finish, ctx->side[0]));
      finish_pending = false; finish_side = true;
    }
    const bool sh_here = shadow && w_off >= opt_split;
    if (peers) {
      STEP(ntf_peer_exchange_adam_impl(ctx, ctx->side[0], a->peers, a->adam_m, a->adam_v, opt_split, a->n_params - opt_split, a->lr, a->beta1,
                                       a->beta2, a->eps, a->adam_t, a->dyn, 1));
      if (sh_here) STEP(ntf_to_half(ctx, (void*)ctx->side[0], a->W[Lo], w_n, const_cast<void*>(a->W16)));
    } else
    STEP(ntf_adam_step_impl(ctx, ctx->side[0], a->params + opt_split, a->grads + opt_split, a->adam_m + opt_split, a->adam_v + opt_split,
                            a->n_params - opt_split, a->lr, a->beta1, a->beta2, a->eps, a->adam_t, a->dyn, sh_here ? const_cast<void*>(a->W16) : nullptr,
                            sh_here ? w_off - opt_split : 0, sh_here ? w_n : 0));
    NTF_CUDA(cudaEventRecord(ctx->ev_join_opt, ctx->side[0]));
  }
  if (finish_pending) { STEP(ntf_out_train_finish(ctx, stream, &o, ws_main, ws_main_bytes)); finish_pending = false; }  // (no side stream this step)
  // ---- backward through the hidden layers ----
  for (int i = Lo - 1; i > 0; --i) {
    STEP(ntf_act_bwd(ctx, stream, a->dact[i], a->act[i], B, h[i], 1, a->dz[i], a->gb[i], ws_main, ws_main_bytes));
    STEP(ntf_dense_bwd(ctx, stream, a->act[i - 1], a->W[i], a->dz[i], B, h[i - 1], h[i], a->gW[i], a->dact[i - 1], ws_main, ws_main_bytes));
  }
  if (!act_fused) STEP(ntf_act_bwd(ctx, stream, a->dact[0], a->act[0], B, h[0], 1, a->dz[0], a->gb[0], ws_main, ws_main_bytes));
  if (a->x_dense) {  // dW0[h0,S] = dz0^T X; the input needs no gradient
    STEP(ntf_dense_bwd(ctx, stream, a->x_dense, a->W[0], a->dz[0], B, a->S, h[0], a->gW[0], nullptr, ws_main, ws_main_bytes));
  } else {
    NTF_CUDA(cudaStreamWaitEvent(st, ctx->ev_join[1], 0));
    STEP(ntf_csr_bag_bwd_reduce_impl(ctx, st, B, a->s_indptr, a->s_indices, a->s_ent_row, a->row_base, a->dz[0], a->S, h[0], a->gW[0], ws_bag, ws_bag_bytes, nullptr,
                                     ctx->side[1]));  // (side 1 is idle since the slot fill: the hot skills' kernels go there)
  }
  // ---- optimiser: fnn.py:139 (skipped when the caller all-reduces the gradients first: data-parallel ranks) ----
  if (finish_side) NTF_CUDA(cudaStreamWaitEvent(st, ctx->ev_finish, 0));  // the fused bias gradient of layer 0 is part of the segment stepped / exchanged next
  if (dp) {
    NTF_CUDA(cudaEventRecord(ctx->ev_bwd, st));
    NTF_CUDA(cudaStreamWaitEvent(ctx->comm_st, ctx->ev_bwd, 0));
    ALLREDUCE(a->grads, opt_split);
    NTF_CUDA(cudaEventRecord(ctx->ev_ar[1], ctx->comm_st));
    NTF_CUDA(cudaStreamWaitEvent(st, ctx->ev_ar[1], 0));
  }
#undef ALLREDUCE
  const bool sh_main = shadow && w_off < opt_split;  // (only when the arena is stepped in one piece)
  if (peers) {
    STEP(ntf_peer_exchange_adam_impl(ctx, st, a->peers, a->adam_m, a->adam_v, 0, opt_split, a->lr, a->beta1, a->beta2, a->eps, a->adam_t, a->dyn, 0));
    if (sh_main) STEP(ntf_to_half(ctx, stream, a->W[Lo], w_n, const_cast<void*>(a->W16)));
  } else if (a->run_adam)
    STEP(ntf_adam_step_impl(ctx, st, a->params, a->grads, a->adam_m, a->adam_v, opt_split, a->lr, a->beta1, a->beta2, a->eps, a->adam_t, a->dyn,
                            sh_main ? const_cast<void*>(a->W16) : nullptr, sh_main ? w_off : 0, sh_main ? w_n : 0));
  if (opt_split < a->n_params) NTF_CUDA(cudaStreamWaitEvent(st, ctx->ev_join_opt, 0));
#undef STEP
  return NTF_OK;
}

// ---- test time: one call per batch of fnn.py:198-218 -- hidden layers, then the fused output layer + sigmoid + top-K (infer_topk.cu) ----
static size_t a16_region(const ntf_fnn_infer_topk_args* a) { return align_up((size_t)a->B * a->hidden[a->n_layers - 2] * 2, 1024); }

extern "C" size_t ntf_fnn_infer_topk_workspace_bytes(const ntf_ctx* ctx, const ntf_fnn_infer_topk_args* a) {
  (void)ctx;
  if (!a || a->n_layers < 2 || a->n_layers > NTF_MAX_LAYERS) return 0;
  return a16_region(a) + ntf_infer_topk_workspace_bytes(a->B, a->hidden[a->n_layers - 2], a->E, a->K);
}

extern "C" int ntf_fnn_infer_topk(ntf_ctx* ctx, void* stream, const ntf_fnn_infer_topk_args* a, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ctx && a, NTF_ERR_BAD_ARG, "fnn_infer_topk: null ctx/args");
  NTF_REQUIRE(a->n_layers >= 2 && a->n_layers <= NTF_MAX_LAYERS, NTF_ERR_UNSUPPORTED, "fnn_infer_topk: %d layers (2..%d)", a->n_layers, NTF_MAX_LAYERS);
  NTF_REQUIRE(a->B > 0 && a->S > 0 && a->E > 0 && a->K > 0, NTF_ERR_BAD_ARG, "fnn_infer_topk: B=%d S=%d E=%d K=%d", a->B, a->S, a->E, a->K);
  NTF_REQUIRE(workspace && workspace_bytes >= ntf_fnn_infer_topk_workspace_bytes(ctx, a), NTF_ERR_WORKSPACE, "fnn_infer_topk: workspace too small");
  const int L = a->n_layers, Lo = L - 1;
  const int* h = a->hidden;
  int rc;
  // one hidden layer: the bag kernel writes the fp16 operand image the tensor-core product reads next to the fp32 activations
  void* A16 = (Lo == 1 && (h[0] % 8) == 0 && !a->x_dense) ? workspace : nullptr;
  if (a->x_dense) rc = ntf_dense_fwd(ctx, stream, a->x_dense, a->W[0], a->b[0], a->B, a->S, h[0], 1, a->act[0]);
  else rc = ntf_csr_bag_fwd_impl(ctx, stream, a->B, a->s_indptr, a->s_indices, a->W[0], a->b[0], a->S, h[0], a->act[0], A16, true);
  if (rc) return rc;
  for (int i = 1; i < Lo; ++i)
    if ((rc = ntf_dense_fwd(ctx, stream, a->act[i - 1], a->W[i], a->b[i], a->B, h[i - 1], h[i], 1, a->act[i]))) return rc;
  ntf_infer_topk_args t;
  memset(&t, 0, sizeof(t));
  t.A = a->act[Lo - 1]; t.A16 = A16; t.W16 = a->W16; t.b = a->b[Lo];
  t.B = a->B; t.h = h[Lo - 1]; t.E = a->E; t.K = a->K; t.e_lo = a->e_lo;
  t.vals = a->vals; t.idx = a->idx;
  return ntf_infer_topk(ctx, stream, &t, (char*)workspace + a16_region(a), workspace_bytes - a16_region(a));
}
