/* ntf_b200.h -- C ABI of libntf_b200.so: the B200 (sm_100a) kernels behind OpeNTF's Fnn/Bnn hot path.
 *
 * The reference (fani-lab/OpeNTF) is pure Python: it has no FFI / operator registry; its "operators" for this
 * path are ATen calls made from src/mdl/fnn.py.  Each entry point below replaces one group of those call
 * sites (cited as file:line under /root/reference) and is what a ctypes/cffi binding inside the reference
 * would call (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer.
 *   - the caller owns every buffer (parameters, optimiser state, CSR, outputs, workspace); the library
 *     allocates nothing that outlives a call.  `*_workspace_bytes` tells how much scratch a call needs.
 *   - every call is asynchronous and ordered on `stream` (a cudaStream_t passed as void*).
 *   - return value: NTF_OK or a negative ntf_status; ntf_last_error() gives the message of the calling
 *     thread's last failure.  No C++ exception crosses this boundary.  There is no CPU fallback: a
 *     device that is not sm_100 makes ntf_create fail.
 *   - matrices are row-major fp32.  A batch is B consecutive rows of a CSR: `indptr` points at the batch's
 *     first row (B+1 entries, ABSOLUTE offsets into `indices`), so a batch of the epoch-permuted CSR built
 *     by ntf_csr_gather is addressed by pointer arithmetic only.
 *   - precision: NTF_FP32 = CUDA-core fp32 FMA chains (bit-stable, the parity mode);
 *                NTF_TF32 = tcgen05 kind::tf32 MMA with fp32 accumulation in TMEM (the fast mode).
 */
#ifndef NTF_B200_H
#define NTF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NTF_ABI_VERSION 2

typedef enum {
  NTF_OK = 0,
  NTF_ERR_BAD_ARG = -1,      /* null pointer, negative size, misaligned buffer ...            */
  NTF_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels are built for                   */
  NTF_ERR_CUDA = -3,         /* a CUDA runtime / driver call failed (message has the string)   */
  NTF_ERR_WORKSPACE = -4,    /* workspace smaller than *_workspace_bytes says                  */
  NTF_ERR_DEVICE = -5        /* not an sm_100 device                                           */
} ntf_status;

typedef enum { NTF_FP32 = 0, NTF_TF32 = 1 } ntf_precision;
typedef enum { NTF_NS_NONE = 0, NTF_NS_UNIFORM = 1, NTF_NS_UNIGRAM = 2, NTF_NS_UNIGRAM_B = 3 } ntf_nsd;

typedef struct ntf_ctx ntf_ctx; /* opaque per-device handle (device properties + TMA encoder entry point) */

int ntf_version(void);
int ntf_last_error(char* buf, size_t n);
/* one handle per device; the reference keeps the equivalent state in torch's globals (ntf.py:12-14) */
int ntf_create(int device, ntf_ctx** out);
int ntf_destroy(ntf_ctx* ctx);
int ntf_sm_count(const ntf_ctx* ctx);
/* number of kernels this library has launched in this process (reset != 0 clears it): launch accounting for bench.py */
unsigned long long ntf_launch_count(int reset);

/* ---- step-varying scalars on the device + CUDA graphs ----------------------------------------------------------------------
 * The reference's loop (fnn.py:118-151) launches ~40 ATen kernels per batch from Python; here a step is ~17 launches, and even
 * those cost the host more than the GPU needs to run them.  The launch sequence of batch i of a split is identical in every epoch
 * (same buffers: the epoch's permutation moves DATA, not pointers) except for three scalars: the sampler's RNG counter, and Adam's
 * learning rate and bias corrections.  Those live in a small DEVICE block (ntf_dyn) that the kernels read when they are handed
 * one, so a step captured once as a CUDA graph is replayed unchanged:   ntf_dyn_update(...); ntf_graph_launch(g, stream);      */
typedef struct {
  uint64_t step;                                                   /* counter of the sampler's Philox stream (ntf_neg_sample)   */
  float one_minus_b1, b2, one_minus_b2, bc2_sqrt, eps, neg_step;   /* Adam: neg_step = -lr / (1 - beta1^t), bc2_sqrt = sqrt(1 - beta2^t) */
} ntf_dyn;
/* writes the block (device memory, caller-owned) for the step about to run; stream-ordered, one tiny launch */
int ntf_dyn_update(ntf_ctx* ctx, void* stream, ntf_dyn* dyn, uint64_t step, double lr, double beta1, double beta2, double eps,
                   int64_t adam_t);
/* host-driven sequences (the Bnn step is ~45 separate calls): while a block is set on the handle, ntf_fill_normal, ntf_fill_sign_bits,
 * ntf_neg_sample, ntf_adam_step and ntf_peer_exchange_adam enqueue kernels that read `step` / the Adam constants from it instead of
 * their host arguments, so a sequence captured between ntf_graph_begin/end replays with this step's values.  NULL clears it. */
int ntf_set_dyn(ntf_ctx* ctx, const ntf_dyn* dyn);
/* capture everything this library enqueues on `stream` between begin and end (thread-local capture mode) into an executable
 * graph; replaying it costs the host one call.  The calls in between must get the same buffers at every replay. */
typedef struct ntf_graph ntf_graph;
int ntf_graph_begin(ntf_ctx* ctx, void* stream);
int ntf_graph_end(ntf_ctx* ctx, void* stream, ntf_graph** out);
int ntf_graph_launch(ntf_graph* g, void* stream);
int ntf_graph_kernels(const ntf_graph* g);   /* kernel nodes in the graph (launch accounting) */
int ntf_graph_destroy(ntf_graph* g);

/* ---- batching: replaces DataLoader(shuffle) + NtfDataset.__getitem__ (fnn.py:95-97, ntf.py:17-25) ------------
 * dst row i = src row rows[i] (rows == NULL: identity).  dst_indptr gets n+1 offsets (exclusive scan of the
 * row lengths), dst_ent_row[p] (nullable) = destination row of entry p.  One call per epoch with the epoch's
 * permutation builds the CSR every batch of the epoch is a slice of. */
size_t ntf_csr_gather_workspace_bytes(int n);
int ntf_csr_gather(ntf_ctx* ctx, void* stream, const int32_t* rows, int n, const int32_t* src_indptr,
                   const int32_t* src_indices, int32_t* dst_indptr, int32_t* dst_indices, int32_t* dst_ent_row,
                   void* workspace, size_t workspace_bytes);

/* dense (embedded) skill input (main.py:148-153 replaces teamsvecs['skill'] by d2v/gnn vectors; ntf.py:24 hands the row as is):
 * dst[i,:] = src[rows[i],:] (rows == NULL: identity) -- the batch-ordered copy of a split, rebuilt per epoch like ntf_csr_gather. */
int ntf_rows_gather(ntf_ctx* ctx, void* stream, const int32_t* rows, int n, int d, const float* src, float* dst);

/* ---- input layer: multi-hot skills --------------------------------------------------------------------------
 * replaces ntf.py:23 (row -> dense float), fnn.py:120 (H2D of dense X) and layer 0 of fnn.py:25
 * (aten::linear over a 99.97%-zero X, + leaky_relu):
 *   A[n,:] = lrelu( b0 + sum_{s in skills(n)} W0T[s,:] ),  W0T is [S,h] (a skill's vector is contiguous). */
int ntf_csr_bag_fwd(ntf_ctx* ctx, void* stream, int B, const int32_t* indptr, const int32_t* indices,
                    const float* W0T, const float* b0, int S, int h, float* A);
/* replaces autograd's addmm backward for layer 0 (fnn.py:137):
 *   dW0T[s,:] = sum_{n: s in skills(n)} dZ[n,:]  for EVERY s in [0,S) (zeros for skills absent from the batch).
 * No floating-point atomics, run-to-run deterministic: a skill row is owned by one warp (or, for the few skills that sit
 * in most teams, by one CTA with a fixed-order combine) and summed in entry order.
 * ent_row[p] - row_base = batch row of CSR entry p (from ntf_csr_gather). */
size_t ntf_csr_bag_bwd_workspace_bytes(int S, int h);
int ntf_csr_bag_bwd(ntf_ctx* ctx, void* stream, int B, const int32_t* indptr, const int32_t* indices,
                    const int32_t* ent_row, int row_base, const float* dZ, int S, int h, float* dW0T,
                    void* workspace, size_t workspace_bytes);

/* ---- hidden layers / dense (embedded) skill input: fnn.py:25 layers i>=1, ntf.py:24 ----------------------------
 * Y = act(A W^T + b), A [B,in], W [out,in] (torch layout); act 0 = identity, 1 = leaky_relu, 2 = sigmoid(leaky_relu). */
int ntf_dense_fwd(ntf_ctx* ctx, void* stream, const float* A, const float* W, const float* b, int B, int in,
                  int out, int act, float* Y);
/* dZ = dY * lrelu'(Y) (act=1; the sign is taken from the stored activation), db = colsum(dZ); dZ may be NULL. */
size_t ntf_act_bwd_workspace_bytes(int B, int h);
int ntf_act_bwd(ntf_ctx* ctx, void* stream, const float* dY, const float* Y, int B, int h, int act, float* dZ,
                float* db, void* workspace, size_t workspace_bytes);
/* dW[out,in] = dZ^T A ;  dA[B,in] = dZ W  (dA may be NULL). */
size_t ntf_dense_bwd_workspace_bytes(int B, int in, int out);
int ntf_dense_bwd(ntf_ctx* ctx, void* stream, const float* A, const float* W, const float* dZ, int B, int in,
                  int out, float* dW, float* dA, void* workspace, size_t workspace_bytes);

/* ---- negative sampling: fnn.py:48-76 -----------------------------------------------------------------------------
 * counts[j] = number of batch teams having expert j (fnn.py:75 `y.sum(0)` as an exact integer; for the global
 * unigram of fnn.py:82 pass all teams once); cdf = inclusive prefix sum. */
size_t ntf_expert_cdf_workspace_bytes(int E);
int ntf_expert_cdf(ntf_ctx* ctx, void* stream, int B, const int32_t* m_indptr, const int32_t* m_indices, int E,
                   uint32_t* counts, uint32_t* cdf, void* workspace, size_t workspace_bytes);
/* neg[n, 0..ns) = distinct experts that are not members of team n, drawn without replacement with probability
 * proportional to the expert counts (unigram: `cdf` of ntf_expert_cdf over all teams; unigram_b: counts over the batch)
 * or uniformly (uniform): the distribution of the reference's torch.multinomial(replacement=False) / top-ns-of-iid-keys.
 * unigram_b needs no histogram: count_j/B is the chance that a uniformly chosen ENTRY of the batch's member CSR is j, so a
 * draw picks entry number floor(u*nnz) of the pool = rows [0, pool_rows) of `pool_indptr` (NULL / 0: the batch itself;
 * a data-parallel rank passes the GLOBAL batch its rows are a slice of).
 * Counter RNG: Philox4x32-10 keyed by seed, counter = (row0+n, draw, step); restated on the CPU in
 * oracle/sampler_oracle.py.  A row whose candidates carry no mass falls back to uniform over ALL experts (fnn.py:67-69).
 * Unfilled slots are -1. */
int ntf_neg_sample(ntf_ctx* ctx, void* stream, int nsd, uint64_t seed, uint64_t step, int row0, int B,
                   const int32_t* m_indptr, const int32_t* m_indices, int E, int ns, const uint32_t* cdf,
                   const int32_t* pool_indptr, int pool_rows, int32_t* neg);
/* special[n, j/32] bit j%32 = 1 iff j is a member of team n or j is in neg[n,:]: the reference's `condition` tensor
 * (fnn.py:33-43), bit-packed.  op 1 sets the bits, op 0 clears the same words again (no full memset per step).
 * Expert-sharded output layer: the plane covers this rank's E columns [e_lo, e_lo + E) of the global expert axis; member
 * and negative ids are GLOBAL and those outside the range are skipped (e_lo = 0, E = all experts otherwise). */
int ntf_special_bits(ntf_ctx* ctx, void* stream, int op, int B, const int32_t* m_indptr, const int32_t* m_indices,
                     const int32_t* neg, int ns, int E, int e_lo, uint32_t* special, int pitch_words);

/* The same two sets, laid out for the tensor-core kernel (NTF_TF32): teams in tiles of 128, per tile one slab of Epad = roundup(E,128)
 * experts x 4 words; word ((n/128)*Epad + j)*4 + (n%128)/32, bit n%32.  special_t: j is a member of team n or in neg[n,:];
 * member_t: j is a member of team n (the target y).  One plane is ntf_special_tiles_bytes(B, E) bytes, zero-initialised by the
 * caller once; op 1 sets the bits of a batch, op 0 clears the same words again (not needed after ntf_out_train(NTF_TF32), which
 * clears what it consumes).  A CTA's (tile, 128-expert) slice is 2 KB contiguous. */
size_t ntf_special_tiles_bytes(int B, int E);
int ntf_special_tiles(ntf_ctx* ctx, void* stream, int op, int B, const int32_t* m_indptr, const int32_t* m_indices,
                      const int32_t* neg, int ns, int E, int e_lo, uint32_t* special_t, uint32_t* member_t);

/* ---- output layer, training: last layer of fnn.py:25 + fnn.py:32-46,135 + its autograd (fnn.py:137) -----------------
 * z = A W^T + b ; x = lrelu(z) ; w = special ? tpw : tnw ; y = [j is a member of team n]
 * loss_out[0] = loss_scale * sum_{n,j} w*((1-y)x + softplus(-x))          (loss_scale = 1/B of the GLOBAL batch)
 * dz = w*(sigmoid(x)-y)*loss_scale*lrelu'(z) ;  dW = dz^T A ; db = colsum(dz) ; dA = dz W
 * A validation step passes dW = db = dA = NULL (loss only, fnn.py:144-151).
 * Flipout (W_delta != NULL; A_s = A*sign_in): z += (A_s W_delta^T + b_delta)*s_out, and the extra gradients
 * dW_delta = (dz*s_out)^T A_s, db_delta = colsum(dz*s_out), dA_s = (dz*s_out) W_delta are produced. */
typedef struct {
  const float* A;              /* [B,h] last hidden activation                                      */
  const float* W;              /* [E,h] (mu for Bnn)                                                 */
  const float* b;              /* [E]                                                                */
  const uint32_t* special;     /* [B,pitch_words] bit plane (NULL: every weight is tnw)              */
  int pitch_words;
  const int32_t* m_indptr;     /* member CSR of the batch (B+1 absolute offsets)                     */
  const int32_t* m_indices;
  int B, h, E;
  float tpw, tnw, loss_scale;
  float* dW;                   /* [E,h] or NULL                                                      */
  float* db;                   /* [E]   or NULL                                                      */
  float* dA;                   /* [B,h] or NULL                                                      */
  float* loss_out;             /* [1]                                                                */
  /* Flipout extras (all NULL for Fnn) */
  const float* A_s;            /* [B,h]  A * sign_in                                                 */
  const float* W_delta;        /* [E,h]  softplus(rho_W) * eps_W                                     */
  const float* b_delta;        /* [E]                                                                */
  const uint32_t* sign_out;    /* [B,pitch_words] bit=1 -> -1                                        */
  float* dW_delta;             /* [E,h]                                                              */
  float* db_delta;             /* [E]                                                                */
  float* dA_s;                 /* [B,h]                                                              */
  /* NTF_TF32 reads these instead of `special` + the member CSR (NULL: every weight tnw, every target 0).
   * They are CONSUMED: the kernel zeroes every word it used, so the planes are clean for the next batch.   */
  uint32_t* special_t;         /* ntf_special_tiles planes                                           */
  uint32_t* member_t;
  int e_lo;                    /* expert-sharded layer: W/b/dW/db and the planes cover columns [e_lo, e_lo+E); m_indices are global */
  const void* A16;             /* NTF_TF32, optional: A as fp16 [B,h] already (the producing kernel wrote it); NULL: converted here */
  const void* W16;             /* NTF_TF32, optional: fp16 image of W [E,h] kept by the caller (ntf_to_half; the optimiser entry points of ntf_fnn_step
                                  rewrite it); NULL: made per call.  Ignored by the Flipout layer */
  const int32_t* neg;          /* NTF_TF32 Fnn (persistent kernel, csrc/out_tc2.cu): the sampled negatives [B,ns] (global expert ids, -1 = none) --
                                  with the member CSR they name the (team, expert) pairs of weight tpw directly; the tensor-core pass treats every
                                  pair as (target 0, weight tnw) and a sparse correction pass puts these few right.  Given `neg` (or ns == 0 and
                                  special_t == NULL) the planes are not needed */
  int ns;
  /* optional fusion of the step that follows when layer n_layers-2 feeds the output layer directly and dA needs no exchange (one hidden
   * layer, not expert-sharded): the correction pass, which owns each team's dA row last, also applies the activation's derivative and
   * the bias gradient of that layer -- dz_prev = dA * lrelu'(act_prev), db_prev = colsum(dz_prev) -- i.e. ntf_act_bwd; dA itself is then
   * not written back.  All three or none. */
  const float* act_prev;       /* [B,h] lrelu output of the previous layer                              */
  float* dz_prev;              /* [B,h]                                                                  */
  float* db_prev;              /* [h]                                                                    */
  int prepared;                /* != 0: the caller ran ntf_out_train_prepare for this call already (off the critical path) */
  int defer_finish;            /* != 0: leave the call's final reductions (loss_out; db_prev when fused) to ntf_out_train_finish, which the caller
                                  runs with the same args / workspace on any stream ordered after this call -- they gate nothing but the optimiser */
} ntf_out_train_args;
int ntf_out_train_finish(ntf_ctx* ctx, void* stream, const ntf_out_train_args* args, void* workspace, size_t workspace_bytes);
/* NTF_TF32 Fnn: what ntf_out_train has to clear before its kernel -- dA and the dW/db rows of expert tiles that two CTAs share.  A caller
 * that knows the shapes early (ntf_fnn_step: at the top of the step, on a side stream) runs it there and sets args.prepared. */
int ntf_out_train_prepare(ntf_ctx* ctx, void* stream, int B, int h, int E, float* dW, float* db, float* dA);
/* 1 if NTF_TF32 has a tcgen05 kernel for this shape (else callers use NTF_FP32; ntf_out_train(NTF_TF32) refuses it) */
int ntf_tc_supported(int B, int h, int E, int flipout);
size_t ntf_out_train_workspace_bytes(const ntf_ctx* ctx, int precision, int B, int h, int E, int flipout);
int ntf_out_train(ntf_ctx* ctx, void* stream, int precision, const ntf_out_train_args* args, void* workspace,
                  size_t workspace_bytes);

/* ---- optimiser: torch.optim.Adam (fnn.py:104,139), dense, one launch over the flat parameter arena; step is 1-based */
int ntf_adam_step(ntf_ctx* ctx, void* stream, float* p, const float* g, float* m, float* v, size_t n, double lr,
                  double beta1, double beta2, double eps, int64_t step);

/* ---- test time: fnn.py:200-213 + pkgmgr.py:125-134 (+ the ranking of evl/metric.py:17-28) ---------------------------
 * P[n,j] = sigmoid(lrelu(A W^T + b)) (Flipout extras as above, may be NULL).  accumulate != 0 adds into P. */
size_t ntf_infer_scores_workspace_bytes(int B, int E, int flipout);
int ntf_infer_scores(ntf_ctx* ctx, void* stream, int precision, const float* A, const float* W, const float* b,
                     int B, int h, int E, const float* A_s, const float* W_delta, const float* b_delta,
                     const uint32_t* sign_out, int pitch_words, int accumulate, float* P, void* workspace,
                     size_t workspace_bytes);
/* per row: the K largest of scale*P[n,:] in rank order (value descending, ties -> lower expert id first); K <= 2048 */
int ntf_topk_select(ntf_ctx* ctx, void* stream, const float* P, int B, int E, int K, float scale, float* vals,
                    int32_t* idx);
/* K5, fused: the K best experts per team straight from the last hidden layer's activations -- output layer product, sigmoid and
 * top-K in one call; the [B,E] score matrix of fnn.py:211 never reaches HBM (csrc/infer_topk.cu: two passes of the tcgen05 product,
 * block maxima -> threshold -> candidates).  Replaces `sigmoid(model(X))` + `topk_sparse` (fnn.py:204-218, pkgmgr.py:125-134) when
 * testcfg.topK < E.  Rank order and values as ntf_infer_scores(NTF_TF32) + ntf_topk_select give (ties between DISTINCT logits whose
 * probabilities round to the same float are ranked by logit).  idx = e_lo + column (global expert ids of an expert shard). */
typedef struct ntf_infer_topk_args {
  const float* A;   /* [B,h] activations of the last hidden layer (read when A16 is NULL) */
  const void* A16;  /* [B,h] the same as fp16 (e.g. written by the input-layer kernel), or NULL */
  const void* W16;  /* [E,h] fp16 image of the output layer's weight (ntf_to_half of layers.L.weight) */
  const float* b;   /* [E] */
  int B, h, E, K, e_lo;
  float* vals;      /* [B,K] out: probabilities, rank order */
  int32_t* idx;     /* [B,K] out: expert ids */
} ntf_infer_topk_args;
int ntf_infer_topk_supported(int B, int h, int E, int K); /* h == 128, K <= 1024, 32*K <= E <= 131072 */
size_t ntf_infer_topk_workspace_bytes(int B, int h, int E, int K);
int ntf_infer_topk(ntf_ctx* ctx, void* stream, const ntf_infer_topk_args* args, void* workspace, size_t workspace_bytes);
/* y[i] = fp16(x[i]) round-to-nearest: the fp16 operand images the tensor-core kernels read (10-bit mantissa = TF32's) */
int ntf_to_half(ntf_ctx* ctx, void* stream, const float* x, size_t n, void* y);
/* merge G per-shard candidate lists [G][B][K] (idx = global expert ids, -1 = empty) into the global top-K per row:
 * the merge step of the expert-sharded output layer. */
int ntf_topk_merge(ntf_ctx* ctx, void* stream, const float* vals_in, const int32_t* idx_in, int G, int B, int K,
                   float* vals, int32_t* idx);
/* Bnn test-time uncertainty (fnn.py:206-208): out[n] (+)= -sum_j q log(q+1e-15), q = scale*P[n,j] */
int ntf_row_entropy(ntf_ctx* ctx, void* stream, const float* P, int B, int E, float scale, int accumulate, float* out);
/* ranking metrics of the top-K output on the device: the per-team loop of src/evl/metric.py:12-33 (argpartition/argsort, two
 * dicts per team, pytrec_eval) called from Ntf.evaluate (src/mdl/ntf.py:59-62).  idx/vals [n,K]: every team's candidates in ANY order
 * (idx < 0 = padding); m_indptr (n+1 absolute offsets) / m_indices: the teams' true members, columns ascending; ks[nk]: cut-offs,
 * ascending, HOST array (<= 16); out [n][5][nk] fp64 (device), families in the order P, recall, ndcg_cut, map_cut, success.
 * Ranking follows trec_eval: score descending, ties by the document id STRING ('d'+str(expert), metric.py:31) descending. K <= 1024. */
int ntf_eval_ranked(ntf_ctx* ctx, void* stream, int n, int K, const int32_t* idx, const float* vals, const int32_t* m_indptr,
                    const int32_t* m_indices, const int* ks, int nk, double* out);
int ntf_axpy(ntf_ctx* ctx, void* stream, size_t n, float a, const float* x, float* y); /* y += a*x (MC mean, fnn.py:209) */

/* ---- staging either side of the path: src/cmn/team.py:148-173, 318-341; src/evl/metric.py:44-73 ------------------
 * ntf_csr_from_lists replaces Team.bucketing / get_one_hot (team.py:148-173: a dense one-hot row per team copied into a lil_matrix):
 * ragged id lists (indptr[n+1], ids[n_ids] in ANY order, duplicates allowed, every id in [0, n_cols)) -> the sorted duplicate-free CSR
 * rows of the multi-hot matrix.  dst_indices needs n_ids slots; *nnz (host) = entries kept.  Synchronises the stream.
 * n_cols <= ~1.8 M (a bitmap of the columns has to fit the shared memory of an SM). */
size_t ntf_csr_from_lists_workspace_bytes(int n, size_t n_ids);
int ntf_csr_from_lists(ntf_ctx* ctx, void* stream, int n, size_t n_ids, const int32_t* indptr, const int32_t* ids, int n_cols,
                       int32_t* dst_indptr, int32_t* dst_indices, int64_t* nnz, void* workspace, size_t workspace_bytes);
/* replaces Team.gen_skill_coverage (team.py:318-341): member^T . skill as a sparse [E,S] matrix, value = number of (not skipped) teams
 * in which expert e and skill s occur together (int32; the reference's uint8 arithmetic wraps at 256 -- the caller narrows).
 * skip (nullable): uint8[T], 1 = leave the team out (team.py:327-332, test teams).  Two calls around the caller's allocation:
 *   _count  transposes the member matrix into the workspace, writes co_indptr[E+1], returns the entry count in *nnz (host; synchronises)
 *   _fill   writes co_indices (ascending per row) and co_values [nnz]; same workspace, untouched in between. */
size_t ntf_cooccur_workspace_bytes(int E, size_t nnz_member);
int ntf_cooccur_count(ntf_ctx* ctx, void* stream, int T, int E, int S, size_t nnz_member, const int32_t* m_indptr,
                      const int32_t* m_indices, const int32_t* s_indptr, const int32_t* s_indices, const uint8_t* skip,
                      int32_t* co_indptr, int64_t* nnz, void* workspace, size_t workspace_bytes);
int ntf_cooccur_fill(ntf_ctx* ctx, void* stream, int T, int E, int S, size_t nnz_member, const int32_t* s_indptr,
                     const int32_t* s_indices, const int32_t* co_indptr, int32_t* co_indices, int32_t* co_values, void* workspace,
                     size_t workspace_bytes);
/* replaces the per-team loop of calculate_skill_coverage (metric.py:53-69): out[t, j] = |skills(t) n U_{r < ks[j]} skills_of(idx[t, r])|
 * / |skills(t)| (fp64; a team without skills gives nan like numpy's 0/0).  idx [n,K]: the team's experts in rank order (-1 = none);
 * x_*: the teams' required skills (CSR rows, n+1 absolute offsets); co_*: the co-occurrence matrix's pattern, columns ascending,
 * entries whose (narrowed) value is 0 removed by the caller; ks: HOST array of nk <= 8 cut-offs. */
int ntf_skill_coverage(ntf_ctx* ctx, void* stream, int n, int K, const int32_t* idx, const int32_t* x_indptr, const int32_t* x_indices,
                       const int32_t* co_indptr, const int32_t* co_indices, const int* ks, int nk, double* out);

/* ---- Bnn / Flipout parameter pass (bayesian-torch 0.5.0 LinearFlipout + get_kl_loss; bnn.py:19-25, fnn.py:136) ---------
 * delta = softplus(rho)*eps ;  kl_out[0] += kl_scale * sum(-log sigma + (sigma^2 + mu^2)/2 - 1/2)   (kl_scale = 1/numel:
 * the reference takes a MEAN per tensor, prior N(0,1)) */
size_t ntf_flipout_prepare_workspace_bytes(const ntf_ctx* ctx);
int ntf_flipout_prepare(ntf_ctx* ctx, void* stream, const float* mu, const float* rho, const float* eps, size_t n,
                        float kl_scale, float* delta, float* kl_out, void* workspace, size_t workspace_bytes);
/* g_mu += kl_gscale*mu ; g_rho = (g_delta*eps + kl_gscale*(sigma - 1/sigma)) * sigmoid(rho)   (kl_gscale = 1/(numel*B)) */
int ntf_flipout_grads(ntf_ctx* ctx, void* stream, const float* mu, const float* rho, const float* eps,
                      const float* g_delta, size_t n, float kl_gscale, float* g_mu, float* g_rho);
/* eps ~ N(0,1): Philox(seed; counter = (i/4, stream_id, step)) + Box-Muller */
int ntf_fill_normal(ntf_ctx* ctx, void* stream, uint64_t seed, uint64_t step, uint32_t stream_id, size_t n, float* out);
/* bit planes of iid fair signs (bit=1 -> -1): Philox(seed; counter = (word/4, stream_id, step)) */
int ntf_fill_sign_bits(ntf_ctx* ctx, void* stream, uint64_t seed, uint64_t step, uint32_t stream_id, size_t n_words,
                       uint32_t* bits);
/* As[n,c] = A[n,c] * (bit(n,c) ? -1 : +1) */
int ntf_apply_sign(ntf_ctx* ctx, void* stream, const float* A, const uint32_t* bits, int pitch_words, int B, int h,
                   float* As);

/* Flipout input layer on multi-hot rows (LinearFlipout with x in {0,1}: the input sign matters only at the nnz positions):
 *   A[n,:] = lrelu( b_mu + sum_p W0T_mu[s_p,:] + (b_delta + sum_p sgn_p W0T_delta[s_p,:]) * s_out[n,:] )
 * ent_sign: one bit per CSR entry of the batch, bit (p - indptr[0]), 1 -> -1 (fill with ntf_fill_sign_bits). */
int ntf_csr_bag_flipout_fwd(ntf_ctx* ctx, void* stream, int B, const int32_t* indptr, const int32_t* indices,
                            const uint32_t* ent_sign, const float* W0T_mu, const float* b_mu, const float* W0T_delta,
                            const float* b_delta, const uint32_t* sign_out, int pitch_words, int S, int h, float* A);
/* dW0T_delta[s,:] = sum_{entries p=(n,s)} sgn_p dZs[n,:]; workspace as ntf_csr_bag_bwd */
int ntf_csr_bag_bwd_signed(ntf_ctx* ctx, void* stream, int B, const int32_t* indptr, const int32_t* indices,
                           const int32_t* ent_row, int row_base, const uint32_t* ent_sign, const float* dZs, int S, int h,
                           float* dW0T_delta, void* workspace, size_t workspace_bytes);
/* Flipout hidden layer: Y = act( A W^T + b + (A_s W_delta^T + b_delta) * s_out ) */
size_t ntf_dense_flipout_fwd_workspace_bytes(int B, int out);
int ntf_dense_flipout_fwd(ntf_ctx* ctx, void* stream, const float* A, const float* W, const float* b, const float* A_s,
                          const float* W_delta, const float* b_delta, const uint32_t* sign_out, int pitch_words, int B, int in,
                          int out, int act, float* Y, void* workspace, size_t workspace_bytes);
/* Y[n,c] += X[n,c] * (bit(n,c) ? -1 : +1) */
int ntf_add_signed(ntf_ctx* ctx, void* stream, const float* X, const uint32_t* bits, int pitch_words, int B, int h, float* Y);

/* ---- data-parallel ranks: gradient exchange + optimiser in one pass over peer memory ---------------------------------
 * Replaces "all-reduce the gradients, then torch.optim.Adam on every rank" (fnn.py:137-139 under data parallelism): rank r owns the
 * r-th slice of the flat arena; it sums that slice of every rank's gradients straight out of the peers' arenas (NVLink loads, rank
 * order), applies Adam to its slice of (p, m, v) and stores the new parameters into EVERY rank's arena (peer stores) -- reduce-scatter,
 * optimiser and all-gather fused, between two flag barriers in peer memory.  Per rank 4P(G-1)/G bytes each way over the links and 28P/G
 * bytes of HBM; m/v are only ever touched on the owner's slice.  The replicas stay bit-identical (one rank computes each element).
 * Arenas and flag blocks are cudaMalloc blocks of the owning process, made visible to the others through CUDA IPC:
 *   ntf_peer_alloc (zeroed) / ntf_peer_free; ntf_peer_export -> 64-byte handle (send it to the other processes by any means);
 *   ntf_peer_import -> the address of that block in this process / ntf_peer_release. */
#define NTF_MAX_PEERS 8
#define NTF_PEER_HANDLE_BYTES 64
#define NTF_PEER_FLAG_BYTES 256
typedef struct {
  int rank, world;
  float* grads[NTF_MAX_PEERS];     /* gradient arena of every rank, as addressed from THIS process (own arena included)  */
  float* params[NTF_MAX_PEERS];    /* parameter arena of every rank                                                      */
  uint32_t* flags[NTF_MAX_PEERS];  /* NTF_PEER_FLAG_BYTES per rank, zero before the first exchange                       */
} ntf_peers;
int ntf_peer_alloc(ntf_ctx* ctx, size_t bytes, void** out);
int ntf_peer_free(ntf_ctx* ctx, void* p);
int ntf_peer_export(ntf_ctx* ctx, void* p, void* handle64);
int ntf_peer_import(ntf_ctx* ctx, const void* handle64, void** out);
int ntf_peer_release(ntf_ctx* ctx, void* imported);
/* floats [offset, offset+n) of the arenas (multiples of 4); every rank calls it with the same arguments, stream-ordered after its
 * gradients are complete.  Returns once enqueued; the parameters are consistent on every rank when the call's work completes.
 * channel (0|1): exchanges that may be in flight at the same time (different streams) must use different channels (flag sets). */
int ntf_peer_exchange_adam(ntf_ctx* ctx, void* stream, const ntf_peers* peers, float* adam_m, float* adam_v, size_t offset, size_t n,
                           double lr, double beta1, double beta2, double eps, int64_t step, int channel);
/* expert-sharded output layer: dst[i] = sum over ranks r (in rank order) of peers->grads[r][i], i < n (a multiple of 4) -- the dA [B,h]
 * exchange of a sharded step as one pass over peer memory between two flag barriers.  peers->grads[r] = rank r's partial (a peer-visible
 * block, ntf_peer_alloc); dst is private and must not alias it; peers->params is not used.  Every rank calls it with the same n. */
int ntf_peer_allreduce(ntf_ctx* ctx, void* stream, const ntf_peers* peers, size_t n, float* dst, int channel);

/* ---- one whole Fnn batch, enqueued by one call: the loop body of fnn.py:118-151 -------------------------------------------------
 * forward (CSR bag, hidden layers), negative sampling (unless neg_given), output layer forward + weighted BCE; and when `train`:
 * the backward pass and (when `run_adam`) the Adam step.  Exactly the sequence of the entry points above; nothing is synchronised. */
#define NTF_MAX_LAYERS 8
/* ncclAllReduce(sendbuff, recvbuff, count, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) -> ncclResult_t (0 = success) */
typedef int (*ntf_allreduce_fn)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, void* stream);
typedef struct {
  int n_layers;                      /* linear layers (hidden layers + 1)                                              */
  int S, E;                          /* E = the output columns THIS call owns (all experts, or one shard of them)      */
  int e_lo, E_total;                 /* expert-sharded output layer: columns [e_lo, e_lo+E) of E_total (0, E otherwise) */
  int phase;                         /* 3 = whole step; 1 = up to and including the output layer; 2 = the rest of the backward pass +
                                        Adam (a sharded layer all-reduces dact[last] and the loss between the two)     */
  int hidden[NTF_MAX_LAYERS];        /* widths of the hidden layers: n_layers - 1 entries                             */
  const float* W[NTF_MAX_LAYERS];    /* layer 0 TRANSPOSED [S,h0] (ntf_csr_bag_fwd); layer i>0 [out,in]                */
  const float* b[NTF_MAX_LAYERS];
  float* gW[NTF_MAX_LAYERS];         /* gradients, same layouts                                                        */
  float* gb[NTF_MAX_LAYERS];
  float* act[NTF_MAX_LAYERS];        /* [B,h_i] activations of hidden layer i (outputs)                               */
  float* dact[NTF_MAX_LAYERS];       /* [B,h_i] scratch: d loss / d act_i                                              */
  float* dz[NTF_MAX_LAYERS];         /* [B,h_i] scratch: d loss / d pre-activation                                     */
  int B;                             /* this call's teams: B consecutive rows of the batch-ordered CSRs               */
  const int32_t* s_indptr;           /* skills: B+1 absolute offsets, first row of the batch                          */
  const int32_t* s_indices;
  const int32_t* s_ent_row;          /* entry -> row of the gathered CSR (ntf_csr_gather); row_base = first batch row */
  int row_base;
  const int32_t* m_indptr;           /* members, likewise                                                             */
  const int32_t* m_indices;
  int gB;                            /* unigram_b over a GLOBAL batch this call is a slice of (data-parallel ranks):  */
  const int32_t* g_m_indptr;         /*   its size and first row; gB = 0: the batch itself                            */
  int nsd;                           /* ntf_nsd                                                                        */
  uint64_t seed, step;
  int row0, ns, neg_given;           /* neg_given != 0: neg[B,ns] holds caller-supplied indices (parity tests)        */
  int32_t* neg;
  uint32_t* counts;                  /* (unused; kept for layout)                                                      */
  uint32_t* cdf;                     /* [E]: NTF_NS_UNIGRAM only, prepared by the caller once (ntf_expert_cdf, fnn.py:82) */
  int precision;                     /* ntf_precision                                                                  */
  float tpw, tnw, loss_scale;
  float* loss_out;
  uint32_t* special; int pitch_words;        /* NTF_FP32: row-major condition plane                                    */
  uint32_t* special_t; uint32_t* member_t;   /* NTF_TF32: tile-transposed planes                                       */
  int train, run_adam;
  float* params; float* grads; float* adam_m; float* adam_v; size_t n_params;   /* the flat arena (ntf_adam_step)      */
  double lr, beta1, beta2, eps; int64_t adam_t;
  void* prof_ev[2];                  /* optional cudaEvent_t pair recorded around the output-layer call (bench.py's roofline timing) */
  const ntf_dyn* dyn;                /* optional (device): `step`, `lr`, `adam_t` are read from this block instead -- graph replays */
  void* comm; void* allreduce;       /* data-parallel ranks: an NCCL communicator as an opaque pointer and the address of
                                        ncclAllReduce (ntf_allreduce_fn).  Non-NULL (train, phase 3, run_adam): the step sums the gradient
                                        arena over the ranks itself, in two segments on its own stream -- the output layer's while the
                                        hidden layers' backward runs, the rest while the output layer's segment is stepped -- so the
                                        whole data-parallel step is ONE capturable launch sequence.  The library does not link NCCL. */
  const ntf_peers* peers;            /* data-parallel ranks, preferred over comm: the exchange and Adam run as ntf_peer_exchange_adam inside
                                        the step (train, phase 3, run_adam), the output layer's segment next to the hidden layers' backward.
                                        EXPERT-SHARDED layer (E < E_total): the table of the dA exchange blocks instead (peers->grads[r] = rank r's
                                        [B,h_last] partial): the output layer writes this rank's dA there and ntf_peer_allreduce sums the shards into
                                        dact[n_layers-2] inside the step, so that a sharded step, too, is one call (phase 3) / one graph */
  const float* x_dense;              /* dense (embedded) skill input, ntf.py:24: [B,S] rows of this batch.  Non-NULL: layer 0 is a dense
                                        layer (ntf_dense_fwd / ntf_dense_bwd), W[0] / gW[0] are in torch layout [h0,S], s_* unused */
  void* W16;                         /* NTF_TF32, optional: fp16 image [E,h_last] of the output layer's weight W[n_layers-1], made once with ntf_to_half.
                                        The output layer TMA-loads it instead of converting W, and the optimiser launch that steps that weight rewrites
                                        it (after a peer-memory exchange: a local conversion pass), so it stays current from step to step */
} ntf_fnn_step_args;
size_t ntf_fnn_step_workspace_bytes(const ntf_ctx* ctx, const ntf_fnn_step_args* args);
int ntf_fnn_step(ntf_ctx* ctx, void* stream, const ntf_fnn_step_args* args, void* workspace, size_t workspace_bytes);

/* one call per test batch (the loop body of fnn.py:198-218): hidden layers (layer 0 = CSR bag, or dense rows when x_dense is set), then
 * ntf_infer_topk.  W[i]/b[i]: layer i's parameters as in ntf_fnn_step_args (layer 0 transposed [S,h0] for the CSR input); act[i]: [B,hidden[i]]
 * scratch; W16: fp16 image of W[n_layers-1]. */
typedef struct ntf_fnn_infer_topk_args {
  int n_layers, S, E, e_lo;
  int hidden[NTF_MAX_LAYERS];
  const float* W[NTF_MAX_LAYERS];
  const float* b[NTF_MAX_LAYERS];
  float* act[NTF_MAX_LAYERS];
  int B;
  const int32_t* s_indptr;  /* B+1 absolute offsets into s_indices */
  const int32_t* s_indices;
  const float* x_dense;     /* [B,S] embedded skills (ntf.py:24) instead of the CSR rows, or NULL */
  const void* W16;
  int K;
  float* vals;
  int32_t* idx;
} ntf_fnn_infer_topk_args;
size_t ntf_fnn_infer_topk_workspace_bytes(const ntf_ctx* ctx, const ntf_fnn_infer_topk_args* args);
int ntf_fnn_infer_topk(ntf_ctx* ctx, void* stream, const ntf_fnn_infer_topk_args* args, void* workspace, size_t workspace_bytes);


/* HOST function (no device work): rows `rows[0..n)` of the skill / member CSR (int32 indptr / indices of all teams) -> the int32 block
 * [s_indptr n+1 | m_indptr n+1 | s_indices cap_s | s_ent_row cap_s | m_indices cap_m] (offsets from 0, segments zero-padded to the capacities)
 * that the streaming entry point copies to the device in one transfer.  Replaces the per-row work of the reference's Dataset / DataLoader
 * (ntf.py:17-25, fnn.py:95,118-121).  `out` is typically pinned host memory.  [lo, hi): the rows this rank trains on -- the skill CSR
 * is packed for them only (empty skill rows elsewhere), the member CSR for all n rows (unigram_b samples from the whole batch). */
int ntf_pack_host_batch(const int32_t* rows, int n, const int32_t* s_indptr, const int32_t* s_indices, const int32_t* m_indptr,
                        const int32_t* m_indices, int cap_s, int cap_m, int lo, int hi, int32_t* out, size_t out_words);

/* The copies of the streaming entry point (host batches in, a loss out per step: fnn.py:118-140) on the context's own copy streams (one per
 * direction), so that batch i+1 goes up while step i computes and loss i comes down while step i+1 computes; `slot` (0/1, alternating)
 * names the device block / the events.
 *   _upload    upload stream: waits for the step that read this slot's block last (the _download call of two steps ago marks it), then
 *              cudaMemcpyAsync(dst_dev <- src_pinned); `stream` then waits for the copy
 *   _download  marks the end of the step on `stream`; the download stream waits for it, then copies *loss_dev to *loss_pinned
 *   _wait      the host blocks until the download of `slot` has landed */
int ntf_host_batch_upload(ntf_ctx* ctx, void* stream, int slot, void* dst_dev, const void* src_pinned, size_t bytes);
int ntf_host_loss_download(ntf_ctx* ctx, void* stream, int slot, const float* loss_dev, float* loss_pinned);
int ntf_host_loss_wait(ntf_ctx* ctx, int slot);

/* out[i] = sum_k parts[k*part_stride + i] in a fixed order (deterministic split reductions) */
int ntf_sum_parts(ntf_ctx* ctx, void* stream, const float* parts, int nparts, size_t n, size_t part_stride, float* out);

#ifdef __cplusplus
}
#endif
#endif /* NTF_B200_H */
