// api.cu -- handle, error string, version; no kernels.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";
unsigned long long g_ntf_launches = 0;

void ntf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int ntf_version(void) { return NTF_ABI_VERSION; }

extern "C" int ntf_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return NTF_ERR_BAD_ARG;
  strncpy(buf, g_err, n - 1);
  buf[n - 1] = 0;
  return NTF_OK;
}

extern "C" int ntf_create(int device, ntf_ctx** out) {
  NTF_REQUIRE(out != nullptr, NTF_ERR_BAD_ARG, "ntf_create: out is NULL");
  *out = nullptr;
  int count = 0;
  NTF_CUDA(cudaGetDeviceCount(&count));
  NTF_REQUIRE(device >= 0 && device < count, NTF_ERR_BAD_ARG, "ntf_create: device %d of %d", device, count);
  cudaDeviceProp prop;
  NTF_CUDA(cudaGetDeviceProperties(&prop, device));
  // no CPU fallback and no other architecture: the kernels are compiled for sm_100a only
  NTF_REQUIRE(prop.major == 10, NTF_ERR_DEVICE, "ntf_create: device %d is sm_%d%d, libntf_b200 needs sm_100 (B200)",
              device, prop.major, prop.minor);
  ntf_ctx* c = new ntf_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  c->encode_tiled = nullptr;
  c->dyn_override = nullptr;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) c->encode_tiled = fn;
  (void)cudaGetLastError();
  int cur = 0;
  NTF_CUDA(cudaGetDevice(&cur));
  NTF_CUDA(cudaSetDevice(device));
  cudaError_t es = cudaSuccess;
  for (int i = 0; i < 2 && es == cudaSuccess; ++i) {
    es = cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking);
    if (es == cudaSuccess) es = cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming);
  }
  if (es == cudaSuccess) es = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
  if (es == cudaSuccess) es = cudaEventCreateWithFlags(&c->ev_fork_opt, cudaEventDisableTiming);
  if (es == cudaSuccess) es = cudaEventCreateWithFlags(&c->ev_join_opt, cudaEventDisableTiming);
  if (es == cudaSuccess) es = cudaEventCreateWithFlags(&c->ev_hot_fork, cudaEventDisableTiming);
  if (es == cudaSuccess) es = cudaEventCreateWithFlags(&c->ev_hot_join, cudaEventDisableTiming);
  if (es == cudaSuccess) es = cudaStreamCreateWithFlags(&c->comm_st, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && es == cudaSuccess; ++i) es = cudaEventCreateWithFlags(&c->ev_ar[i], cudaEventDisableTiming);
  if (es == cudaSuccess) es = cudaEventCreateWithFlags(&c->ev_bwd, cudaEventDisableTiming);
  if (es == cudaSuccess) es = cudaEventCreateWithFlags(&c->ev_finish, cudaEventDisableTiming);
  if (es == cudaSuccess) es = cudaStreamCreateWithFlags(&c->copy_st, cudaStreamNonBlocking);
  if (es == cudaSuccess) es = cudaStreamCreateWithFlags(&c->copy_dn, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && es == cudaSuccess; ++i) {
    es = cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming);
    if (es == cudaSuccess) es = cudaEventCreateWithFlags(&c->ev_step[i], cudaEventDisableTiming);
    if (es == cudaSuccess) es = cudaEventCreateWithFlags(&c->ev_loss[i], cudaEventDisableTiming);
  }
  cudaSetDevice(cur);
  if (es != cudaSuccess) {
    ntf_set_error("ntf_create: side streams / events: %s", cudaGetErrorString(es));
    delete c;
    return NTF_ERR_CUDA;
  }
  *out = c;
  return NTF_OK;
}

extern "C" int ntf_destroy(ntf_ctx* ctx) {
  if (ctx) {
    for (int i = 0; i < 2; ++i) { cudaStreamDestroy(ctx->side[i]); cudaEventDestroy(ctx->ev_join[i]); }
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_fork_opt); cudaEventDestroy(ctx->ev_join_opt);
    cudaEventDestroy(ctx->ev_hot_fork); cudaEventDestroy(ctx->ev_hot_join);
    cudaStreamDestroy(ctx->comm_st); cudaEventDestroy(ctx->ev_ar[0]); cudaEventDestroy(ctx->ev_ar[1]); cudaEventDestroy(ctx->ev_bwd); cudaEventDestroy(ctx->ev_finish);
    cudaStreamDestroy(ctx->copy_st); cudaStreamDestroy(ctx->copy_dn);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(ctx->ev_up[i]); cudaEventDestroy(ctx->ev_step[i]); cudaEventDestroy(ctx->ev_loss[i]); }
  }
  delete ctx;
  return NTF_OK;
}

extern "C" int ntf_host_batch_upload(ntf_ctx* ctx, void* stream, int slot, void* dst_dev, const void* src_pinned, size_t bytes) {
  NTF_REQUIRE(ctx && dst_dev && src_pinned && (slot == 0 || slot == 1), NTF_ERR_BAD_ARG, "host_batch_upload: null pointer or slot %d", slot);
  NTF_CUDA(cudaStreamWaitEvent(ctx->copy_st, ctx->ev_step[slot], 0));  // the step that read this slot's device block last (two calls ago) must be through
  NTF_CUDA(cudaMemcpyAsync(dst_dev, src_pinned, bytes, cudaMemcpyHostToDevice, ctx->copy_st));
  NTF_CUDA(cudaEventRecord(ctx->ev_up[slot], ctx->copy_st));
  NTF_CUDA(cudaStreamWaitEvent(as_stream(stream), ctx->ev_up[slot], 0));
  return NTF_OK;
}

extern "C" int ntf_host_loss_download(ntf_ctx* ctx, void* stream, int slot, const float* loss_dev, float* loss_pinned) {
  NTF_REQUIRE(ctx && loss_dev && loss_pinned && (slot == 0 || slot == 1), NTF_ERR_BAD_ARG, "host_loss_download: null pointer or slot %d", slot);
  NTF_CUDA(cudaEventRecord(ctx->ev_step[slot], as_stream(stream)));
  NTF_CUDA(cudaStreamWaitEvent(ctx->copy_dn, ctx->ev_step[slot], 0));
  NTF_CUDA(cudaMemcpyAsync(loss_pinned, loss_dev, sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_dn));
  NTF_CUDA(cudaEventRecord(ctx->ev_loss[slot], ctx->copy_dn));
  return NTF_OK;
}

extern "C" int ntf_host_loss_wait(ntf_ctx* ctx, int slot) {
  NTF_REQUIRE(ctx && (slot == 0 || slot == 1), NTF_ERR_BAD_ARG, "host_loss_wait: slot %d", slot);
  NTF_CUDA(cudaEventSynchronize(ctx->ev_loss[slot]));
  return NTF_OK;
}

extern "C" int ntf_set_dyn(ntf_ctx* ctx, const ntf_dyn* dyn) {
  NTF_REQUIRE(ctx, NTF_ERR_BAD_ARG, "set_dyn: null ctx");
  ctx->dyn_override = dyn;
  return NTF_OK;
}

extern "C" int ntf_sm_count(const ntf_ctx* ctx) { return ctx ? ctx->sm_count : NTF_ERR_BAD_ARG; }

extern "C" unsigned long long ntf_launch_count(int reset) {
  const unsigned long long v = g_ntf_launches;
  if (reset) g_ntf_launches = 0;
  return v;
}

// ---- CUDA graphs: a captured step replayed with one host call (include/ntf_b200.h) ----------------------------------------
struct ntf_graph {
  cudaGraphExec_t exec;
  int kernels;
};
static thread_local unsigned long long g_capture_mark = 0;

extern "C" int ntf_graph_begin(ntf_ctx* ctx, void* stream) {
  NTF_REQUIRE(ctx, NTF_ERR_BAD_ARG, "graph_begin: null ctx");
  g_capture_mark = g_ntf_launches;
  NTF_CUDA(cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeThreadLocal));
  return NTF_OK;
}

extern "C" int ntf_graph_end(ntf_ctx* ctx, void* stream, ntf_graph** out) {
  NTF_REQUIRE(ctx && out, NTF_ERR_BAD_ARG, "graph_end: null pointer");
  *out = nullptr;
  cudaGraph_t graph = nullptr;
  NTF_CUDA(cudaStreamEndCapture(as_stream(stream), &graph));
  const int kernels = (int)(g_ntf_launches - g_capture_mark);
  g_ntf_launches -= (unsigned long long)kernels;  // captured launches did not run; replays are counted by ntf_graph_launch
  cudaGraphExec_t exec = nullptr;
  const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  NTF_REQUIRE(e == cudaSuccess, NTF_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
  ntf_graph* g = new ntf_graph();
  g->exec = exec;
  g->kernels = kernels;
  *out = g;
  return NTF_OK;
}

extern "C" int ntf_graph_launch(ntf_graph* g, void* stream) {
  NTF_REQUIRE(g && g->exec, NTF_ERR_BAD_ARG, "graph_launch: null graph");
  NTF_CUDA(cudaGraphLaunch(g->exec, as_stream(stream)));
  g_ntf_launches += (unsigned long long)g->kernels;
  return NTF_OK;
}

extern "C" int ntf_graph_kernels(const ntf_graph* g) { return g ? g->kernels : NTF_ERR_BAD_ARG; }

extern "C" int ntf_graph_destroy(ntf_graph* g) {
  if (!g) return NTF_OK;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  delete g;
  return NTF_OK;
}

// ---- host side of the streaming entry point (no device work): one global batch of the teamsvecs CSR -> the pinned block Engine.step_host
// copies to the device in ONE transfer.  Layout (int32 words): [s_indptr n+1 | m_indptr n+1 | s_indices cap_s | s_ent_row cap_s | m_indices cap_m];
// offsets start at 0; the index segments are padded to fixed capacities so that every batch of a run lands at the same device addresses (the
// step then replays one captured graph).  The reference does this work per row in Python (ntf.py:17-25: lil row -> dense float vector, default
// collate); here it is a memcpy per row of the already-CSR teamsvecs.
// Data-parallel ranks train on rows [lo, hi) of the global batch but sample negatives from the whole batch's members (unigram_b): the member
// CSR is packed for all n rows, the skill CSR for the slice only (the other rows get empty skill rows, so the offsets stay those of the batch).
extern "C" int ntf_pack_host_batch(const int32_t* rows, int n, const int32_t* s_indptr, const int32_t* s_indices, const int32_t* m_indptr,
                                   const int32_t* m_indices, int cap_s, int cap_m, int lo, int hi, int32_t* out, size_t out_words) {
  NTF_REQUIRE(rows && s_indptr && s_indices && m_indptr && m_indices && out && n > 0, NTF_ERR_BAD_ARG, "pack_host_batch: null pointer / empty batch");
  NTF_REQUIRE(0 <= lo && lo <= hi && hi <= n, NTF_ERR_BAD_ARG, "pack_host_batch: slice [%d,%d) of %d rows", lo, hi, n);
  const size_t need = 2 * ((size_t)n + 1) + 2 * (size_t)cap_s + (size_t)cap_m;
  NTF_REQUIRE(out_words >= need, NTF_ERR_WORKSPACE, "pack_host_batch: block of %zu words, %zu needed", out_words, need);
  int32_t* sp = out;
  int32_t* mp = sp + n + 1;
  int32_t* si = mp + n + 1;
  int32_t* sr = si + cap_s;
  int32_t* mi = sr + cap_s;
  int ns = 0, nm = 0;
  sp[0] = 0; mp[0] = 0;
  for (int i = 0; i < n; ++i) {
    const int r = rows[i];
    const int a = s_indptr[r], ls = (i >= lo && i < hi) ? s_indptr[r + 1] - a : 0, b = m_indptr[r], lm = m_indptr[r + 1] - b;
    NTF_REQUIRE(ns + ls <= cap_s && nm + lm <= cap_m, NTF_ERR_BAD_ARG, "pack_host_batch: capacities (%d, %d) too small at row %d", cap_s, cap_m, i);
    memcpy(si + ns, s_indices + a, (size_t)ls * sizeof(int32_t));
    for (int k = 0; k < ls; ++k) sr[ns + k] = i;
    memcpy(mi + nm, m_indices + b, (size_t)lm * sizeof(int32_t));
    ns += ls; nm += lm;
    sp[i + 1] = ns; mp[i + 1] = nm;
  }
  memset(si + ns, 0, (size_t)(cap_s - ns) * sizeof(int32_t));
  memset(sr + ns, 0, (size_t)(cap_s - ns) * sizeof(int32_t));
  memset(mi + nm, 0, (size_t)(cap_m - nm) * sizeof(int32_t));
  return NTF_OK;
}
