// staging.cu -- SURVEY 8(f-4): the sparse staging work either side of the training path, on the device.
//
//   ntf_csr_from_lists     team.py:148-173 (bucketing / Team.get_one_hot): a team's id list -> the sorted, duplicate-free row of the
//                          multi-hot matrix.  Warp bitonic sort for rows of <= 32 ids, a shared-memory BITMAP over the column range for
//                          longer ones (set the bits, popcount-scan the words, emit the set bits in order: no comparison sort at all).
//   ntf_cooccur_count/fill team.py:318-341 (gen_skill_coverage): member^T . skill, the [E,S] matrix of co-occurrence counts, without
//                          forming or sorting the (expert, skill) pairs: the member matrix is transposed by a counting sort (histogram,
//                          scan, scatter), then one CTA per expert ORs the skill rows of the expert's teams into a bitmap of the S skills;
//                          the popcount prefix of the bitmap is at once the row length, the sorted column list and the rank of every
//                          skill, so the counts are integer atomics on rank addresses.  Integer work, bit-exact, order-independent.
//   ntf_skill_coverage     metric.py:44-73 (calculate_skill_coverage): per team, the fraction of the required skills held by the union
//                          of the first k ranked experts; a warp per team, binary searches in the co-occurrence rows.
//
// All HBM/L2-bound integer work: no tensor cores here.  Algorithmic bytes: from_lists 8 B per id; cooccur
// 8 B per member entry for the transpose + 2 x 4 B per (team of expert, skill) pair + 8 B per output entry.
#include <stdlib.h>

#include "common.cuh"

int ntf_scan_u32_impl(cudaStream_t st, const uint32_t* in, size_t n, uint32_t* out, int inclusive, uint32_t* grand_total, void* ws, size_t ws_bytes);
size_t ntf_scan_workspace_bytes(size_t n);

namespace {
constexpr int ROW_THREADS = 256;

// exclusive scan of one value per thread over the CTA (ROW_THREADS threads); *total = the sum.  red: 4 + 1 words of shared memory
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, uint32_t* red, uint32_t* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();  // (red may still be read from a previous call)
  if (lane == 31) red[w] = incl;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int q = 0; q < ROW_THREADS / 32; ++q) {
    const uint32_t r = red[q];
    if (q < w) base += r;
    tot += r;
  }
  *total = tot;
  return base + incl - v;
}

// ---- bitmap of a column range in shared memory: the words [0, nw) ----
__device__ __forceinline__ void bitmap_zero(uint32_t* bits, int nw) {
  for (int i = threadIdx.x; i < nw; i += ROW_THREADS) bits[i] = 0u;
}
// every thread owns the contiguous words [w0, w1): returns the exclusive popcount prefix of its chunk, *total = bits set in the map
__device__ __forceinline__ uint32_t bitmap_chunk_prefix(const uint32_t* bits, int nw, int* w0, int* w1, uint32_t* red, uint32_t* total) {
  const int per = (nw + ROW_THREADS - 1) / ROW_THREADS;
  *w0 = min(nw, (int)threadIdx.x * per); *w1 = min(nw, *w0 + per);
  uint32_t c = 0;
  for (int i = *w0; i < *w1; ++i) c += __popc(bits[i]);
  return cta_exclusive_scan(c, red, total);
}

// =========================================================================================================
// ntf_csr_from_lists
// =========================================================================================================
// short rows (<= 32 ids): one warp per row, a bitonic network over the lanes; duplicates and the padding collapse under "differs from the
// lane before".  Longer rows are queued for the bitmap kernel.  out (the row's slots of the staging buffer) and len[row].
__global__ void __launch_bounds__(256) lists_short_kernel(int n, const int32_t* __restrict__ indptr, const int32_t* __restrict__ ids, int32_t* __restrict__ tmp,
                                                          uint32_t* __restrict__ len, int32_t* __restrict__ long_rows, int* __restrict__ n_long) {
  const int lane = threadIdx.x & 31;
  const int row = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (row >= n) return;
  const int beg = indptr[row], L = indptr[row + 1] - beg;
  if (L > 32) {
    if (lane == 0) long_rows[atomicAdd(n_long, 1)] = row;
    return;
  }
  uint32_t v = lane < L ? (uint32_t)ids[beg + lane] : 0xffffffffu;  // (ids are >= 0: the padding sorts last)
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint32_t o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool up = (lane & k) == 0, lower = (lane & j) == 0;
      v = (lower == up) ? min(v, o) : max(v, o);
    }
  const uint32_t prev = __shfl_up_sync(0xffffffffu, v, 1);
  const bool keep = v != 0xffffffffu && (lane == 0 || v != prev);
  const uint32_t m = __ballot_sync(0xffffffffu, keep);
  if (keep) tmp[beg + __popc(m & ((1u << lane) - 1u))] = (int32_t)v;
  if (lane == 0) len[row] = __popc(m);
}

// long rows: a CTA per queued row; the row's ids set bits of a shared-memory map of the n_cols columns, the set bits leave in order
__global__ void __launch_bounds__(ROW_THREADS) lists_long_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ ids, int n_cols,
                                                                  int32_t* __restrict__ tmp, uint32_t* __restrict__ len,
                                                                  const int32_t* __restrict__ long_rows, const int* __restrict__ n_long) {
  extern __shared__ uint32_t bits[];
  __shared__ uint32_t red[ROW_THREADS / 32];
  const int nw = (n_cols + 31) >> 5;
  const int nl = *n_long;
  for (int q = blockIdx.x; q < nl; q += gridDim.x) {
    const int row = long_rows[q], beg = indptr[row], L = indptr[row + 1] - beg;
    __syncthreads();
    bitmap_zero(bits, nw);
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += ROW_THREADS) {
      const int c = ids[beg + i];
      atomicOr(&bits[c >> 5], 1u << (c & 31));
    }
    __syncthreads();
    int w0, w1;
    uint32_t total;
    uint32_t pos = bitmap_chunk_prefix(bits, nw, &w0, &w1, red, &total);
    for (int i = w0; i < w1; ++i) {
      uint32_t b = bits[i];
      while (b) {
        const int t = __ffs(b) - 1;
        b &= b - 1;
        tmp[beg + pos++] = i * 32 + t;
      }
    }
    if (threadIdx.x == 0) len[row] = total;
  }
}

// dst row = the first len entries of the row's staging slots (warp per row)
__global__ void __launch_bounds__(256) lists_compact_kernel(int n, const int32_t* __restrict__ indptr, const int32_t* __restrict__ tmp,
                                                            const int32_t* __restrict__ dst_indptr, int32_t* __restrict__ dst_indices) {
  const int lane = threadIdx.x & 31;
  const int row = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (row >= n) return;
  const int src = indptr[row], dst = dst_indptr[row], L = dst_indptr[row + 1] - dst;
  for (int i = lane; i < L; i += 32) dst_indices[dst + i] = tmp[src + i];
}

__global__ void check_ids_kernel(size_t n, const int32_t* __restrict__ ids, int n_cols, int* __restrict__ bad) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if ((uint32_t)ids[i] >= (uint32_t)n_cols) *bad = 1;
}

struct ListsWs { size_t len, tmp, lrows, flags, scan, total; };
ListsWs lists_ws(int n, size_t n_ids) {
  ListsWs w;
  w.len = 0;
  w.tmp = w.len + align_up((size_t)(n + 1) * 4, 256);
  w.lrows = w.tmp + align_up((n_ids ? n_ids : 1) * 4, 256);
  w.flags = w.lrows + align_up((size_t)(n ? n : 1) * 4, 256);
  w.scan = w.flags + 256;
  w.total = w.scan + ntf_scan_workspace_bytes((size_t)n + 1);
  return w;
}
}  // namespace

extern "C" size_t ntf_csr_from_lists_workspace_bytes(int n, size_t n_ids) { return n < 0 ? 0 : lists_ws(n, n_ids).total; }

extern "C" int ntf_csr_from_lists(ntf_ctx* ctx, void* stream, int n, size_t n_ids, const int32_t* indptr, const int32_t* ids, int n_cols, int32_t* dst_indptr,
                                  int32_t* dst_indices, int64_t* nnz, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ctx && indptr && dst_indptr && nnz && (n_ids == 0 || (ids && dst_indices)), NTF_ERR_BAD_ARG, "csr_from_lists: null pointer");
  NTF_REQUIRE(n >= 0 && n_cols > 0, NTF_ERR_BAD_ARG, "csr_from_lists: n=%d n_cols=%d", n, n_cols);
  const ListsWs w = lists_ws(n, n_ids);
  NTF_REQUIRE(workspace && workspace_bytes >= w.total, NTF_ERR_WORKSPACE, "csr_from_lists: workspace %zu < %zu", workspace_bytes, w.total);
  const size_t smem = (size_t)((n_cols + 31) / 32) * 4;
  NTF_REQUIRE(smem <= ctx->smem_optin - 1024, NTF_ERR_UNSUPPORTED, "csr_from_lists: %d columns need a %zu-byte bitmap (limit %zu)", n_cols, smem, ctx->smem_optin - 1024);
  cudaStream_t st = as_stream(stream);
  char* ws = (char*)workspace;
  uint32_t* len = (uint32_t*)(ws + w.len);
  int32_t* tmp = (int32_t*)(ws + w.tmp);
  int32_t* lrows = (int32_t*)(ws + w.lrows);
  int* flags = (int*)(ws + w.flags);  // [0] long rows queued, [1] an id out of range
  *nnz = 0;
  NTF_CUDA(cudaMemsetAsync(dst_indptr, 0, 4, st));
  if (n == 0) return NTF_OK;
  NTF_CUDA(cudaMemsetAsync(flags, 0, 8, st));
  NTF_CUDA(cudaMemsetAsync(len + n, 0, 4, st));
  if (n_ids) {
    const size_t cb = (n_ids + 255) / 256;
    NTF_COUNT_LAUNCH; check_ids_kernel<<<(unsigned)(cb < (size_t)ctx->sm_count * 8 ? cb : (size_t)ctx->sm_count * 8), 256, 0, st>>>(n_ids, ids, n_cols, flags + 1);
  }
  int bad = 0;
  NTF_CUDA(cudaMemcpyAsync(&bad, flags + 1, 4, cudaMemcpyDeviceToHost, st));
  NTF_CUDA(cudaStreamSynchronize(st));
  NTF_REQUIRE(!bad, NTF_ERR_BAD_ARG, "csr_from_lists: an id outside [0, %d)", n_cols);
  NTF_COUNT_LAUNCH; lists_short_kernel<<<cdiv(n, 8), 256, 0, st>>>(n, indptr, ids, tmp, len, lrows, flags);
  NTF_CUDA(cudaFuncSetAttribute(lists_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NTF_COUNT_LAUNCH; lists_long_kernel<<<ctx->sm_count * 4, ROW_THREADS, smem, st>>>(indptr, ids, n_cols, tmp, len, lrows, flags);
  NTF_LAUNCH_CHECK();
  uint32_t* total = (uint32_t*)(flags + 2);
  int rc = ntf_scan_u32_impl(st, len, (size_t)n + 1, (uint32_t*)dst_indptr, 0, total, ws + w.scan, workspace_bytes - w.scan);
  if (rc) return rc;
  NTF_COUNT_LAUNCH; lists_compact_kernel<<<cdiv(n, 8), 256, 0, st>>>(n, indptr, tmp, dst_indptr, dst_indices);
  NTF_LAUNCH_CHECK();
  uint32_t tot = 0;
  NTF_CUDA(cudaMemcpyAsync(&tot, total, 4, cudaMemcpyDeviceToHost, st));
  NTF_CUDA(cudaStreamSynchronize(st));
  *nnz = (int64_t)tot;
  return NTF_OK;
}

// =========================================================================================================
// ntf_cooccur_count / ntf_cooccur_fill
// =========================================================================================================
namespace {
// warp per team: the team's members bump their experts' counters (COUNT) or take a slot of the expert's team list (SCATTER)
template <bool SCATTER>
__global__ void __launch_bounds__(256) transpose_kernel(int T, const int32_t* __restrict__ m_indptr, const int32_t* __restrict__ m_indices,
                                                        const uint8_t* __restrict__ skip, uint32_t* __restrict__ cnt, const uint32_t* __restrict__ et_ptr,
                                                        int32_t* __restrict__ et_teams) {
  const int lane = threadIdx.x & 31;
  const int t = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (t >= T || (skip && skip[t])) return;
  for (int p = m_indptr[t] + lane; p < m_indptr[t + 1]; p += 32) {
    const int e = m_indices[p];
    const uint32_t slot = atomicAdd(&cnt[e], 1u);
    if (SCATTER) et_teams[et_ptr[e] + slot] = t;
  }
}

// CTA per expert: OR the skill rows of the expert's teams into the bitmap of the S skills.  FILL = false: the row length (bits set).
// FILL = true: the set bits in order are the row's columns; a second sweep over the same skill rows adds 1 at the rank of each skill.
template <bool FILL>
__global__ void __launch_bounds__(ROW_THREADS) cooccur_row_kernel(int E, int S, const uint32_t* __restrict__ et_ptr, const int32_t* __restrict__ et_teams,
                                                                  const int32_t* __restrict__ s_indptr, const int32_t* __restrict__ s_indices,
                                                                  uint32_t* __restrict__ len, const int32_t* __restrict__ co_indptr,
                                                                  int32_t* __restrict__ co_indices, int32_t* __restrict__ co_values) {
  extern __shared__ uint32_t sm[];
  __shared__ uint32_t red[ROW_THREADS / 32];
  const int nw = (S + 31) >> 5;
  uint32_t* bits = sm;
  uint32_t* pre = sm + nw;  // FILL: exclusive popcount prefix per word
  for (int e = blockIdx.x; e < E; e += gridDim.x) {
    const uint32_t t0 = et_ptr[e], t1 = et_ptr[e + 1];
    if (t0 == t1) {
      if (!FILL && threadIdx.x == 0) len[e] = 0u;
      continue;
    }
    __syncthreads();
    bitmap_zero(bits, nw);
    __syncthreads();
    // a THREAD per team of the expert (a team has a handful of skills; a prolific expert has thousands of teams: what counts is how many of the
    // dependent load chains team -> row pointers -> skills are in flight)
    for (uint32_t q = t0 + threadIdx.x; q < t1; q += ROW_THREADS) {
      const int t = et_teams[q];
      for (int p = s_indptr[t], pe = s_indptr[t + 1]; p < pe; ++p) {
        const int s = s_indices[p];
        atomicOr(&bits[s >> 5], 1u << (s & 31));
      }
    }
    __syncthreads();
    int w0, w1;
    uint32_t total;
    uint32_t pos = bitmap_chunk_prefix(bits, nw, &w0, &w1, red, &total);
    if (!FILL) {
      if (threadIdx.x == 0) len[e] = total;
      continue;
    }
    const int base = co_indptr[e];
    for (int i = w0; i < w1; ++i) {
      uint32_t b = bits[i];
      pre[i] = pos;
      while (b) {
        const int t = __ffs(b) - 1;
        b &= b - 1;
        co_indices[base + pos] = i * 32 + t;
        co_values[base + pos] = 0;
        ++pos;
      }
    }
    __syncthreads();  // (block-wide: makes the zeroed values visible to the atomics below, global memory included)
    for (uint32_t q = t0 + threadIdx.x; q < t1; q += ROW_THREADS) {
      const int t = et_teams[q];
      for (int p = s_indptr[t], pe = s_indptr[t + 1]; p < pe; ++p) {
        const int s = s_indices[p];
        const uint32_t rank = pre[s >> 5] + __popc(bits[s >> 5] & ((1u << (s & 31)) - 1u));
        atomicAdd(&co_values[base + rank], 1);
      }
    }
  }
}

struct CoWs { size_t cnt, ptr, teams, len, flags, scan, total; };
CoWs co_ws(int E, size_t nnz_member) {
  CoWs w;
  w.cnt = 0;
  w.ptr = w.cnt + align_up((size_t)(E + 1) * 4, 256);
  w.teams = w.ptr + align_up((size_t)(E + 1) * 4, 256);
  w.len = w.teams + align_up((nnz_member ? nnz_member : 1) * 4, 256);
  w.flags = w.len + align_up((size_t)(E + 1) * 4, 256);
  w.scan = w.flags + 256;
  w.total = w.scan + ntf_scan_workspace_bytes((size_t)E + 1);
  return w;
}

int co_check(ntf_ctx* ctx, int T, int E, int S, const void* a, const void* b, const void* c, const void* d, void* ws, size_t ws_bytes, size_t nnz_member, size_t* smem,
             int fill) {
  NTF_REQUIRE(ctx && a && b && c && d, NTF_ERR_BAD_ARG, "cooccur: null pointer");
  NTF_REQUIRE(T >= 0 && E > 0 && S > 0, NTF_ERR_BAD_ARG, "cooccur: T=%d E=%d S=%d", T, E, S);
  NTF_REQUIRE(ws && ws_bytes >= co_ws(E, nnz_member).total, NTF_ERR_WORKSPACE, "cooccur: workspace %zu < %zu", ws_bytes, co_ws(E, nnz_member).total);
  *smem = (size_t)((S + 31) / 32) * 4 * (fill ? 2 : 1);
  NTF_REQUIRE(*smem <= ctx->smem_optin - 1024, NTF_ERR_UNSUPPORTED, "cooccur: %d skills need %zu bytes of shared memory (limit %zu)", S, *smem, ctx->smem_optin - 1024);
  return NTF_OK;
}
}  // namespace

extern "C" size_t ntf_cooccur_workspace_bytes(int E, size_t nnz_member) { return E <= 0 ? 0 : co_ws(E, nnz_member).total; }

extern "C" int ntf_cooccur_count(ntf_ctx* ctx, void* stream, int T, int E, int S, size_t nnz_member, const int32_t* m_indptr, const int32_t* m_indices,
                                 const int32_t* s_indptr, const int32_t* s_indices, const uint8_t* skip, int32_t* co_indptr, int64_t* nnz, void* workspace,
                                 size_t workspace_bytes) {
  size_t smem;
  int rc = co_check(ctx, T, E, S, m_indptr, s_indptr, co_indptr, nnz, workspace, workspace_bytes, nnz_member, &smem, 0);
  if (rc) return rc;
  NTF_REQUIRE((nnz_member == 0 || m_indices) && s_indices, NTF_ERR_BAD_ARG, "cooccur_count: null pointer");
  const CoWs w = co_ws(E, nnz_member);
  cudaStream_t st = as_stream(stream);
  char* ws = (char*)workspace;
  uint32_t* cnt = (uint32_t*)(ws + w.cnt);
  uint32_t* ptr = (uint32_t*)(ws + w.ptr);
  int32_t* teams = (int32_t*)(ws + w.teams);
  uint32_t* len = (uint32_t*)(ws + w.len);
  uint32_t* total = (uint32_t*)(ws + w.flags);
  // member^T by a counting sort: expert -> the teams it is a member of (order inside a list: arrival; nothing below depends on it)
  NTF_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(E + 1) * 4, st));
  if (T > 0) { NTF_COUNT_LAUNCH; transpose_kernel<false><<<cdiv(T, 8), 256, 0, st>>>(T, m_indptr, m_indices, skip, cnt, nullptr, nullptr); }
  if ((rc = ntf_scan_u32_impl(st, cnt, (size_t)E + 1, ptr, 0, nullptr, ws + w.scan, workspace_bytes - w.scan))) return rc;
  NTF_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(E + 1) * 4, st));
  if (T > 0) { NTF_COUNT_LAUNCH; transpose_kernel<true><<<cdiv(T, 8), 256, 0, st>>>(T, m_indptr, m_indices, skip, cnt, ptr, teams); }
  NTF_CUDA(cudaFuncSetAttribute(cooccur_row_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NTF_CUDA(cudaMemsetAsync(len + E, 0, 4, st));
  NTF_COUNT_LAUNCH; cooccur_row_kernel<false><<<min(E, ctx->sm_count * 8), ROW_THREADS, smem, st>>>(E, S, ptr, teams, s_indptr, s_indices, len, nullptr, nullptr, nullptr);
  NTF_LAUNCH_CHECK();
  if ((rc = ntf_scan_u32_impl(st, len, (size_t)E + 1, (uint32_t*)co_indptr, 0, total, ws + w.scan, workspace_bytes - w.scan))) return rc;
  uint32_t tot = 0;
  NTF_CUDA(cudaMemcpyAsync(&tot, total, 4, cudaMemcpyDeviceToHost, st));
  NTF_CUDA(cudaStreamSynchronize(st));
  NTF_REQUIRE(tot <= 0x7fffffffu, NTF_ERR_UNSUPPORTED, "cooccur_count: %u entries overflow the int32 row pointers", tot);
  *nnz = (int64_t)tot;
  return NTF_OK;
}

extern "C" int ntf_cooccur_fill(ntf_ctx* ctx, void* stream, int T, int E, int S, size_t nnz_member, const int32_t* s_indptr, const int32_t* s_indices,
                                const int32_t* co_indptr, int32_t* co_indices, int32_t* co_values, void* workspace, size_t workspace_bytes) {
  size_t smem;
  int rc = co_check(ctx, T, E, S, s_indptr, s_indices, co_indptr, co_indptr, workspace, workspace_bytes, nnz_member, &smem, 1);
  if (rc) return rc;
  const CoWs w = co_ws(E, nnz_member);
  char* ws = (char*)workspace;
  NTF_CUDA(cudaFuncSetAttribute(cooccur_row_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NTF_COUNT_LAUNCH;
  cooccur_row_kernel<true><<<min(E, ctx->sm_count * 8), ROW_THREADS, smem, as_stream(stream)>>>(E, S, (const uint32_t*)(ws + w.ptr), (const int32_t*)(ws + w.teams), s_indptr,
                                                                                               s_indices, nullptr, co_indptr, co_indices, co_values);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

// =========================================================================================================
// ntf_skill_coverage
// =========================================================================================================
namespace {
constexpr int COV_MAXK = 8;  // cut-offs per call
struct CovKs { int k[COV_MAXK]; };

// warp per team.  Lane i takes required skill i (+32, ...): the rank of the first recommended expert that holds it (binary search in the
// expert's sorted row of the co-occurrence matrix); the cut-off k covers the skill when that rank is < k.
__global__ void __launch_bounds__(256) skill_coverage_kernel(int n, int K, const int32_t* __restrict__ idx, const int32_t* __restrict__ x_indptr,
                                                             const int32_t* __restrict__ x_indices, const int32_t* __restrict__ co_indptr,
                                                             const int32_t* __restrict__ co_indices, int nk, int kmax, CovKs ksv,
                                                             double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int t = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (t >= n) return;
  const int* ks = ksv.k;
  const int beg = x_indptr[t], L = x_indptr[t + 1] - beg;
  int covered[COV_MAXK];
#pragma unroll
  for (int j = 0; j < COV_MAXK; ++j) covered[j] = 0;
  const int depth = min(K, kmax);
  for (int i0 = 0; i0 < L; i0 += 32) {
    const int i = i0 + lane;
    int first = 0x7fffffff;
    if (i < L) {
      const int s = x_indices[beg + i];
      for (int r = 0; r < depth; ++r) {
        const int e = idx[(size_t)t * K + r];
        if (e < 0) continue;
        int lo = co_indptr[e], hi = co_indptr[e + 1];
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (co_indices[mid] < s) lo = mid + 1; else hi = mid;
        }
        if (lo < co_indptr[e + 1] && co_indices[lo] == s) { first = r; break; }
      }
    }
#pragma unroll
    for (int j = 0; j < COV_MAXK; ++j) covered[j] += __popc(__ballot_sync(0xffffffffu, first < ks[j]));
  }
  if (lane == 0)
    for (int j = 0; j < nk; ++j) out[(size_t)t * nk + j] = (double)covered[j] / (double)L;  // (L = 0: 0/0 = nan, as numpy's division gives)
}
}  // namespace

extern "C" int ntf_skill_coverage(ntf_ctx* ctx, void* stream, int n, int K, const int32_t* idx, const int32_t* x_indptr, const int32_t* x_indices,
                                  const int32_t* co_indptr, const int32_t* co_indices, const int* ks, int nk, double* out) {
  NTF_REQUIRE(ctx && idx && x_indptr && x_indices && co_indptr && co_indices && ks && out, NTF_ERR_BAD_ARG, "skill_coverage: null pointer");
  NTF_REQUIRE(n >= 0 && K >= 1 && nk >= 1 && nk <= COV_MAXK, NTF_ERR_BAD_ARG, "skill_coverage: n=%d K=%d nk=%d (1..%d cut-offs)", n, K, nk, COV_MAXK);
  if (n == 0) return NTF_OK;
  CovKs kv = {};  // ks: a HOST array of nk cut-offs (any order); the deepest one bounds the searches
  int kmax = 0;
  for (int j = 0; j < nk; ++j) {
    NTF_REQUIRE(ks[j] >= 1, NTF_ERR_BAD_ARG, "skill_coverage: cut-off %d", ks[j]);
    kv.k[j] = ks[j];
    kmax = ks[j] > kmax ? ks[j] : kmax;
  }
  NTF_COUNT_LAUNCH; skill_coverage_kernel<<<cdiv(n, 8), 256, 0, as_stream(stream)>>>(n, K, idx, x_indptr, x_indices, co_indptr, co_indices, nk, kmax, kv, out);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
