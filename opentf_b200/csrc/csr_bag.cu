// csr_bag.cu -- the sparse multi-hot input layer (K1/K2 of SURVEY.md) and CSR staging.
//
//   ntf_scan_u32      : device-wide inclusive/exclusive prefix sum (two-level, deterministic)
//   ntf_csr_gather    : permuted/compacted copy of CSR rows (the DataLoader shuffle, fnn.py:95-97, on device)
//   ntf_csr_bag_fwd   : A = lrelu(b0 + sum_s W0T[s,:])                    -- HBM-bound gather-sum
//   ntf_csr_bag_bwd   : dW0T[s,:] = sum_{n: s in skills(n)} dZ[n,:]       -- atomic-free, owner-computes
//   ntf_act_bwd       : dZ = dY*lrelu'(Y), db = colsum(dZ)
#include <cuda_fp16.h>

#include <stdlib.h>

#include "common.cuh"

// =========================================================================================================
// prefix sum: block-local scan + one-block scan of the block totals + offset add
// =========================================================================================================
namespace {
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  // returns the exclusive prefix of v over the block; *total (all threads) = block sum
  __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t t = lane < SCAN_THREADS / 32 ? warp_tot[lane] : 0u, ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    if (lane < SCAN_THREADS / 32) warp_tot[lane] = ti - t;  // exclusive warp offsets
    if (lane == SCAN_THREADS / 32 - 1) *total = ti;
  }
  __syncthreads();
  const uint32_t r = warp_tot[w] + inc - v;
  __syncthreads();
  return r;
}

// pass 1: per-tile totals.  in may be a length array or (if from_indptr) derived as indptr[rows[i]+1]-indptr[rows[i]]
__global__ void scan_tile_totals(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ tile_tot) {
  __shared__ uint32_t tot;
  const size_t base = (size_t)blockIdx.x * SCAN_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const size_t k = base + (size_t)i * SCAN_THREADS + threadIdx.x;
    if (k < n) s += in[k];
  }
  (void)block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) tile_tot[blockIdx.x] = tot;
}

// pass 2: one block turns tile totals into exclusive tile offsets (serial over chunks of SCAN_THREADS)
__global__ void scan_tile_offsets(uint32_t* __restrict__ tile_tot, int ntiles) {
  __shared__ uint32_t tot;
  uint32_t carry = 0;
  for (int base = 0; base < ntiles; base += SCAN_THREADS) {
    const int k = base + threadIdx.x;
    const uint32_t v = k < ntiles ? tile_tot[k] : 0u;
    const uint32_t ex = block_exclusive_scan(v, &tot);
    if (k < ntiles) tile_tot[k] = carry + ex;
    carry += tot;
    __syncthreads();
  }
}

// pass 3: scan inside each tile (thread owns SCAN_ITEMS consecutive items) + tile offset
__global__ void scan_apply(const uint32_t* __restrict__ in, size_t n, const uint32_t* __restrict__ tile_off,
                           uint32_t* __restrict__ out, int inclusive, uint32_t* __restrict__ grand_total) {
  __shared__ uint32_t tot;
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = base + i < n ? in[base + i] : 0u;
    s += v[i];
  }
  uint32_t run = block_exclusive_scan(s, &tot) + tile_off[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = inclusive ? run + v[i] : run;
    run += v[i];
  }
  if (grand_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *grand_total = run;
}
}  // namespace

size_t ntf_scan_workspace_bytes(size_t n) { return align_up(((n + SCAN_TILE - 1) / SCAN_TILE + 1) * sizeof(uint32_t), 256); }

// out may alias in.  grand_total (device, nullable) receives the sum of all items.
int ntf_scan_u32_impl(cudaStream_t st, const uint32_t* in, size_t n, uint32_t* out, int inclusive, uint32_t* grand_total,
                      void* ws, size_t ws_bytes) {
  if (n == 0) {
    if (grand_total) NTF_CUDA(cudaMemsetAsync(grand_total, 0, sizeof(uint32_t), st));
    return NTF_OK;
  }
  NTF_REQUIRE(ws_bytes >= ntf_scan_workspace_bytes(n), NTF_ERR_WORKSPACE, "scan: workspace %zu < %zu", ws_bytes,
              ntf_scan_workspace_bytes(n));
  const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  uint32_t* tile = (uint32_t*)ws;
  NTF_COUNT_LAUNCH; scan_tile_totals<<<ntiles, SCAN_THREADS, 0, st>>>(in, n, tile);
  NTF_COUNT_LAUNCH; scan_tile_offsets<<<1, SCAN_THREADS, 0, st>>>(tile, ntiles);
  NTF_COUNT_LAUNCH; scan_apply<<<ntiles, SCAN_THREADS, 0, st>>>(in, n, tile, out, inclusive, grand_total);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

// =========================================================================================================
// CSR gather: dst row i = src row rows[i]
// =========================================================================================================
namespace {
__global__ void csr_row_lengths(const int32_t* __restrict__ rows, int n, const int32_t* __restrict__ indptr,
                                uint32_t* __restrict__ len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int r = rows ? rows[i] : i;
    len[i] = (uint32_t)(indptr[r + 1] - indptr[r]);
  }
}

// one warp per destination row; also records the destination row of every entry (ent_row)
__global__ void csr_copy_rows(const int32_t* __restrict__ rows, int n, const int32_t* __restrict__ src_indptr,
                              const int32_t* __restrict__ src_indices, int32_t* __restrict__ dst_indptr,
                              int32_t* __restrict__ dst_indices, int32_t* __restrict__ dst_ent_row) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int r = rows ? rows[warp] : warp;
  const int sb = src_indptr[r], len = src_indptr[r + 1] - sb;
  const int db = dst_indptr[warp];
  for (int k = lane; k < len; k += 32) {
    dst_indices[db + k] = src_indices[sb + k];
    if (dst_ent_row) dst_ent_row[db + k] = warp;
  }
  if (warp == n - 1 && lane == 0) dst_indptr[n] = db + len;
}
}  // namespace

extern "C" size_t ntf_csr_gather_workspace_bytes(int n) { return ntf_scan_workspace_bytes((size_t)n); }

extern "C" int ntf_csr_gather(ntf_ctx* ctx, void* stream, const int32_t* rows, int n, const int32_t* src_indptr,
                              const int32_t* src_indices, int32_t* dst_indptr, int32_t* dst_indices,
                              int32_t* dst_ent_row, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ctx && src_indptr && src_indices && dst_indptr && dst_indices, NTF_ERR_BAD_ARG, "csr_gather: null pointer");
  NTF_REQUIRE(n >= 0, NTF_ERR_BAD_ARG, "csr_gather: n=%d", n);
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    NTF_CUDA(cudaMemsetAsync(dst_indptr, 0, sizeof(int32_t), st));
    return NTF_OK;
  }
  // lengths are staged in dst_indptr[0..n) and scanned in place (exclusive)
  NTF_COUNT_LAUNCH; csr_row_lengths<<<cdiv(n, 256), 256, 0, st>>>(rows, n, src_indptr, (uint32_t*)dst_indptr);
  int rc = ntf_scan_u32_impl(st, (const uint32_t*)dst_indptr, (size_t)n, (uint32_t*)dst_indptr, 0, nullptr, workspace,
                             workspace_bytes);
  if (rc) return rc;
  NTF_COUNT_LAUNCH; csr_copy_rows<<<cdiv(n * 32, 256), 256, 0, st>>>(rows, n, src_indptr, src_indices, dst_indptr, dst_indices, dst_ent_row);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

// ---- dense (embedded) skill input: ntf.py:24 -- a team's row is a d-vector instead of a multi-hot set ----------------------------
// dst[i,:] = src[rows[i],:]: the per-epoch shuffle of fnn.py:95 for dense rows (HBM-bound: 8*d bytes per team).
namespace {
__global__ void __launch_bounds__(256) rows_gather_kernel(const int32_t* __restrict__ rows, int n, int d, const float* __restrict__ src,
                                                          float* __restrict__ dst) {
  const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const size_t r = rows ? (size_t)rows[i] : (size_t)i;
    const float* s = src + r * d;
    float* t = dst + (size_t)i * d;
    if ((d & 3) == 0) {
      for (int c = lane; c < (d >> 2); c += 32) reinterpret_cast<float4*>(t)[c] = __ldg(reinterpret_cast<const float4*>(s) + c);
    } else {
      for (int c = lane; c < d; c += 32) t[c] = __ldg(s + c);
    }
  }
}
}  // namespace

extern "C" int ntf_rows_gather(ntf_ctx* ctx, void* stream, const int32_t* rows, int n, int d, const float* src, float* dst) {
  NTF_REQUIRE(ctx && src && dst, NTF_ERR_BAD_ARG, "rows_gather: null pointer");
  NTF_REQUIRE(n >= 0 && d > 0, NTF_ERR_BAD_ARG, "rows_gather: n=%d d=%d", n, d);
  if (n == 0) return NTF_OK;
  const int blocks = min(cdiv(n * 32, 256), ctx->sm_count * 8);
  NTF_COUNT_LAUNCH; rows_gather_kernel<<<blocks, 256, 0, as_stream(stream)>>>(rows, n, d, src, dst);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

// =========================================================================================================
// K1: embedding-bag forward.  One warp per team; a lane owns float4 slices of the h-vector.
// Algorithmic bytes per team: n_s*(4h+4) + 8 + 4h   
// =========================================================================================================
namespace {
template <bool VEC4>
__global__ void __launch_bounds__(256) csr_bag_fwd_kernel(int B, const int32_t* __restrict__ indptr,
                                                          const int32_t* __restrict__ indices,
                                                          const float* __restrict__ W0T, const float* __restrict__ b0,
                                                          int h, float* __restrict__ A, __half* __restrict__ A16) {
  // programmatic dependent launch (infer_topk.cu): the kernel chained behind this one may set itself up meanwhile; launched `chained`, this
  // one starts while the kernel before it drains and waits here for its memory (no-ops in a plain launch)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < B; n += warps) {
    const int beg = indptr[n], end = indptr[n + 1];
    if (VEC4) {
      const int h4 = h >> 2;
      for (int c0 = 0; c0 < h4; c0 += 32) {  // (the shuffles need every lane: the column guard is on the memory accesses only)
        const int c = min(c0 + lane, h4 - 1);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int p = beg;
        for (; p + 4 <= end; p += 4) {  // 4 independent 16-byte loads in flight, summed in CSR order
          const int s0 = indices[p], s1 = indices[p + 1], s2 = indices[p + 2], s3 = indices[p + 3];
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(W0T + (size_t)s0 * h) + c);
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(W0T + (size_t)s1 * h) + c);
          const float4 v2 = __ldg(reinterpret_cast<const float4*>(W0T + (size_t)s2 * h) + c);
          const float4 v3 = __ldg(reinterpret_cast<const float4*>(W0T + (size_t)s3 * h) + c);
          acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
          acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
          acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
          acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
        }
        for (; p < end; ++p) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(W0T + (size_t)indices[p] * h) + c);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b0) + c);
        float4 o;
        o.x = lrelu(acc.x + bb.x); o.y = lrelu(acc.y + bb.y); o.z = lrelu(acc.z + bb.z); o.w = lrelu(acc.w + bb.w);
        reinterpret_cast<float4*>(A + (size_t)n * h)[c] = o;
        if (A16) {  // the tensor-core output layer reads fp16 operands: hand it the copy now instead of a conversion pass later
          const __half2 lo = __floats2half2_rn(o.x, o.y), hi = __floats2half2_rn(o.z, o.w);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
          reinterpret_cast<uint2*>(A16 + (size_t)n * h)[c] = pk;
        }
      }
    } else {
      for (int c = lane; c < h; c += 32) {
        float acc = 0.f;
        for (int p = beg; p < end; ++p) acc += __ldg(W0T + (size_t)indices[p] * h + c);
        const float o = lrelu(acc + b0[c]);
        A[(size_t)n * h + c] = o;
        if (A16) A16[(size_t)n * h + c] = __float2half_rn(o);
      }
    }
  }
}
}  // namespace

// A16 (nullable): also write the activations as fp16 [B,h] (operand of the tensor-core output layer); chained: launch with the programmatic
// stream-serialization attribute (the test-time chain of infer_topk.cu, one call per batch back to back)
int ntf_csr_bag_fwd_impl(ntf_ctx* ctx, void* stream, int B, const int32_t* indptr, const int32_t* indices, const float* W0T,
                         const float* b0, int S, int h, float* A, void* A16v, bool chained) {
  __half* A16 = (__half*)A16v;
  NTF_REQUIRE(ctx && indptr && indices && W0T && b0 && A, NTF_ERR_BAD_ARG, "csr_bag_fwd: null pointer");
  NTF_REQUIRE(B >= 0 && S > 0 && h > 0, NTF_ERR_BAD_ARG, "csr_bag_fwd: B=%d S=%d h=%d", B, S, h);
  if (B == 0) return NTF_OK;
  const int blocks = min(cdiv(B, 8), ctx->sm_count * 8);
  const bool vec = (h % 4 == 0) && (((uintptr_t)W0T | (uintptr_t)b0 | (uintptr_t)A) % 16 == 0) && ((uintptr_t)A16 % 8 == 0);
  NTF_COUNT_LAUNCH;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = blocks; cfg.blockDim = 256; cfg.stream = as_stream(stream);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = chained ? 1 : 0;
  NTF_CUDA(cudaLaunchKernelEx(&cfg, vec ? csr_bag_fwd_kernel<true> : csr_bag_fwd_kernel<false>, B, indptr, indices, W0T, b0, h, A, A16));
  return NTF_OK;
}

extern "C" int ntf_csr_bag_fwd(ntf_ctx* ctx, void* stream, int B, const int32_t* indptr, const int32_t* indices,
                               const float* W0T, const float* b0, int S, int h, float* A) {
  return ntf_csr_bag_fwd_impl(ctx, stream, B, indptr, indices, W0T, b0, S, h, A, nullptr, false);
}

// Bnn / Flipout input layer (bayesian-torch LinearFlipout on a multi-hot row):
//   A[n,:] = lrelu( b_mu + sum_p W_mu[s_p,:] + (b_delta + sum_p sgn_p * W_delta[s_p,:]) * s_out[n,:] )
// sgn_p = sign_in[n, s_p] exists only at the nnz positions (x*sign = 0 elsewhere): one bit per CSR entry of the batch
// (bit p - indptr[0] of ent_sign, 1 -> -1).  One warp per team, lanes stride the h-vector (128-byte coalesced rows).
namespace {
__global__ void __launch_bounds__(256) csr_bag_flipout_fwd_kernel(int B, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                                  const uint32_t* __restrict__ ent_sign, const float* __restrict__ Wmu,
                                                                  const float* __restrict__ bmu, const float* __restrict__ Wd,
                                                                  const float* __restrict__ bd, const uint32_t* __restrict__ sign_out,
                                                                  int pitch, int h, float* __restrict__ A) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int p0 = indptr[0];
  for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < B; n += warps) {
    const int beg = indptr[n], end = indptr[n + 1];
    for (int c = lane; c < h; c += 32) {
      float am = 0.f, ad = 0.f;
      for (int p = beg; p < end; ++p) {
        const int sk = indices[p];
        const int q = p - p0;
        const bool neg = (__ldg(ent_sign + (q >> 5)) >> (q & 31)) & 1u;
        const float wd = __ldg(Wd + (size_t)sk * h + c);
        am += __ldg(Wmu + (size_t)sk * h + c);
        ad += neg ? -wd : wd;
      }
      const bool so = (__ldg(sign_out + (size_t)n * pitch + (c >> 5)) >> (c & 31)) & 1u;
      const float t = ad + bd[c];
      A[(size_t)n * h + c] = lrelu((am + bmu[c]) + (so ? -t : t));
    }
  }
}
}  // namespace

extern "C" int ntf_csr_bag_flipout_fwd(ntf_ctx* ctx, void* stream, int B, const int32_t* indptr, const int32_t* indices,
                                       const uint32_t* ent_sign, const float* W0T_mu, const float* b_mu, const float* W0T_delta,
                                       const float* b_delta, const uint32_t* sign_out, int pitch_words, int S, int h, float* A) {
  NTF_REQUIRE(ctx && indptr && indices && ent_sign && W0T_mu && b_mu && W0T_delta && b_delta && sign_out && A, NTF_ERR_BAD_ARG,
              "csr_bag_flipout_fwd: null pointer");
  NTF_REQUIRE(B >= 0 && S > 0 && h > 0 && pitch_words * 32 >= h, NTF_ERR_BAD_ARG, "csr_bag_flipout_fwd: B=%d S=%d h=%d pitch=%d", B, S, h, pitch_words);
  if (B == 0) return NTF_OK;
  NTF_COUNT_LAUNCH;
  csr_bag_flipout_fwd_kernel<<<min(cdiv(B, 8), ctx->sm_count * 8), 256, 0, as_stream(stream)>>>(B, indptr, indices, ent_sign, W0T_mu, b_mu, W0T_delta,
                                                                                                 b_delta, sign_out, pitch_words, h, A);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

// =========================================================================================================
// K2: embedding-bag backward, atomic-free and run-to-run deterministic.
//   dW0T[s,:] = sum over the batch entries (s, n) of dZ[n,:], written exactly once for EVERY s (zeros for skills
//   absent from the batch: Adam is dense, SURVEY.md "Hard parts").
// Skill popularity is heavy-tailed (a few skills sit in most teams), so the work is split by the per-batch entry count of a skill:
//   cold skills (<= HOT entries): pass 1 drops every batch entry into a slot of its skill (integer atomics: count and position);
//     pass 2 gives each skill a warp, which rank-sorts the <= 32 row ids and adds the dZ rows in ascending team order, 4 rows
//     in flight; skills absent from the batch get their zero row from the same warp;
//   hot skills (> HOT entries): one CTA per skill compacts that skill's entries in entry order, its 8 warps sum
//     fixed slices of the list and the 8 partials are combined in a fixed order.
// The summation order depends on the data only.  Algorithmic bytes per team: n_s*(4h+4) + 4h read, 4h*S/B written.
// =========================================================================================================
namespace {
constexpr int BWD_WARPS = 8;
constexpr int HOT = 32;      // a skill with more batch entries than this is reduced by the per-skill CTA kernel
constexpr int HCAP = 8192;
constexpr int HOT_G = 8;      // team ranges a hot skill's sum is cut into
constexpr int HOT_MAX = 1024; // hot skills that get the split treatment (a batch has at most entries/33 hot skills)

// pass 1: every batch entry (team n, skill s) takes the next slot of its skill: slots[s][pos] = batch row (| sign bit for Flipout).
// The slot order is arbitrary (integer atomics), pass 2 sorts it; cnt[s] ends as the exact number of entries of skill s.
__global__ void bag_bwd_fill_kernel(int B, const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                    const int32_t* __restrict__ ent_row, int row_base, const uint32_t* __restrict__ ent_sign,
                                    uint32_t* __restrict__ cnt, int32_t* __restrict__ slots) {
  const int p0 = indptr[0], p1 = indptr[B];
  for (int p = p0 + blockIdx.x * blockDim.x + threadIdx.x; p < p1; p += gridDim.x * blockDim.x) {
    const int s = __ldg(indices + p);
    int rr = __ldg(ent_row + p) - row_base;
    if (ent_sign) {
      const int q = p - p0;
      if ((__ldg(ent_sign + (q >> 5)) >> (q & 31)) & 1u) rr |= (int)0x80000000u;
    }
    const uint32_t pos = atomicAdd(cnt + s, 1u);
    if (pos < HOT) slots[(size_t)s * HOT + pos] = rr;
  }
}

// pass 2: one warp per skill (all S of them: Adam is dense, absent skills get an explicit zero row).  The <= 32 batch rows of the
// skill are rank-sorted by row id, so the fp32 sum runs in ascending team order whatever order pass 1 filled the slots in:
// no floating-point atomics, bit-reproducible.  Skills with more entries go to the hot list.
// skills with more than HOT entries in the batch -> the hot list.  Needs the counts only, so it runs with the slot fill (off the
// critical path of a step); list order is irrelevant: every hot skill is reduced on its own.
__global__ void bag_bwd_hotlist_kernel(int S, const uint32_t* __restrict__ cnt, int32_t* __restrict__ hot, uint32_t* __restrict__ nhot) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < S && cnt[s] > (uint32_t)HOT) hot[atomicAdd(nhot, 1u)] = s;
}

template <bool VEC4>
__global__ void __launch_bounds__(256) bag_bwd_reduce_kernel(int S, int h, const uint32_t* __restrict__ cnt, const int32_t* __restrict__ slots,
                                                             const float* __restrict__ dZ, float* __restrict__ dW0T, int32_t* __restrict__ hot,
                                                             uint32_t* __restrict__ nhot) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < S; s += warps) {
    const int n = (int)__ldg(cnt + s);
    if (n > HOT) continue;  // on the hot list (bag_bwd_hotlist_kernel): reduced by csr_bag_bwd_hot_kernel, possibly at the same time
    int rid = 0x7fffffff;
    if (lane < n) rid = __ldg(slots + (size_t)s * HOT + lane);
    const int key = rid & 0x7fffffff;
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (__shfl_sync(0xffffffffu, key, j) < key);
    // lane i now needs the entry whose rank is i
    int sorted = 0;
    for (int i = 0; i < n; ++i) {
      const unsigned m = __ballot_sync(0xffffffffu, lane < n && rank == i);
      const int v = __shfl_sync(0xffffffffu, rid, __ffs(m) - 1);
      if (lane == i) sorted = v;
    }
    if (VEC4) {
      const int h4 = h >> 2;
      for (int c0 = 0; c0 < h4; c0 += 32) {  // (the shuffles need every lane: the column guard is on the memory accesses only)
        const int c = min(c0 + lane, h4 - 1);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int i = 0;
        for (; i + 4 <= n; i += 4) {  // 4 rows in flight, added in order
          int r[4]; float4 v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) r[q] = __shfl_sync(0xffffffffu, sorted, i + q);
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = __ldg(reinterpret_cast<const float4*>(dZ + (size_t)(r[q] & 0x7fffffff) * h) + c);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float sg = r[q] < 0 ? -1.f : 1.f;
            acc.x += sg * v[q].x; acc.y += sg * v[q].y; acc.z += sg * v[q].z; acc.w += sg * v[q].w;
          }
        }
        for (; i < n; ++i) {
          const int r = __shfl_sync(0xffffffffu, sorted, i);
          const float4 v = __ldg(reinterpret_cast<const float4*>(dZ + (size_t)(r & 0x7fffffff) * h) + c);
          const float sg = r < 0 ? -1.f : 1.f;
          acc.x += sg * v.x; acc.y += sg * v.y; acc.z += sg * v.z; acc.w += sg * v.w;
        }
        if (c0 + lane < h4) reinterpret_cast<float4*>(dW0T + (size_t)s * h)[c] = acc;
      }
    } else {
      for (int c0 = 0; c0 < h; c0 += 32) {  // (the shuffles need every lane: the column guard is on the memory accesses only)
        const int c = c0 + lane;
        float acc = 0.f;
        for (int i = 0; i < n; ++i) {
          const int r = __shfl_sync(0xffffffffu, sorted, i);
          const float v = c < h ? __ldg(dZ + (size_t)(r & 0x7fffffff) * h + c) : 0.f;
          acc += r < 0 ? -v : v;
        }
        if (c < h) dW0T[(size_t)s * h + c] = acc;
      }
    }
  }
}

__global__ void __launch_bounds__(BWD_WARPS * 32) csr_bag_bwd_hot_kernel(int B, const int32_t* __restrict__ indptr,
                                                                          const int32_t* __restrict__ indices,
                                                                          const int32_t* __restrict__ ent_row, int row_base,
                                                                          const float* __restrict__ dZ, int h,
                                                                          const int32_t* __restrict__ hot,
                                                                          const uint32_t* __restrict__ nhot_p, float* __restrict__ dW0T,
                                                                          const uint32_t* __restrict__ ent_sign, float* __restrict__ hot_part) {
  // Work item = (hot skill, one of HOT_G fixed ranges of the batch's teams): a popular skill's sum is cut across CTAs by TEAM
  // range, each item adds its teams in entry order, and bag_bwd_hot_combine_kernel adds the HOT_G partials in range order --
  // the result depends on the data only.  (Hot skills past HOT_MAX, if any, are summed whole by one CTA each.)
  extern __shared__ float sm[];
  float* total = sm;                        // [h]
  float* partial = sm + h;                  // [BWD_WARPS][h]
  int* hits = (int*)(sm + (size_t)(1 + BWD_WARPS) * h);  // [HCAP]
  __shared__ uint32_t warp_cnt[BWD_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int p_first = indptr[0];
  const int nhot = (int)*nhot_p;
  const int nsplit = min(nhot, HOT_MAX);
  const int nitems = nsplit * HOT_G + (nhot - nsplit);
  constexpr int PER = 4;                    // consecutive entries per thread per pass (4 loads in flight)
  constexpr int PASS = BWD_WARPS * 32 * PER;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const bool whole = item >= nsplit * HOT_G;
    const int hi = whole ? nsplit + (item - nsplit * HOT_G) : item / HOT_G;
    const int gr = whole ? 0 : item % HOT_G;
    const int r0 = whole ? 0 : (int)((long long)B * gr / HOT_G), r1 = whole ? B : (int)((long long)B * (gr + 1) / HOT_G);
    const int p_beg = indptr[r0], p_end = indptr[r1];
    const int s = hot[hi];
    for (int c = tid; c < h; c += blockDim.x) total[c] = 0.f;
    int filled = 0;
    __syncthreads();
    for (int base = p_beg; base < p_end; base += PASS) {
      // ordered compaction of this skill's entries: thread tid owns entries base + PER*tid .. +PER-1
      const int p0 = base + tid * PER;
      int sk[PER];
#pragma unroll
      for (int q = 0; q < PER; ++q) sk[q] = p0 + q < p_end ? __ldg(indices + p0 + q) : -1;
      int mine = 0;
#pragma unroll
      for (int q = 0; q < PER; ++q) mine += (sk[q] == s);
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) warp_cnt[w] = (uint32_t)incl;
      __syncthreads();
      int off = 0, tot = 0;
#pragma unroll
      for (int q = 0; q < BWD_WARPS; ++q) { if (q < w) off += warp_cnt[q]; tot += warp_cnt[q]; }
      int slot = filled + off + incl - mine;
#pragma unroll
      for (int q = 0; q < PER; ++q)
        if (sk[q] == s) {
          int rr = __ldg(ent_row + p0 + q) - row_base;
          if (ent_sign) {
            const int e = p0 + q - p_first;
            if ((__ldg(ent_sign + (e >> 5)) >> (e & 31)) & 1u) rr |= (int)0x80000000u;
          }
          hits[slot++] = rr;
        }
      filled += tot;
      __syncthreads();
      const bool last = base + PASS >= p_end;
      if (filled > HCAP - PASS || last) {  // reduce this segment of the entry list: 8 warps x fixed slices, 8 rows in flight
        const int per = (filled + BWD_WARPS - 1) / BWD_WARPS;
        const int i0 = min(filled, w * per), i1 = min(filled, (w + 1) * per);
        for (int c = lane; c < h; c += 32) {
          float a = 0.f;
          int i = i0;
          for (; i + 8 <= i1; i += 8) {
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = __ldg(dZ + (size_t)(hits[i + q] & 0x7fffffff) * h + c);
#pragma unroll
            for (int q = 0; q < 8; ++q) a += hits[i + q] < 0 ? -v[q] : v[q];
          }
          for (; i < i1; ++i) { const float v1 = __ldg(dZ + (size_t)(hits[i] & 0x7fffffff) * h + c); a += hits[i] < 0 ? -v1 : v1; }
          partial[(size_t)w * h + c] = a;
        }
        __syncthreads();
        for (int c = tid; c < h; c += blockDim.x) {
          float t = total[c];
#pragma unroll
          for (int q = 0; q < BWD_WARPS; ++q) t += partial[(size_t)q * h + c];
          total[c] = t;
        }
        filled = 0;
        __syncthreads();
      }
    }
    float* dst = whole ? dW0T + (size_t)s * h : hot_part + ((size_t)hi * HOT_G + gr) * h;
    for (int c = tid; c < h; c += blockDim.x) dst[c] = total[c];
    __syncthreads();
  }
}

__global__ void bag_bwd_hot_combine_kernel(int h, const int32_t* __restrict__ hot, const uint32_t* __restrict__ nhot_p,
                                           const float* __restrict__ hot_part, float* __restrict__ dW0T) {
  const int nsplit = min((int)*nhot_p, HOT_MAX);
  for (int hi = blockIdx.x; hi < nsplit; hi += gridDim.x) {
    const int s = hot[hi];
    for (int c = threadIdx.x; c < h; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int gr = 0; gr < HOT_G; ++gr) t += hot_part[((size_t)hi * HOT_G + gr) * h + c];
      dW0T[(size_t)s * h + c] = t;
    }
  }
}
}  // namespace

extern "C" size_t ntf_csr_bag_bwd_workspace_bytes(int S, int h) {
  return align_up(((size_t)S * (2 + HOT) + 64) * sizeof(uint32_t), 256) + align_up((size_t)HOT_MAX * HOT_G * h * sizeof(float), 256);
}

struct BagBwdWs {
  uint32_t* cnt; uint32_t* nhot; int32_t* hot; int32_t* slots; float* hot_part;
};
static BagBwdWs bag_bwd_ws(void* workspace, int S) {
  BagBwdWs w;
  w.cnt = (uint32_t*)workspace;        // [S]
  w.nhot = w.cnt + S;                  // [1] (padded to 64)
  w.hot = (int32_t*)(w.cnt + S + 64);  // [S]
  w.slots = w.hot + S;                 // [S][HOT]
  w.hot_part = (float*)((char*)workspace + align_up(((size_t)S * (2 + HOT) + 64) * sizeof(uint32_t), 256));  // [HOT_MAX][HOT_G][h]
  return w;
}

// pass 1 depends on the batch's CSR only (not on dZ): ntf_fnn_step runs it on a side stream while the output layer computes
int ntf_csr_bag_bwd_fill_impl(ntf_ctx* ctx, cudaStream_t st, int B, const int32_t* indptr, const int32_t* indices, const int32_t* ent_row,
                              int row_base, int S, int h, void* workspace, size_t workspace_bytes, const uint32_t* ent_sign) {
  NTF_REQUIRE(ctx && indptr && indices && ent_row && workspace, NTF_ERR_BAD_ARG, "csr_bag_bwd: null pointer");
  NTF_REQUIRE(B > 0 && S > 0 && h > 0, NTF_ERR_BAD_ARG, "csr_bag_bwd: B=%d S=%d h=%d", B, S, h);
  NTF_REQUIRE(workspace_bytes >= ntf_csr_bag_bwd_workspace_bytes(S, h), NTF_ERR_WORKSPACE, "csr_bag_bwd: workspace too small");
  const BagBwdWs w = bag_bwd_ws(workspace, S);
  NTF_CUDA(cudaMemsetAsync(w.cnt, 0, (size_t)(S + 64) * sizeof(uint32_t), st));
  NTF_COUNT_LAUNCH; bag_bwd_fill_kernel<<<min(cdiv(B * 8, 256), ctx->sm_count * 8), 256, 0, st>>>(B, indptr, indices, ent_row, row_base, ent_sign, w.cnt, w.slots);
  NTF_COUNT_LAUNCH; bag_bwd_hotlist_kernel<<<cdiv(S, 256), 256, 0, st>>>(S, w.cnt, w.hot, w.nhot);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

int ntf_csr_bag_bwd_reduce_impl(ntf_ctx* ctx, cudaStream_t st, int B, const int32_t* indptr, const int32_t* indices, const int32_t* ent_row,
                                int row_base, const float* dZ, int S, int h, float* dW0T, void* workspace, size_t workspace_bytes,
                                const uint32_t* ent_sign, cudaStream_t st_hot) {
  // st_hot (nullable): the hot skills' kernels run there, next to the warp-per-skill reduction on `st` (disjoint rows of dW0T); joined
  // back into `st` before returning.  Both need dZ, which is complete on `st` when this is called.
  NTF_REQUIRE(ctx && indptr && indices && ent_row && dZ && dW0T && workspace, NTF_ERR_BAD_ARG, "csr_bag_bwd: null pointer");
  NTF_REQUIRE(h <= 2048, NTF_ERR_UNSUPPORTED, "csr_bag_bwd: first hidden width %d > 2048", h);
  NTF_REQUIRE(workspace_bytes >= ntf_csr_bag_bwd_workspace_bytes(S, h), NTF_ERR_WORKSPACE, "csr_bag_bwd: workspace too small");
  const BagBwdWs w = bag_bwd_ws(workspace, S);
  const bool vec = (h % 4 == 0) && (((uintptr_t)dZ | (uintptr_t)dW0T) % 16 == 0);
  const int blocks = min(cdiv(S, 8), ctx->sm_count * 8);
  if (st_hot != nullptr && st_hot != st) NTF_CUDA(cudaEventRecord(ctx->ev_hot_fork, st));
  NTF_COUNT_LAUNCH;
  if (vec) bag_bwd_reduce_kernel<true><<<blocks, 256, 0, st>>>(S, h, w.cnt, w.slots, dZ, dW0T, w.hot, w.nhot);
  else bag_bwd_reduce_kernel<false><<<blocks, 256, 0, st>>>(S, h, w.cnt, w.slots, dZ, dW0T, w.hot, w.nhot);
  const size_t smem_hot = (size_t)(1 + BWD_WARPS) * h * sizeof(float) + (size_t)HCAP * sizeof(int);
  NTF_CUDA(cudaFuncSetAttribute(csr_bag_bwd_hot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_hot));
  const bool fork = st_hot != nullptr && st_hot != st;
  cudaStream_t sh = fork ? st_hot : st;
  if (fork) NTF_CUDA(cudaStreamWaitEvent(sh, ctx->ev_hot_fork, 0));  // (recorded on st before the reduce kernel was enqueued, below)
  NTF_COUNT_LAUNCH; csr_bag_bwd_hot_kernel<<<ctx->sm_count * 2, BWD_WARPS * 32, smem_hot, sh>>>(B, indptr, indices, ent_row, row_base, dZ, h, w.hot, w.nhot, dW0T, ent_sign, w.hot_part);
  NTF_COUNT_LAUNCH; bag_bwd_hot_combine_kernel<<<ctx->sm_count, 128, 0, sh>>>(h, w.hot, w.nhot, w.hot_part, dW0T);
  NTF_LAUNCH_CHECK();
  if (fork) {
    NTF_CUDA(cudaEventRecord(ctx->ev_hot_join, sh));
    NTF_CUDA(cudaStreamWaitEvent(st, ctx->ev_hot_join, 0));
  }
  return NTF_OK;
}

static int csr_bag_bwd_impl(ntf_ctx* ctx, void* stream, int B, const int32_t* indptr, const int32_t* indices,
                            const int32_t* ent_row, int row_base, const float* dZ, int S, int h, float* dW0T,
                            void* workspace, size_t workspace_bytes, const uint32_t* ent_sign) {
  const int rc = ntf_csr_bag_bwd_fill_impl(ctx, as_stream(stream), B, indptr, indices, ent_row, row_base, S, h, workspace, workspace_bytes, ent_sign);
  if (rc) return rc;
  return ntf_csr_bag_bwd_reduce_impl(ctx, as_stream(stream), B, indptr, indices, ent_row, row_base, dZ, S, h, dW0T, workspace, workspace_bytes, ent_sign, nullptr);
}

extern "C" int ntf_csr_bag_bwd(ntf_ctx* ctx, void* stream, int B, const int32_t* indptr, const int32_t* indices,
                               const int32_t* ent_row, int row_base, const float* dZ, int S, int h, float* dW0T,
                               void* workspace, size_t workspace_bytes) {
  return csr_bag_bwd_impl(ctx, stream, B, indptr, indices, ent_row, row_base, dZ, S, h, dW0T, workspace, workspace_bytes, nullptr);
}

// Flipout: dW0T_delta[s,:] = sum_{entries (n,s)} sgn_p * dZs[n,:]   (same owner-computes reduction, sign per CSR entry)
extern "C" int ntf_csr_bag_bwd_signed(ntf_ctx* ctx, void* stream, int B, const int32_t* indptr, const int32_t* indices,
                                      const int32_t* ent_row, int row_base, const uint32_t* ent_sign, const float* dZs, int S, int h,
                                      float* dW0T_delta, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ent_sign != nullptr, NTF_ERR_BAD_ARG, "csr_bag_bwd_signed: ent_sign is NULL");
  return csr_bag_bwd_impl(ctx, stream, B, indptr, indices, ent_row, row_base, dZs, S, h, dW0T_delta, workspace, workspace_bytes, ent_sign);
}

// =========================================================================================================
// dZ = dY * lrelu'(Y) ; db = colsum(dZ).  Column sums are taken by one thread per column over a fixed row
// order inside a block of rows, then across row-blocks by a second fixed-order pass: deterministic.
// =========================================================================================================
namespace {
constexpr int ACT_ROWS = 16;  // rows per block (B=1000: 63 blocks; the partials are combined in block order)

__global__ void act_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ Y, int B, int h, int act,
                               float* __restrict__ dZ, float* __restrict__ part) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= h) return;
  const int r0 = blockIdx.y * ACT_ROWS, r1 = min(B, r0 + ACT_ROWS);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) {
    const size_t k = (size_t)r * h + c;
    float g = dY[k];
    if (act) g *= (Y[k] > 0.f ? 1.f : NTF_LRELU_SLOPE);
    if (dZ) dZ[k] = g;
    s += g;
  }
  part[(size_t)blockIdx.y * h + c] = s;
}

__global__ void colsum_parts_kernel(const float* __restrict__ part, int nparts, int h, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= h) return;
  float s = 0.f;
  for (int k = 0; k < nparts; ++k) s += part[(size_t)k * h + c];
  out[c] = s;
}
}  // namespace

extern "C" size_t ntf_act_bwd_workspace_bytes(int B, int h) { return align_up((size_t)cdiv(B, ACT_ROWS) * h * sizeof(float), 256); }

extern "C" int ntf_act_bwd(ntf_ctx* ctx, void* stream, const float* dY, const float* Y, int B, int h, int act, float* dZ,
                           float* db, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ctx && dY && db && workspace && (!act || Y), NTF_ERR_BAD_ARG, "act_bwd: null pointer");
  NTF_REQUIRE(B > 0 && h > 0, NTF_ERR_BAD_ARG, "act_bwd: B=%d h=%d", B, h);
  NTF_REQUIRE(workspace_bytes >= ntf_act_bwd_workspace_bytes(B, h), NTF_ERR_WORKSPACE, "act_bwd: workspace too small");
  const int nparts = cdiv(B, ACT_ROWS);
  dim3 grid(cdiv(h, 128), nparts);
  NTF_COUNT_LAUNCH; act_bwd_kernel<<<grid, 128, 0, as_stream(stream)>>>(dY, Y, B, h, act, dZ, (float*)workspace);
  NTF_COUNT_LAUNCH; colsum_parts_kernel<<<cdiv(h, 128), 128, 0, as_stream(stream)>>>((const float*)workspace, nparts, h, db);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
