// sampler.cu -- negative sampling (K4) and the bit-packed `condition` plane of fnn.py:33-43.
//
// The reference draws B x E random keys per step and keeps the top-ns per row (uniform: fnn.py:48-56; unigram /
// unigram_b: torch.multinomial without replacement == top-ns of p_j/Exp(1), fnn.py:58-76).  Both are "ns draws
// without replacement from the row's negatives, with probability proportional to p_j".  Here that distribution
// is sampled directly by successive draws, rejecting the row's own members and experts already drawn:
//   unigram   : inversion of the integer CDF of the all-team expert counts (built once per fold), O(ns log E) per team;
//   unigram_b : p_j = (teams of the batch having expert j) / B, so a draw is simply a uniformly chosen ENTRY of the batch's
//               member CSR (expert j owns count_j of the nnz entries) -- no histogram, no prefix sum, O(ns) per team;
//   uniform   : a uniform expert id.
// All integer arithmetic, so that oracle/sampler_oracle.py reproduces every index bit for bit.
#include "common.cuh"

int ntf_scan_u32_impl(cudaStream_t st, const uint32_t* in, size_t n, uint32_t* out, int inclusive, uint32_t* grand_total,
                      void* ws, size_t ws_bytes);
size_t ntf_scan_workspace_bytes(size_t n);

namespace {
constexpr uint32_t PURPOSE_NEG = 0x6e656730u;  // key1 ^= this for the negative sampler stream

__global__ void count_members_kernel(int B, const int32_t* __restrict__ m_indptr, const int32_t* __restrict__ m_indices,
                                     uint32_t* __restrict__ counts) {
  const int p0 = m_indptr[0], p1 = m_indptr[B];
  for (int p = p0 + blockIdx.x * blockDim.x + threadIdx.x; p < p1; p += gridDim.x * blockDim.x)
    atomicAdd(counts + m_indices[p], 1u);  // integer adds: the result does not depend on the order
}

__device__ __forceinline__ uint64_t draw64(uint64_t seed, uint32_t row, uint32_t t, uint64_t step) {
  const Philox4 r = philox4x32_10(row, t, (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)seed,
                                  (uint32_t)(seed >> 32) ^ PURPOSE_NEG);
  return ((uint64_t)r.v[1] << 32) | r.v[0];
}

// smallest j with cdf[j] > x
__device__ __forceinline__ int upper_bound_u32(const uint32_t* __restrict__ cdf, int E, uint32_t x) {
  int lo = 0, hi = E;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(cdf + mid) > x) hi = mid; else lo = mid + 1;
  }
  return lo;
}

constexpr int NS_MAX = 64;
constexpr int NS_WARPS = 4;

// One warp per team.  The reference semantics (restated sequentially in oracle/sampler_oracle.py) are: walk the draw sequence
// t = 0,1,2,..., accept draw j_t unless it is a member of the team or already accepted, stop at ns accepted.  "Accepted" = the
// first ns DISTINCT non-member values of the sequence, so 32 draws can be evaluated at once: lane l takes draw t_base+l
// (its own Philox counter and its own binary search), duplicates are resolved with match_any (earlier lane wins) and against
// the accepted list, and the survivors are appended in lane order.  max_tries is a multiple of 32, so the hand-over to the
// uniform top-up happens at the same draw index as in the sequential statement.
__global__ void __launch_bounds__(NS_WARPS * 32) neg_sample_kernel(int nsd, uint64_t seed, uint64_t step, int row0, int B,
                                                                   const int32_t* __restrict__ m_indptr, const int32_t* __restrict__ m_indices,
                                                                   int E, int ns, const uint32_t* __restrict__ cdf, int32_t* __restrict__ neg,
                                                                   const ntf_dyn* __restrict__ dyn, const int32_t* __restrict__ pool_indptr,
                                                                   int pool_rows) {
  __shared__ int sacc[NS_WARPS][NS_MAX];
  if (dyn) step = dyn->step;  // a replayed graph: this step's RNG counter comes from the block ntf_dyn_update wrote
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int* acc = sacc[w];
  const uint32_t lt = (1u << lane) - 1u;
  for (int n = blockIdx.x * NS_WARPS + w; n < B; n += gridDim.x * NS_WARPS) {
    const int pb = m_indptr[n], pe = m_indptr[n + 1], npos = pe - pb;
    auto is_pos = [&](int j) { for (int p = pb; p < pe; ++p) if (__ldg(m_indices + p) == j) return true; return false; };
    const int max_tries = 32 * ns + 64;
    int got = 0;
    uint32_t t = 0;
    bool weighted = (nsd == NTF_NS_UNIGRAM || nsd == NTF_NS_UNIGRAM_B);
    const bool pool = nsd == NTF_NS_UNIGRAM_B;  // draws are entries of the (global) batch's member CSR
    bool all_experts = false;  // fnn.py:67-69 fallback: uniform over ALL experts, members included
    uint32_t T = 0;
    int q0 = 0;
    if (pool) {
      q0 = pool_indptr[0];
      const int q1 = pool_indptr[pool_rows];
      T = (uint32_t)(q1 - q0);
      // all the mass on this team's own members <=> every entry of the pool is one of them (normally settled by the first 32 entries)
      bool other = false;
      for (int base = q0; base < q1 && !other; base += 32) {
        const int p = base + lane;
        other = __any_sync(0xffffffffu, p < q1 && !is_pos(__ldg(m_indices + p)));
      }
      if (!other) { weighted = false; all_experts = true; }
    } else if (weighted) {
      T = __ldg(cdf + E - 1);
      uint32_t pos_mass = 0;
      for (int p = pb; p < pe; ++p) { const int j = __ldg(m_indices + p); pos_mass += __ldg(cdf + j) - (j ? __ldg(cdf + j - 1) : 0u); }
      if (T == pos_mass) { weighted = false; all_experts = true; }
    }
    auto round32 = [&](bool use_w, bool members_ok, int want) {
      const uint64_t r = draw64(seed, (uint32_t)(row0 + n), t + lane, step);
      int j;
      if (!use_w) j = (int)__umul64hi(r, (uint64_t)E);
      else if (pool) j = __ldg(m_indices + q0 + (int)__umul64hi(r, (uint64_t)T));
      else j = upper_bound_u32(cdf, E, (uint32_t)__umul64hi(r, (uint64_t)T));
      bool keep = members_ok || !is_pos(j);
      for (int q = 0; q < got; ++q) keep = keep && (acc[q] != j);
      const uint32_t same = __match_any_sync(0xffffffffu, j);
      keep = keep && ((same & lt) == 0u);
      const uint32_t kb = __ballot_sync(0xffffffffu, keep);
      const int slot = got + __popc(kb & lt);
      __syncwarp();
      if (keep && slot < want) acc[slot] = j;
      __syncwarp();
      got = min(want, got + __popc(kb));
      t += 32;
    };
    if (weighted)
      for (int tries = 0; got < ns && tries < max_tries; tries += 32) round32(true, false, ns);
    // uniform mode, the all-expert fallback, and the top-up when fewer than ns weighted candidates exist
    const int avail = all_experts ? E : E - npos;
    const int want = min(ns, avail);
    for (int tries = 0; got < want && tries < max_tries; tries += 32) round32(false, all_experts, want);
    if (got < want && lane == 0) {  // pathological rows (E - npos barely >= ns): deterministic sweep
      for (int j = 0; got < want && j < E; ++j) {
        bool ok = all_experts || !is_pos(j);
        for (int q = 0; q < got; ++q) ok = ok && (acc[q] != j);
        if (ok) acc[got++] = j;
      }
    }
    got = __shfl_sync(0xffffffffu, got, 0);
    __syncwarp();
    for (int q = lane; q < ns; q += 32) neg[(size_t)n * ns + q] = q < got ? acc[q] : -1;
    __syncwarp();
  }
}

__global__ void special_bits_kernel(int op, int B, const int32_t* __restrict__ m_indptr, const int32_t* __restrict__ m_indices,
                                    const int32_t* __restrict__ neg, int ns, int E, int e_lo, uint32_t* __restrict__ special, int pitch) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= B) return;
  uint32_t* row = special + (size_t)n * pitch;
  // one thread owns the whole row of the plane, so plain read-modify-write is race free
  for (int p = m_indptr[n]; p < m_indptr[n + 1]; ++p) {
    const int j = m_indices[p] - e_lo;  // expert-sharded output layer: this rank's columns are [e_lo, e_lo + E)
    if (j < 0 || j >= E) continue;
    if (op) row[j >> 5] |= 1u << (j & 31); else row[j >> 5] = 0u;
  }
  if (neg)
    for (int q = 0; q < ns; ++q) {
      if (neg[(size_t)n * ns + q] < 0) continue;
      const int j = neg[(size_t)n * ns + q] - e_lo;
      if (j < 0 || j >= E) continue;
      if (op) row[j >> 5] |= 1u << (j & 31); else row[j >> 5] = 0u;
    }
}

// tile-transposed planes for the tensor-core output kernel: word ((n/128)*Epad + j)*4 + (n%128)/32, bit n%32
__global__ void special_tiles_kernel(int op, int B, const int32_t* __restrict__ m_indptr, const int32_t* __restrict__ m_indices,
                                     const int32_t* __restrict__ neg, int ns, int E, int e_lo, int Epad, uint32_t* __restrict__ special_t,
                                     uint32_t* __restrict__ member_t) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= B) return;
  const size_t slab = (size_t)(n >> 7) * Epad;
  const int wq = (n & 127) >> 5;
  const uint32_t bit = 1u << (n & 31);
  for (int p = m_indptr[n]; p < m_indptr[n + 1]; ++p) {
    const int jm = m_indices[p] - e_lo;
    if (jm < 0 || jm >= E) continue;
    const size_t w = (slab + jm) * 4 + wq;
    if (op) { atomicOr(special_t + w, bit); atomicOr(member_t + w, bit); } else { special_t[w] = 0u; member_t[w] = 0u; }  // clearing: every writer stores 0
  }
  if (neg)
    for (int q = 0; q < ns; ++q) {
      if (neg[(size_t)n * ns + q] < 0) continue;
      const int j = neg[(size_t)n * ns + q] - e_lo;
      if (j < 0 || j >= E) continue;
      const size_t w = (slab + j) * 4 + wq;
      if (op) atomicOr(special_t + w, bit); else special_t[w] = 0u;
    }
}
}  // namespace

extern "C" size_t ntf_special_tiles_bytes(int B, int E) { return (size_t)cdiv(B, 128) * (size_t)(cdiv(E, 128) * 128) * 16; }

extern "C" int ntf_special_tiles(ntf_ctx* ctx, void* stream, int op, int B, const int32_t* m_indptr, const int32_t* m_indices,
                                 const int32_t* neg, int ns, int E, int e_lo, uint32_t* special_t, uint32_t* member_t) {
  NTF_REQUIRE(ctx && m_indptr && m_indices && special_t && member_t, NTF_ERR_BAD_ARG, "special_tiles: null pointer");
  NTF_REQUIRE(B > 0 && E > 0, NTF_ERR_BAD_ARG, "special_tiles: B=%d E=%d", B, E);
  NTF_REQUIRE((((uintptr_t)special_t | (uintptr_t)member_t) & 15) == 0, NTF_ERR_BAD_ARG, "special_tiles: planes must be 16-byte aligned");
  NTF_COUNT_LAUNCH; special_tiles_kernel<<<cdiv(B, 128), 128, 0, as_stream(stream)>>>(op, B, m_indptr, m_indices, neg, ns, E, e_lo, cdiv(E, 128) * 128, special_t, member_t);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

extern "C" size_t ntf_expert_cdf_workspace_bytes(int E) { return ntf_scan_workspace_bytes((size_t)E); }

extern "C" int ntf_expert_cdf(ntf_ctx* ctx, void* stream, int B, const int32_t* m_indptr, const int32_t* m_indices, int E,
                              uint32_t* counts, uint32_t* cdf, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ctx && m_indptr && m_indices && counts && cdf, NTF_ERR_BAD_ARG, "expert_cdf: null pointer");
  NTF_REQUIRE(B > 0 && E > 0, NTF_ERR_BAD_ARG, "expert_cdf: B=%d E=%d", B, E);
  cudaStream_t st = as_stream(stream);
  NTF_CUDA(cudaMemsetAsync(counts, 0, (size_t)E * sizeof(uint32_t), st));
  NTF_COUNT_LAUNCH; count_members_kernel<<<min(cdiv(B * 4, 256), ctx->sm_count * 8), 256, 0, st>>>(B, m_indptr, m_indices, counts);
  NTF_LAUNCH_CHECK();
  return ntf_scan_u32_impl(st, counts, (size_t)E, cdf, 1, nullptr, workspace, workspace_bytes);
}

int ntf_neg_sample_impl(ntf_ctx* ctx, void* stream, int nsd, uint64_t seed, uint64_t step, int row0, int B, const int32_t* m_indptr,
                        const int32_t* m_indices, int E, int ns, const uint32_t* cdf, const int32_t* pool_indptr, int pool_rows,
                        int32_t* neg, const ntf_dyn* dyn) {
  NTF_REQUIRE(ctx && m_indptr && m_indices && neg, NTF_ERR_BAD_ARG, "neg_sample: null pointer");
  NTF_REQUIRE(nsd >= NTF_NS_UNIFORM && nsd <= NTF_NS_UNIGRAM_B, NTF_ERR_BAD_ARG, "neg_sample: nsd=%d", nsd);
  NTF_REQUIRE(nsd != NTF_NS_UNIGRAM || cdf, NTF_ERR_BAD_ARG, "neg_sample: unigram needs the cdf of the all-team expert counts");
  if (!pool_indptr || pool_rows <= 0) { pool_indptr = m_indptr; pool_rows = B; }  // unigram_b: the batch itself
  NTF_REQUIRE(B > 0 && E > 0 && ns > 0 && ns <= NS_MAX, NTF_ERR_UNSUPPORTED, "neg_sample: B=%d E=%d ns=%d (ns<=%d)", B, E, ns, NS_MAX);
  NTF_COUNT_LAUNCH; neg_sample_kernel<<<min(cdiv(B, NS_WARPS), ctx->sm_count * 16), NS_WARPS * 32, 0, as_stream(stream)>>>(nsd, seed, step, row0, B, m_indptr, m_indices, E, ns, cdf, neg, dyn, pool_indptr, pool_rows);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

extern "C" int ntf_neg_sample(ntf_ctx* ctx, void* stream, int nsd, uint64_t seed, uint64_t step, int row0, int B,
                              const int32_t* m_indptr, const int32_t* m_indices, int E, int ns, const uint32_t* cdf,
                              const int32_t* pool_indptr, int pool_rows, int32_t* neg) {
  return ntf_neg_sample_impl(ctx, stream, nsd, seed, step, row0, B, m_indptr, m_indices, E, ns, cdf, pool_indptr, pool_rows, neg, ctx ? ctx->dyn_override : nullptr);
}

extern "C" int ntf_special_bits(ntf_ctx* ctx, void* stream, int op, int B, const int32_t* m_indptr, const int32_t* m_indices,
                                const int32_t* neg, int ns, int E, int e_lo, uint32_t* special, int pitch_words) {
  NTF_REQUIRE(ctx && m_indptr && m_indices && special, NTF_ERR_BAD_ARG, "special_bits: null pointer");
  NTF_REQUIRE(B > 0 && E > 0 && pitch_words * 32 >= E, NTF_ERR_BAD_ARG, "special_bits: B=%d E=%d pitch=%d", B, E, pitch_words);
  NTF_COUNT_LAUNCH; special_bits_kernel<<<cdiv(B, 128), 128, 0, as_stream(stream)>>>(op, B, m_indptr, m_indices, neg, ns, E, e_lo, special, pitch_words);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
