// topk.cu -- test-time ranking (K5 selection stage): per team, the K largest scores in rank order.
// Replaces torch.topk over the host-resident [N,E] prediction matrix (pkgmgr.py:125-134) and the per-row
// argpartition/argsort of evl/metric.py:17-28.  Rank order: value descending, ties -> lower expert id first
// (the reference leaves tie order unspecified).
//
// ntf_topk_select: one CTA per team.  4-pass 8-bit radix select on the order-preserving integer image of the
// fp32 score finds the K-th value exactly, an index-ordered compaction takes everything above it plus the
// lowest-id ties, and a shared-memory bitonic sort puts the K survivors in rank order.  HBM/L2-bound:
// 5 reads of the row (4*E bytes each; L2-resident after the first).
#include "common.cuh"

namespace {
constexpr int TK_THREADS = 256;
constexpr int TK_KMAX = 2048;

__device__ __forceinline__ uint32_t ordered_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// descending bitonic sort of n (power of two) 64-bit keys in shared memory
__device__ void bitonic_desc(unsigned long long* a, int n) {
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = a[i], y = a[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (x < y) : (x > y)) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
}

__global__ void __launch_bounds__(TK_THREADS) topk_select_kernel(const float* __restrict__ P, int E, int K, int Kpad, float scale,
                                                                 float* __restrict__ vals, int32_t* __restrict__ idx) {
  extern __shared__ unsigned long long sel[];  // Kpad composite keys: (ordered score << 32) | ~expert
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_bucket, s_remaining, s_count, s_ties;
  __shared__ uint32_t warp_gt[TK_THREADS / 32], warp_eq[TK_THREADS / 32];
  const float* p = P + (size_t)blockIdx.x * E;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

  uint32_t prefix = 0, mask = 0;
  if (tid == 0) s_remaining = (uint32_t)K;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    hist[tid] = 0;  // TK_THREADS == 256
    __syncthreads();
    for (int j = tid; j < E; j += TK_THREADS) {
      const uint32_t k = ordered_key(__ldg(p + j) * scale);
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t c = 0, rem = s_remaining;
      int b = 255;
      for (; b > 0; --b) { if (c + hist[b] >= rem) break; c += hist[b]; }
      s_bucket = (uint32_t)b; s_remaining = rem - c;
    }
    __syncthreads();
    prefix |= s_bucket << shift; mask |= 255u << shift;
  }
  const uint32_t thr = prefix;          // ordered key of the K-th largest score
  const uint32_t need_ties = s_remaining;  // how many elements equal to thr belong to the top K
  if (tid == 0) { s_count = 0; s_ties = 0; }
  for (int i = tid; i < Kpad; i += TK_THREADS) sel[i] = 0ull;
  __syncthreads();
  // index-ordered compaction
  for (int base = 0; base < E; base += TK_THREADS) {
    const int j = base + tid;
    uint32_t k = 0; bool gt = false, eq = false;
    if (j < E) { k = ordered_key(__ldg(p + j) * scale); gt = k > thr; eq = k == thr; }
    const unsigned bgt = __ballot_sync(0xffffffffu, gt), beq = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) { warp_gt[w] = __popc(bgt); warp_eq[w] = __popc(beq); }
    __syncthreads();
    uint32_t off_gt = 0, off_eq = 0, tot_gt = 0, tot_eq = 0;
#pragma unroll
    for (int q = 0; q < TK_THREADS / 32; ++q) {
      if (q < w) { off_gt += warp_gt[q]; off_eq += warp_eq[q]; }
      tot_gt += warp_gt[q]; tot_eq += warp_eq[q];
    }
    const uint32_t lm = (1u << lane) - 1u;
    const uint32_t tie_rank = s_ties + off_eq + __popc(beq & lm);
    const bool take = gt || (eq && tie_rank < need_ties);
    // slot: survivors are appended in index order; count taken ties before me in this chunk
    const unsigned btake = __ballot_sync(0xffffffffu, take);
    __shared__ uint32_t warp_take[TK_THREADS / 32];
    if (lane == 0) warp_take[w] = __popc(btake);
    __syncthreads();
    uint32_t off_take = 0, tot_take = 0;
#pragma unroll
    for (int q = 0; q < TK_THREADS / 32; ++q) { if (q < w) off_take += warp_take[q]; tot_take += warp_take[q]; }
    if (take) {
      const uint32_t slot = s_count + off_take + __popc(btake & lm);
      if (slot < (uint32_t)Kpad) sel[slot] = ((unsigned long long)k << 32) | (uint32_t)(~(uint32_t)j);
    }
    __syncthreads();
    if (tid == 0) { s_count += tot_take; s_ties += tot_eq; }
    __syncthreads();
    (void)tot_gt; (void)off_gt;
  }
  bitonic_desc(sel, Kpad);
  for (int i = tid; i < K; i += TK_THREADS) {
    const unsigned long long c = sel[i];
    vals[(size_t)blockIdx.x * K + i] = key_to_float((uint32_t)(c >> 32));
    idx[(size_t)blockIdx.x * K + i] = (int32_t)(~(uint32_t)c);
  }
}

// Small K (<= 32: the k = 2..10 of the evaluation metrics, evl/metric.py:17-28): ONE pass over the row.  Each warp keeps the K best of
// its share of the row as 64-bit composites (ordered score << 32 | ~expert: larger = better score, then lower id), rank r in lane r.
// A lane compares its four new scores with the warp's K-th best -- after the first few hundred elements almost every score fails that
// single compare, so the warp-wide insert (one shuffle-shift) is rare -- and at the end warp 0 merges the 8 warp lists.
// HBM/L2-bound: 4*E bytes per team, read once.
__global__ void __launch_bounds__(TK_THREADS) topk_small_kernel(const float* __restrict__ P, int E, int K, float scale,
                                                                float* __restrict__ vals, int32_t* __restrict__ idx) {
  __shared__ unsigned long long wbest[TK_THREADS];  // [8 warps][32 ranks]
  const float* p = P + (size_t)blockIdx.x * E;
  const int tid = threadIdx.x, lane = tid & 31;
  unsigned long long mine = 0ull;   // rank `lane` of this warp's list (0 = empty; lanes >= K stay empty)
  unsigned long long kth = 0ull;    // the warp's K-th best = the bar a new score has to clear
  auto insert = [&](unsigned long long c) {  // warp-uniform c > kth
    const unsigned long long up = __shfl_up_sync(0xffffffffu, mine, 1);
    const unsigned long long prev = lane == 0 ? ~0ull : up;
    if (c > mine) mine = c > prev ? prev : c;  // ranks below the insertion point shift down by one, the point takes c
    if (lane >= K) mine = 0ull;
    kth = __shfl_sync(0xffffffffu, mine, K - 1);
  };
  auto offer4 = [&](float4 v, int j, bool ok) {
    unsigned long long c[4];
    c[0] = ((unsigned long long)ordered_key(v.x * scale) << 32) | (uint32_t)(~(uint32_t)j);
    c[1] = ((unsigned long long)ordered_key(v.y * scale) << 32) | (uint32_t)(~(uint32_t)(j + 1));
    c[2] = ((unsigned long long)ordered_key(v.z * scale) << 32) | (uint32_t)(~(uint32_t)(j + 2));
    c[3] = ((unsigned long long)ordered_key(v.w * scale) << 32) | (uint32_t)(~(uint32_t)(j + 3));
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      unsigned hit = __ballot_sync(0xffffffffu, ok && j + q < E && c[q] > kth);
      while (hit) {  // rare
        const int L = __ffs(hit) - 1;
        hit &= hit - 1;
        const unsigned long long cand = __shfl_sync(0xffffffffu, c[q], L);
        if (cand > kth) insert(cand);  // (an earlier insert of this round may have raised the bar)
      }
    }
  };
  const bool vec = (E & 3) == 0 && ((uintptr_t)p & 15) == 0;
  const int n4 = (E + 3) >> 2;  // groups of 4 consecutive scores; thread tid takes groups tid, tid + 256, ...
  for (int g0 = 0; g0 < n4; g0 += TK_THREADS * 4) {  // 4 loads in flight per thread
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int g = g0 + u * TK_THREADS + tid;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < n4) {
        if (vec) v[u] = __ldg(reinterpret_cast<const float4*>(p) + g);
        else {
          const int j = 4 * g;
          v[u].x = __ldg(p + j);
          if (j + 1 < E) v[u].y = __ldg(p + j + 1);
          if (j + 2 < E) v[u].z = __ldg(p + j + 2);
          if (j + 3 < E) v[u].w = __ldg(p + j + 3);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int g = g0 + u * TK_THREADS + tid;
      offer4(v[u], 4 * g, g < n4);
    }
  }
  wbest[tid] = mine;
  __syncthreads();
  if (tid < 32) {  // merge the 8 sorted warp lists: K rounds of "largest head wins and is popped"; lane l < 8 plays warp l's list
    int head = 0;
    unsigned long long out = 0ull;
    for (int r = 0; r < K; ++r) {
      unsigned long long h = (lane < TK_THREADS / 32 && head < K) ? wbest[lane * 32 + head] : 0ull;
      unsigned long long m = h;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
      }
      if (lane == r) out = m;
      if (h == m && m != 0ull) ++head;  // composites are unique (they carry the expert id): exactly one list pops
    }
    if (lane < K) {
      vals[(size_t)blockIdx.x * K + lane] = key_to_float((uint32_t)(out >> 32));
      idx[(size_t)blockIdx.x * K + lane] = (int32_t)(~(uint32_t)out);
    }
  }
}

// merge G candidate lists of length K per team: sort G*K composites, keep K
__global__ void __launch_bounds__(TK_THREADS) topk_merge_kernel(const float* __restrict__ vals_in, const int32_t* __restrict__ idx_in,
                                                                int G, int B, int K, int npad, float* __restrict__ vals,
                                                                int32_t* __restrict__ idx) {
  extern __shared__ unsigned long long sel[];
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < npad; i += TK_THREADS) {
    unsigned long long c = 0ull;
    if (i < G * K) {
      const int g = i / K, q = i % K;
      const size_t src = ((size_t)g * B + n) * K + q;
      const int32_t j = idx_in[src];
      if (j >= 0) c = ((unsigned long long)ordered_key(vals_in[src]) << 32) | (uint32_t)(~(uint32_t)j);
    }
    sel[i] = c;
  }
  __syncthreads();
  bitonic_desc(sel, npad);
  for (int i = threadIdx.x; i < K; i += TK_THREADS) {
    const unsigned long long c = sel[i];
    vals[(size_t)n * K + i] = key_to_float((uint32_t)(c >> 32));
    idx[(size_t)n * K + i] = c ? (int32_t)(~(uint32_t)c) : -1;
  }
}

// out[n] (+)= -sum_j q log(q + 1e-15), q = scale*P[n,j]
__global__ void __launch_bounds__(256) row_entropy_kernel(const float* __restrict__ P, int E, float scale, int accumulate,
                                                          float* __restrict__ out) {
  const float* p = P + (size_t)blockIdx.x * E;
  float s = 0.f;
  for (int j = threadIdx.x; j < E; j += 256) { const float q = p[j] * scale; s -= q * logf(q + 1e-15f); }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int q = 0; q < 8; ++q) t += red[q];
    out[blockIdx.x] = accumulate ? out[blockIdx.x] + t : t;
  }
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
}  // namespace

extern "C" int ntf_topk_select(ntf_ctx* ctx, void* stream, const float* P, int B, int E, int K, float scale, float* vals,
                               int32_t* idx) {
  NTF_REQUIRE(ctx && P && vals && idx, NTF_ERR_BAD_ARG, "topk_select: null pointer");
  NTF_REQUIRE(B > 0 && E > 0 && K > 0 && K <= E, NTF_ERR_BAD_ARG, "topk_select: B=%d E=%d K=%d", B, E, K);
  NTF_REQUIRE(K <= TK_KMAX, NTF_ERR_UNSUPPORTED, "topk_select: K=%d > %d", K, TK_KMAX);
  NTF_REQUIRE(scale > 0.f, NTF_ERR_BAD_ARG, "topk_select: scale must be positive");
  if (K <= 32) {  // single-pass kernel (a zero composite = "no candidate" cannot collide: a real score has a non-zero ordered key)
    NTF_COUNT_LAUNCH; topk_small_kernel<<<B, TK_THREADS, 0, as_stream(stream)>>>(P, E, K, scale, vals, idx);
    NTF_LAUNCH_CHECK();
    return NTF_OK;
  }
  const int Kpad = next_pow2(K);
  NTF_COUNT_LAUNCH; topk_select_kernel<<<B, TK_THREADS, (size_t)Kpad * 8, as_stream(stream)>>>(P, E, K, Kpad, scale, vals, idx);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

extern "C" int ntf_topk_merge(ntf_ctx* ctx, void* stream, const float* vals_in, const int32_t* idx_in, int G, int B, int K,
                              float* vals, int32_t* idx) {
  NTF_REQUIRE(ctx && vals_in && idx_in && vals && idx, NTF_ERR_BAD_ARG, "topk_merge: null pointer");
  NTF_REQUIRE(G > 0 && B > 0 && K > 0, NTF_ERR_BAD_ARG, "topk_merge: G=%d B=%d K=%d", G, B, K);
  const int npad = next_pow2(G * K);
  NTF_REQUIRE(npad <= 16384, NTF_ERR_UNSUPPORTED, "topk_merge: G*K=%d > 16384", G * K);
  const size_t smem = (size_t)npad * 8;
  NTF_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NTF_COUNT_LAUNCH; topk_merge_kernel<<<B, TK_THREADS, smem, as_stream(stream)>>>(vals_in, idx_in, G, B, K, npad, vals, idx);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

extern "C" int ntf_row_entropy(ntf_ctx* ctx, void* stream, const float* P, int B, int E, float scale, int accumulate, float* out) {
  NTF_REQUIRE(ctx && P && out, NTF_ERR_BAD_ARG, "row_entropy: null pointer");
  NTF_REQUIRE(B > 0 && E > 0, NTF_ERR_BAD_ARG, "row_entropy: B=%d E=%d", B, E);
  NTF_COUNT_LAUNCH; row_entropy_kernel<<<B, 256, 0, as_stream(stream)>>>(P, E, scale, accumulate, out);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
