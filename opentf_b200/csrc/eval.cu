// eval.cu -- ranking metrics of the top-K output on the device: the consumer of ntf_topk_select.
//
// Replaces the per-team python loop of src/evl/metric.py:12-33 (argpartition/argsort per row, two dicts per team, then
// pytrec_eval's C extension): one CTA per team ranks the team's K candidates exactly as trec_eval does -- score descending,
// ties by document id STRING descending ('d' + str(expert), metric.py:31) -- looks the ranked experts up in the team's member
// row (the qrels, metric.py:30) and accumulates P_k, recall_k, ndcg_cut_k, map_cut_k and success_k for every requested cut-off.
// Only [n, 5, nk] numbers leave the GPU; they are computed in fp64 like trec_eval's.
//
// Integer / index work plus a handful of fp64 flops per team: latency-bound, HBM traffic = 8*K + 4*n_members bytes per team.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int EV_THREADS = 256;
constexpr int EV_MAXK = 1024;  // candidates per team held in shared memory (the reference's default first-stage topK is 1000)
constexpr int EV_MAXCUT = 16;  // cut-offs per call

// order of the decimal strings str(a) vs str(b) as python / trec_eval compare them: left-align the digits; a proper prefix is smaller
__device__ __forceinline__ uint64_t doc_key(int32_t col) {
  uint32_t c = (uint32_t)col;
  int len = 1;
  for (uint32_t t = c; t >= 10u; t /= 10u) ++len;
  uint64_t k = c;
  for (int i = len; i < 10; ++i) k *= 10ull;  // int32 ids have at most 10 digits
  return k * 16ull + (uint64_t)len;
}

// monotone map float -> uint32 (larger float = larger key); NaNs sort last like trec_eval's comparison leaves them
__device__ __forceinline__ uint32_t val_key(float v) {
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

struct Cut { int k[EV_MAXCUT]; int n; };

// "a ranks before b": higher score first, ties by larger document-id string first
__device__ __forceinline__ bool before(uint32_t va, uint64_t da, uint32_t vb, uint64_t db) { return va != vb ? va > vb : da > db; }

__global__ void __launch_bounds__(EV_THREADS) eval_ranked_kernel(int n, int K, const int32_t* __restrict__ idx, const float* __restrict__ vals,
                                                                 const int32_t* __restrict__ m_indptr, const int32_t* __restrict__ m_indices,
                                                                 Cut cuts, double* __restrict__ out) {
  __shared__ uint32_t sv[EV_MAXK];
  __shared__ uint64_t sd[EV_MAXK];
  __shared__ int32_t sc[EV_MAXK];
  __shared__ uint8_t hit[EV_MAXK];
  const int team = blockIdx.x;
  if (team >= n) return;
  int P2 = 1;
  while (P2 < K) P2 <<= 1;
  for (int i = threadIdx.x; i < P2; i += EV_THREADS) {
    int32_t c = -1;
    float v = 0.f;
    if (i < K) { c = idx[(size_t)team * K + i]; v = vals[(size_t)team * K + i]; }
    const bool real = c >= 0;  // padding (fewer than K candidates) sorts behind every real candidate
    sc[i] = c;
    sv[i] = real ? val_key(v) : 0u;
    sd[i] = real ? doc_key(c) : 0ull;
    if (real && sv[i] == 0u) sv[i] = 1u;
  }
  __syncthreads();
  // bitonic sort into ranking order
  for (int size = 2; size <= P2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (P2 >> 1); t += EV_THREADS) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride), hi = lo + stride;
        const bool up = ((lo & size) == 0);  // this sub-sequence sorts "best first"
        const bool lo_first = before(sv[lo], sd[lo], sv[hi], sd[hi]);
        const bool tie = (sv[lo] == sv[hi]) && (sd[lo] == sd[hi]);
        if (!tie && (lo_first != up)) {
          const uint32_t tv = sv[lo]; sv[lo] = sv[hi]; sv[hi] = tv;
          const uint64_t td = sd[lo]; sd[lo] = sd[hi]; sd[hi] = td;
          const int32_t tc = sc[lo]; sc[lo] = sc[hi]; sc[hi] = tc;
        }
      }
      __syncthreads();
    }
  }
  const int mb = m_indptr[team], me = m_indptr[team + 1];
  const int kmax = cuts.k[cuts.n - 1] < K ? cuts.k[cuts.n - 1] : K;
  for (int r = threadIdx.x; r < kmax; r += EV_THREADS) {  // is the expert at rank r a member? (member rows are sorted)
    const int32_t c = sc[r];
    int a = mb, b = me;
    while (a < b) {
      const int mid = (a + b) >> 1;
      if (m_indices[mid] < c) a = mid + 1; else b = mid;
    }
    hit[r] = (c >= 0 && a < me && m_indices[a] == c) ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int R = me - mb;
    double* o = out + (size_t)team * 5 * cuts.n;
    int ci = 0, hits = 0;
    double dcg = 0.0, idcg = 0.0, ap = 0.0;
    for (int r = 0; ci < cuts.n; ++r) {  // r = rank (0-based); a cut-off beyond the candidates keeps counting misses
      while (ci < cuts.n && cuts.k[ci] == r) {
        const int k = cuts.k[ci];
        o[0 * cuts.n + ci] = (double)hits / (double)k;                    // P_k
        o[1 * cuts.n + ci] = R ? (double)hits / (double)R : 0.0;          // recall_k
        o[2 * cuts.n + ci] = (R && idcg > 0.0) ? dcg / idcg : 0.0;        // ndcg_cut_k
        o[3 * cuts.n + ci] = R ? ap / (double)R : 0.0;                    // map_cut_k
        o[4 * cuts.n + ci] = hits > 0 ? 1.0 : 0.0;                        // success_k
        ++ci;
      }
      if (ci >= cuts.n) break;
      const double disc = 1.0 / log2((double)r + 2.0);
      if (r < R) idcg += disc;
      if (r < kmax && hit[r]) {
        ++hits;
        dcg += disc;
        ap += (double)hits / (double)(r + 1);
      }
    }
  }
}

}  // namespace

// idx / vals [n,K]: the candidates of every team in ANY order (idx < 0 = padding); m_indptr (n+1 absolute offsets) / m_indices: the
// teams' true members, columns sorted; ks[nk]: cut-offs, ascending, HOST array; out [n][5][nk] fp64 (device):
// families in the order P, recall, ndcg_cut, map_cut, success.
extern "C" int ntf_eval_ranked(ntf_ctx* ctx, void* stream, int n, int K, const int32_t* idx, const float* vals, const int32_t* m_indptr,
                               const int32_t* m_indices, const int* ks, int nk, double* out) {
  NTF_REQUIRE(ctx && idx && vals && m_indptr && m_indices && ks && out, NTF_ERR_BAD_ARG, "eval_ranked: null pointer");
  NTF_REQUIRE(n >= 0 && K >= 1 && K <= EV_MAXK, NTF_ERR_UNSUPPORTED, "eval_ranked: K=%d (1..%d candidates per team)", K, EV_MAXK);
  NTF_REQUIRE(nk >= 1 && nk <= EV_MAXCUT, NTF_ERR_UNSUPPORTED, "eval_ranked: %d cut-offs (1..%d)", nk, EV_MAXCUT);
  Cut c;
  c.n = nk;
  for (int i = 0; i < nk; ++i) {
    NTF_REQUIRE(ks[i] >= 1 && (i == 0 || ks[i] > ks[i - 1]), NTF_ERR_BAD_ARG, "eval_ranked: cut-offs must be positive and ascending");
    c.k[i] = ks[i];
  }
  if (n == 0) return NTF_OK;
  NTF_COUNT_LAUNCH; eval_ranked_kernel<<<n, EV_THREADS, 0, as_stream(stream)>>>(n, K, idx, vals, m_indptr, m_indices, c, out);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
