// infer_topk.cu -- K5: test-time top-K expert ranking fused with the output layer (fnn.py:204-218, pkgmgr.py:125-134).
//
// The reference computes sigmoid(model(X)) as a dense [N,E] matrix, moves it to the host and runs torch.topk over it.  Here the
// [B,E] score matrix never exists: the logits live in TMEM for the time a CTA looks at them, and only O(K) numbers per team
// leave the SM.  Selection is done on the logit z = a.w + b (sigmoid(lrelu(.)) is monotone, so the K best logits are the K best
// probabilities; the probability is evaluated for the K winners only), in TWO passes over the same tcgen05 product:
//
//   pass 1   Z[128 teams x 128 experts] = A16 . W16^T per (batch tile, expert tile); every epilogue thread (team, block of 32
//            consecutive experts) keeps only the block's MAXIMUM -> bm[B, E/32]   (4*E/32 bytes per team instead of 4*E)
//   select   per team the K-th largest block maximum (blockmax_threshold_kernel: bisection on the ordered keys): value Tz, block bK.
//            K blocks have a maximum >= Tz, so Tz is a lower bound of the K-th best logit, and everything that can be in the
//            top K has  z > Tz  or  (z == Tz and its block <= bK)  -- at most 32*K elements, all inside those K blocks.
//   pass 2   the same product again; elements passing that test are appended to the team's candidate list (global atomics on a
//            per-team counter: a few dozen per team), everything else is dropped in registers.
//   final    per team: sort the candidates by (z descending, expert id ascending), emit the first K as (probability, expert id).
//
// Recomputing the product is cheaper than writing and re-reading [B,E] (FlashAttention's trade): per team 2*2*h*E flops on the
// tensor pipe against 8*E bytes of HBM traffic.  Exact: the result equals the K best of the full score row under the library's
// rank order (value descending, ties -> lower expert id), as ntf_topk_select gives on ntf_infer_scores' output, whenever the
// logits are distinct where the probabilities are (equal probabilities of distinct logits -- sigmoid saturation -- are ranked
// by logit: a permutation among exact score ties).
//
// Layout: CTA = (a PAIR of batch tiles of 128 teams, chunk of consecutive expert tiles); grid = pairs x chunks sized to one wave.
//   warp 16     TMA: the two fp16 activation tiles once, then the fp16 W tiles of the chunk through a 3-stage ring
//   warp 17     MMA issuer: per W tile two products (M = 128 teams, N = 128 experts, K = 128 hidden: 8 x kind::f16), 4 accumulator stages in TMEM
//   warps 0-15  epilogue: thread = (team = TMEM lane, 32-expert column block): tcgen05.ld, + bias, max / threshold test
// W16 is an fp16 image of the layer's weight kept by the caller (ntf_to_half; rewritten when the parameters change).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace {

constexpr int TT = 128;   // teams per batch tile (UMMA M)
constexpr int TX = 128;   // experts per expert tile (UMMA N)
constexpr int HK = 128;   // hidden width
constexpr uint32_t CHUNK = 128 * 128;        // one swizzle chunk: 128 rows x 128 bytes (64 halfs)
constexpr uint32_t TILE_BYTES = 2 * CHUNK;   // an fp16 [128 x 128] operand tile: 2 chunks (hidden 0-63 | 64-127)
constexpr int W_STAGES = 3;
constexpr int Z_STAGES = 4;
constexpr int NH = 2;     // batch tiles per CTA: every W tile that comes in is multiplied with both (halves the W traffic out of L2, which bound the
                          // one-tile version: 8 batch tiles x 10 MB per pass at ~4.4 TB/s = 18.6 us measured)
constexpr uint32_t OFF_A = 0;
constexpr uint32_t OFF_W = OFF_A + NH * TILE_BYTES;
constexpr uint32_t OFF_BAR = OFF_W + W_STAGES * TILE_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256;
enum { BAR_A_FULL = 0, BAR_W_FULL = 1, BAR_W_EMPTY = 1 + W_STAGES, BAR_Z_FULL = 1 + 2 * W_STAGES, BAR_Z_EMPTY = 1 + 2 * W_STAGES + Z_STAGES,
       NUM_BARS = 1 + 2 * W_STAGES + 2 * Z_STAGES };
static_assert(NUM_BARS * 8 + 8 <= 256, "barrier block");
constexpr int EPI_WARPS = 16, WARP_TMA = 16, WARP_MMA = 17, NT = 18 * 32;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr uint32_t IDESC = instr_desc(0, 0, 0, TT, TX);  // Z = A16(K-major) . W16(K-major)^T, fp16 operands, fp32 accumulate
constexpr int BLK = 32;  // experts per block (= the 32 TMEM columns one epilogue thread reads)

__device__ __forceinline__ uint32_t ordered_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

struct ItArgs {
  const float* bias;     // [E]
  int B, E, K;
  int nbt, nct, nchunk;  // batch tiles, expert tiles, chunks of expert tiles
  int nblk;              // row pitch of bm = 4 * nct (blocks of 32 experts, the last tile padded)
  float* bm;             // pass 1 out: [B, nblk] block maxima (-inf for blocks past E)
  float* thr_val;        // pass 2 in: [B] the K-th largest block maximum of the team (blockmax_threshold_kernel), and
  int32_t* thr_blk;      //            [B] the block it belongs to under the rank order (value descending, block ascending)
  unsigned long long* cand;  // pass 2 out: [B, cap] composites (ordered logit << 32 | ~expert)
  int* cnt;                  // [B] candidates appended so far (zeroed by the caller)
  int cap;
  long long* timing;         // debug (NTF_IT_TIMING): per CTA 256 slots: [0] start clock, [1] end clock; from 8, 4 per product q < 60: W tile landed (MMA warp),
                             // product issued, logits ready (epilogue thread 0), epilogue done
};

template <int PASS>
__global__ void __launch_bounds__(NT, 1) infer_topk_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, ItArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if ((sbase & 1023u) != 0u) __trap();
  const uint32_t bars = sbase + OFF_BAR;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + OFF_BAR + NUM_BARS * 8);
  auto bar = [&](int i) { return bars + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // consecutive CTAs take different batch-tile pairs of the SAME chunk: they stream the same W tiles at the same time (L2 hits)
  const int nbp = (g.nbt + NH - 1) / NH;  // batch-tile pairs
  const int bp = blockIdx.x % nbp, chunk = blockIdx.x / nbp;
  const int nh = min(NH, g.nbt - bp * NH);  // batch tiles this CTA really has
  const int t0 = (int)((long long)chunk * g.nct / g.nchunk), t1 = (int)((long long)(chunk + 1) * g.nct / g.nchunk);
  const int ntiles = t1 - t0;
  if (g.timing && threadIdx.x == 0) g.timing[256 * blockIdx.x] = clock64();

  if (threadIdx.x == 0) {
    mbar_init(bar(BAR_A_FULL), 1);
    for (int s = 0; s < W_STAGES; ++s) { mbar_init(bar(BAR_W_FULL + s), 1); mbar_init(bar(BAR_W_EMPTY + s), 1); }
    for (int s = 0; s < Z_STAGES; ++s) { mbar_init(bar(BAR_Z_FULL + s), 1); mbar_init(bar(BAR_Z_EMPTY + s), EPI_THREADS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + NUM_BARS * 8), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == WARP_TMA) {
    if (lane == 0) {
      mbar_expect_tx(bar(BAR_A_FULL), nh * TILE_BYTES);
      for (int hf = 0; hf < nh; ++hf)
        for (int c = 0; c < 2; ++c) tma_load_2d(sbase + OFF_A + hf * TILE_BYTES + c * CHUNK, &map_a, c * 64, (bp * NH + hf) * TT, bar(BAR_A_FULL));
      for (int it = 0; it < ntiles; ++it) {
        const int s = it % W_STAGES;
        mbar_wait(bar(BAR_W_EMPTY + s), ((it / W_STAGES) & 1) ^ 1);
        mbar_expect_tx(bar(BAR_W_FULL + s), TILE_BYTES);
        for (int c = 0; c < 2; ++c) tma_load_2d(sbase + OFF_W + s * TILE_BYTES + c * CHUNK, &map_w, c * 64, (t0 + it) * TX, bar(BAR_W_FULL + s));
      }
    }
  } else if (warp == WARP_MMA) {
    // the whole warp runs this loop converged; one elected lane issues the MMAs and the commits (tc_common.cuh: elect_one)
    mbar_wait(bar(BAR_A_FULL), 0);
    for (int it = 0; it < ntiles; ++it) {
      const int s = it % W_STAGES;
      mbar_wait(bar(BAR_W_FULL + s), (it / W_STAGES) & 1);
      const uint64_t d_w = smem_desc(sbase + OFF_W + s * TILE_BYTES, 16, 1024);
      for (int hf = 0; hf < nh; ++hf) {
        const int q = it * nh + hf, zs = q % Z_STAGES;  // accumulator stages are handed out per product
        if (g.timing && lane == 0 && q < 60) g.timing[256 * blockIdx.x + 8 + 4 * q] = clock64();
        mbar_wait(bar(BAR_Z_EMPTY + zs), ((q / Z_STAGES) & 1) ^ 1);
        tc_fence_after();
        const uint64_t d_a = smem_desc(sbase + OFF_A + hf * TILE_BYTES, 16, 1024);
#pragma unroll
        for (int i = 0; i < HK / 16; ++i) {  // 8 k-steps of 16 halfs: chunk i/4, 32-byte slice i%4 of the 128-byte swizzle row
          const uint64_t koff = (uint64_t)(((i >> 2) * CHUNK + (i & 3) * 32) >> 4);
          if (elect_one()) mma_f16(tmem + zs * TX, d_a + koff, d_w + koff, IDESC, i > 0);
        }
        if (elect_one()) tc_commit(bar(BAR_Z_FULL + zs));
        if (g.timing && lane == 0 && q < 60) g.timing[256 * blockIdx.x + 8 + 4 * q + 1] = clock64();
      }
      if (elect_one()) tc_commit(bar(BAR_W_EMPTY + s));
    }
  } else {
    // ---- epilogue: thread = (team = TMEM lane, column block cb of 32 experts) ----
    const int qd = warp & 3, cb = warp >> 2;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    int team[NH];
    bool team_ok[NH];
    float Tz[NH];
    int bK[NH];
#pragma unroll
    for (int hf = 0; hf < NH; ++hf) {
      team[hf] = (bp * NH + hf) * TT + qd * 32 + lane;
      team_ok[hf] = hf < nh && team[hf] < g.B;
      Tz[hf] = 0.f; bK[hf] = 0;
      if (PASS == 2 && team_ok[hf]) { Tz[hf] = __ldg(g.thr_val + team[hf]); bK[hf] = __ldg(g.thr_blk + team[hf]); }
    }
    const float NEG_INF = __int_as_float(0xff800000);
    for (int it = 0; it < ntiles; ++it) {
      const int e_base = (t0 + it) * TX + cb * BLK;  // first expert of this thread's block
      const int blk = (t0 + it) * 4 + cb;
      // the block's biases (the same 32 addresses in every lane: broadcast loads); past E: -inf so that the column can never win
      float bv[BLK];
      if (e_base + BLK <= g.E) {
#pragma unroll
        for (int u = 0; u < BLK / 4; ++u) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(g.bias + e_base) + u);
          bv[4 * u] = f.x; bv[4 * u + 1] = f.y; bv[4 * u + 2] = f.z; bv[4 * u + 3] = f.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < BLK; ++i) bv[i] = (e_base + i < g.E) ? __ldg(g.bias + e_base + i) : NEG_INF;
      }
#pragma unroll
      for (int hf = 0; hf < NH; ++hf) {
        if (hf >= nh) break;
        const int q = it * nh + hf, zs = q % Z_STAGES;
        mbar_wait(bar(BAR_Z_FULL + zs), (q / Z_STAGES) & 1);
        if (g.timing && threadIdx.x == 0 && q < 60) g.timing[256 * blockIdx.x + 8 + 4 * q + 2] = clock64();
        tc_fence_after();
        float z[BLK];
        tmem_ld32(tmem + lane_base + zs * TX + cb * BLK, z);
        tc_fence_before();
        mbar_arrive(bar(BAR_Z_EMPTY + zs));
        float m = NEG_INF;
#pragma unroll
        for (int i = 0; i < BLK; ++i) { z[i] += bv[i]; m = fmaxf(m, z[i]); }
        if (PASS == 1) {
          if (team_ok[hf]) g.bm[(size_t)team[hf] * g.nblk + blk] = m;
        } else {
          const float tz = Tz[hf];
          const bool tie_ok = blk <= bK[hf];
          if (team_ok[hf] && m >= tz && (m > tz || tie_ok)) {  // rare: K of the E/32 blocks of a team get here
#pragma unroll
            for (int i = 0; i < BLK; ++i) {
              const float v = z[i];
              if (v > tz || (v == tz && tie_ok)) {
                const int slot = atomicAdd(g.cnt + team[hf], 1);
                if (slot < g.cap) g.cand[(size_t)team[hf] * g.cap + slot] = ((unsigned long long)ordered_key(v) << 32) | (uint32_t)(~(uint32_t)(e_base + i));
              }
            }
          }
        }
        if (g.timing && threadIdx.x == 0 && q < 60) g.timing[256 * blockIdx.x + 8 + 4 * q + 3] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (g.timing && threadIdx.x == 0) g.timing[256 * blockIdx.x + 1] = clock64();
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// per team: the candidates in rank order -> the first K as (probability, global expert id).  One CTA per team; bitonic sort in shared memory.
constexpr int FIN_THREADS = 128;
__global__ void __launch_bounds__(FIN_THREADS) infer_topk_final_kernel(const unsigned long long* __restrict__ cand, const int* __restrict__ cnt, int cap, int K,
                                                                        int e_lo, float* __restrict__ vals, int32_t* __restrict__ idx) {
  extern __shared__ unsigned long long sel[];
  const int n = blockIdx.x;
  const int c = min(cnt[n], cap);
  int npad = 32;
  while (npad < c) npad <<= 1;
  for (int i = threadIdx.x; i < npad; i += FIN_THREADS) sel[i] = i < c ? cand[(size_t)n * cap + i] : 0ull;
  __syncthreads();
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npad; i += FIN_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = sel[i], y = sel[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (x < y) : (x > y)) { sel[i] = y; sel[ixj] = x; }
        }
      }
      __syncthreads();
    }
  constexpr float LOG2E = 1.4426950408889634f;
  for (int i = threadIdx.x; i < K; i += FIN_THREADS) {
    const unsigned long long cmp = i < c ? sel[i] : 0ull;
    float p = 0.f;
    int32_t e = -1;
    if (cmp) {
      const float zz = key_to_float((uint32_t)(cmp >> 32));
      const float x = fmaxf(zz, NTF_LRELU_SLOPE * zz);           // the arithmetic of ntf_infer_scores' tensor-core epilogue (out_tc.cu, MODE 1)
      const float ex = ex2_approx(fabsf(x) * -LOG2E);
      const float r = rcp_approx(1.f + ex);
      p = zz > 0.f ? r : ex * r;
      e = e_lo + (int32_t)(~(uint32_t)cmp);
    }
    vals[(size_t)n * K + i] = p;
    idx[(size_t)n * K + i] = e;
  }
}

// per team (one CTA of 4 warps): Tz = the K-th largest of the row's block maxima and bK = the block that holds it under the rank order (value
// descending, ties -> lower block first).  Bisection on the order-preserving integer image of the fp32 maxima: 32 rounds of "how many keys
// >= candidate" (register-resident keys, PL per thread in contiguous chunks so that ties can be ranked by block id; one barrier per round).
template <int PL>
__global__ void __launch_bounds__(128) blockmax_threshold_kernel(const float* __restrict__ bm, int nblk, int K, float* __restrict__ thr_val,
                                                                 int32_t* __restrict__ thr_blk) {
  __shared__ int cnt[2][4];
  __shared__ int wsum[2][4];
  const int team = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int per = (nblk + 127) / 128;  // <= PL (checked by the host)
  const float* row = bm + (size_t)team * nblk;
  const int first = tid * per;
  uint32_t key[PL];
#pragma unroll
  for (int i = 0; i < PL; ++i) key[i] = (i < per && first + i < nblk) ? ordered_key(__ldg(row + first + i)) : 0u;  // 0 < every real key
  uint32_t T = 0u;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    int c0 = 0, c1 = 0;
#pragma unroll
    for (int i = 0; i < PL; i += 2) { c0 += key[i] >= cand; if (i + 1 < PL) c1 += key[i + 1] >= cand; }
    const int c = __reduce_add_sync(0xffffffffu, c0 + c1);
    if (lane == 0) cnt[bit & 1][w] = c;
    __syncthreads();  // (the buffer of round r is rewritten in round r+2, after everybody passed round r+1's barrier)
    if (cnt[bit & 1][0] + cnt[bit & 1][1] + cnt[bit & 1][2] + cnt[bit & 1][3] >= K) T = cand;
  }
  int gt = 0, eq = 0;
#pragma unroll
  for (int i = 0; i < PL; ++i) { gt += key[i] > T; eq += key[i] == T; }
  const int gtw = __reduce_add_sync(0xffffffffu, gt);
  int incl = eq;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 0) wsum[0][w] = gtw;
  if (lane == 31) wsum[1][w] = incl;
  __syncthreads();
  const int need = K - (wsum[0][0] + wsum[0][1] + wsum[0][2] + wsum[0][3]);  // the need-th block (in id order) whose maximum equals Tz is the K-th block
  int before = incl - eq;
  for (int q = 0; q < w; ++q) before += wsum[1][q];
  if (before < need && need <= before + eq) {
    int seen = before, blk = first;
#pragma unroll
    for (int i = 0; i < PL; ++i)
      if (key[i] == T && ++seen == need) blk = first + i;
    thr_val[team] = key_to_float(T);
    thr_blk[team] = blk;
  }
}

__global__ void to_half_kernel(const float* __restrict__ x, size_t n, __half* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = __float2half_rn(x[i]);
}

struct ItWs { size_t a16, bm, tv, ti, cnt, cand, total; };
ItWs it_ws(int B, int h, int E, int K) {
  const size_t nblk = (size_t)cdiv(E, TX) * 4;
  ItWs w;
  w.a16 = 0;
  w.bm = w.a16 + align_up((size_t)B * h * sizeof(__half), 1024);
  w.tv = w.bm + align_up((size_t)B * nblk * sizeof(float), 1024);
  w.ti = w.tv + align_up((size_t)B * sizeof(float), 1024);
  w.cnt = w.ti + align_up((size_t)B * sizeof(int32_t), 1024);
  w.cand = w.cnt + align_up((size_t)B * sizeof(int), 1024);
  w.total = w.cand + align_up((size_t)B * (size_t)(BLK * K) * sizeof(unsigned long long), 1024);
  return w;
}
}  // namespace

extern "C" int ntf_to_half(ntf_ctx* ctx, void* stream, const float* x, size_t n, void* y) {
  NTF_REQUIRE(ctx && x && y, NTF_ERR_BAD_ARG, "to_half: null pointer");
  if (n == 0) return NTF_OK;
  const size_t blocks = (n + 255) / 256;
  NTF_COUNT_LAUNCH; to_half_kernel<<<(unsigned)(blocks < (size_t)ctx->sm_count * 8 ? blocks : (size_t)ctx->sm_count * 8), 256, 0, as_stream(stream)>>>(x, n, (__half*)y);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

// the fused path needs the tensor-core width, at least K blocks of 32 experts per team, and a candidate list that sorts in shared memory
extern "C" int ntf_infer_topk_supported(int B, int h, int E, int K) {
  return (h == HK && B >= 1 && K >= 1 && K <= 128 && (long long)K * BLK <= (long long)E && cdiv(cdiv(E, TX) * 4, 128) <= 32) ? 1 : 0;  // (E <= 131072 per shard)
}

extern "C" size_t ntf_infer_topk_workspace_bytes(int B, int h, int E, int K) { return it_ws(B, h, E, K).total; }

extern "C" int ntf_infer_topk(ntf_ctx* ctx, void* stream, const ntf_infer_topk_args* a, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ctx && a && a->W16 && a->b && a->vals && a->idx && (a->A || a->A16), NTF_ERR_BAD_ARG, "infer_topk: null pointer");
  NTF_REQUIRE(ntf_infer_topk_supported(a->B, a->h, a->E, a->K), NTF_ERR_UNSUPPORTED, "infer_topk: B=%d h=%d E=%d K=%d (needs h=%d, K <= 128, 32*K <= E <= 131072)", a->B, a->h, a->E,
              a->K, HK);
  const ItWs w = it_ws(a->B, a->h, a->E, a->K);
  NTF_REQUIRE(workspace && workspace_bytes >= w.total, NTF_ERR_WORKSPACE, "infer_topk: workspace too small");
  NTF_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)a->W16 & 15) == 0 && ((uintptr_t)a->b & 15) == 0, NTF_ERR_BAD_ARG, "infer_topk: misaligned pointer");
  cudaStream_t st = as_stream(stream);
  char* ws = (char*)workspace;
  const __half* A16 = (const __half*)a->A16;
  if (!A16) {
    NTF_REQUIRE(((uintptr_t)a->A & 15) == 0, NTF_ERR_BAD_ARG, "infer_topk: misaligned activations");
    int rc = ntf_to_half(ctx, stream, a->A, (size_t)a->B * a->h, ws + w.a16);
    if (rc) return rc;
    A16 = (const __half*)(ws + w.a16);
  }
  NTF_REQUIRE(((uintptr_t)A16 & 15) == 0, NTF_ERR_BAD_ARG, "infer_topk: misaligned fp16 activations");
  ItArgs g{};
  g.bias = a->b; g.B = a->B; g.E = a->E; g.K = a->K;
  g.nbt = cdiv(a->B, TT); g.nct = cdiv(a->E, TX);
  const int sm = ctx->sm_count > 0 ? ctx->sm_count : 148;
  const int nbp = cdiv(g.nbt, NH);
  g.nchunk = nbp >= sm ? 1 : sm / nbp;  // one wave: batch-tile pairs x chunks <= SMs
  if (g.nchunk > g.nct) g.nchunk = g.nct;
  g.nblk = g.nct * 4;
  g.bm = (float*)(ws + w.bm);
  float* tv = (float*)(ws + w.tv);
  int32_t* ti = (int32_t*)(ws + w.ti);
  g.thr_val = tv; g.thr_blk = ti;
  g.cnt = (int*)(ws + w.cnt);
  g.cand = (unsigned long long*)(ws + w.cand);
  g.cap = BLK * a->K;
  const char* tim = getenv("NTF_IT_TIMING");
  g.timing = tim ? (long long*)(uintptr_t)strtoull(tim, nullptr, 0) : nullptr;
  CUtensorMap ma, mw;
  int rc;
  if ((rc = make_map(ctx, &ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, A16, (uint64_t)a->B, HK, TT, 64))) return rc;
  if ((rc = make_map(ctx, &mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a->W16, (uint64_t)a->E, HK, TX, 64))) return rc;
  const int grid = nbp * g.nchunk;
  static bool attr_set[64] = {};  // (per device: the attribute call costs about as much as a launch)
  if (!attr_set[ctx->device & 63]) {
    NTF_CUDA(cudaFuncSetAttribute(infer_topk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    NTF_CUDA(cudaFuncSetAttribute(infer_topk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set[ctx->device & 63] = true;
  }
  NTF_CUDA(cudaMemsetAsync(g.cnt, 0, (size_t)a->B * sizeof(int), st));
  NTF_COUNT_LAUNCH; infer_topk_kernel<1><<<grid, NT, SMEM_BYTES, st>>>(ma, mw, g);
  NTF_LAUNCH_CHECK();
  {
    const int per = cdiv(g.nblk, 128);
    NTF_COUNT_LAUNCH;
    if (per <= 4) blockmax_threshold_kernel<4><<<a->B, 128, 0, st>>>(g.bm, g.nblk, a->K, tv, ti);
    else if (per <= 12) blockmax_threshold_kernel<12><<<a->B, 128, 0, st>>>(g.bm, g.nblk, a->K, tv, ti);
    else blockmax_threshold_kernel<32><<<a->B, 128, 0, st>>>(g.bm, g.nblk, a->K, tv, ti);
    NTF_LAUNCH_CHECK();
  }
  NTF_COUNT_LAUNCH; infer_topk_kernel<2><<<grid, NT, SMEM_BYTES, st>>>(ma, mw, g);
  NTF_LAUNCH_CHECK();
  int npad = 32;
  while (npad < g.cap) npad <<= 1;
  NTF_COUNT_LAUNCH; infer_topk_final_kernel<<<a->B, FIN_THREADS, (size_t)npad * 8, st>>>(g.cand, g.cnt, g.cap, a->K, a->e_lo, a->vals, a->idx);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
