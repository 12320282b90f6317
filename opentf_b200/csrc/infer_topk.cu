// infer_topk.cu -- K5: test-time top-K expert ranking fused with the output layer (fnn.py:204-218, pkgmgr.py:125-134).
//
// The reference computes sigmoid(model(X)) as a dense [N,E] matrix, moves it to the host and runs torch.topk over it.  Here the
// [B,E] score matrix never exists: the logits live in TMEM for the time a CTA looks at them, and only O(K) numbers per team
// leave the SM.  Selection is done on the logit z = a.w + b (sigmoid(lrelu(.)) is monotone, so the K best logits are the K best
// probabilities; the probability is evaluated for the K winners only), in TWO passes over the same tcgen05 product:
//
//   pass 1   Z[128 teams x 128 experts] = A16 . W16^T per (batch tile, expert tile); every epilogue thread (team, block of 32
//            consecutive experts) keeps only the block's MAXIMUM -> bm[B, E/32]   (4*E/32 bytes per team instead of 4*E)
//   select   per team the K-th largest block maximum Tz under the order (value descending, block ascending) -- blockmax_threshold_kernel,
//            an adaptive radix select -- and per block a bit (one of the K blocks at or above it?) and for those a byte, the slot 0..K-1.
//            K blocks have a maximum >= Tz, so Tz is a lower bound of the K-th best logit and everything that can be in the top K
//            has z >= Tz and lies inside those K blocks (an element equal to Tz in a later block loses the tie against the K-th block's).
//   pass 2   the same product again; a thread whose block has a slot stores the block's 32 (logit, expert) composites at
//            cand[team][slot][0..31] -- fixed positions, no atomics, no counters.  A warp none of whose 32 (team, block) pairs has a
//            slot (the block bits sit in registers, 8 expert tiles at a time) does not even read its logits from TMEM.
//   final    per team: the candidates >= Tz in rank order (z descending, expert id ascending) -> the first K as (probability, expert id).
//
// Recomputing the product is cheaper than writing and re-reading [B,E] (FlashAttention's trade): per team 2*2*h*E flops on the
// tensor pipe against 8*E bytes of HBM traffic.  Exact: the result equals the K best of the full score row under the library's
// rank order (value descending, ties -> lower expert id), as ntf_topk_select gives on ntf_infer_scores' output, whenever the
// logits are distinct where the probabilities are (equal probabilities of distinct logits -- sigmoid saturation -- are ranked
// by logit: a permutation among exact score ties).
//
// Layout: CTA = (NH = 4 batch tiles of 128 teams, chunk of consecutive expert tiles); grid = groups x chunks sized to one wave.
//   warp 16       TMA: the fp16 activation tiles once, then the fp16 W tiles of the chunk through a 3-stage ring
//   warps 17-18   MMA issuers: per W tile one product per batch tile (M = 128 teams, N = 128 experts, K = 128 hidden: 8 x kind::f16),
//                 4 accumulators in TMEM
//   warps 0-15    epilogue: thread = (team = TMEM lane, 32-expert column block): tcgen05.ld, + bias, block maximum / candidate store
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
constexpr int NH = 4;     // batch tiles per CTA: every W tile that comes in is multiplied with all of them.  The W traffic out of L2 bounds the
                          // kernel otherwise: one tile per CTA = 8 x 11.5 MB per pass on the C2 shape (18.6 us measured), two = 46 MB (~12 us);
                          // four fill the 512 TMEM columns (one accumulator each) and the shared memory (128 KB of A + 3 W stages)
constexpr uint32_t OFF_A = 0;
constexpr uint32_t OFF_W = OFF_A + NH * TILE_BYTES;
constexpr uint32_t OFF_BAR = OFF_W + W_STAGES * TILE_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256;
enum { BAR_A_FULL = 0, BAR_W_FULL = NH, BAR_W_EMPTY = NH + W_STAGES, BAR_Z_FULL = NH + 2 * W_STAGES, BAR_Z_EMPTY = NH + 2 * W_STAGES + Z_STAGES,
       NUM_BARS = NH + 2 * W_STAGES + 2 * Z_STAGES };
static_assert(NUM_BARS * 8 + 8 <= 256, "barrier block");
constexpr int EPI_WARPS = 16, WARP_TMA = 16, WARP_MMA = 17, N_MMA = 2, NT = (17 + N_MMA) * 32;  // warps 17, 18 issue the products of the even / odd batch tiles of the CTA
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr uint32_t IDESC = instr_desc(0, 0, 0, TT, TX);  // Z = A16(K-major) . W16(K-major)^T, fp16 operands, fp32 accumulate
constexpr int BLK = 32;  // experts per block (= the 32 TMEM columns one epilogue thread reads)

// Programmatic dependent launch: every kernel of the chain lets the next one start its CTAs (launch latency, barrier / TMEM set-up) while it is
// still running, and itself touches global memory only after the kernel before it has completed and flushed (griddepcontrol.wait; without
// the launch attribute both are no-ops).  The chain is transitive: a kernel that has completed has seen its predecessor complete.
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
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
  const uint32_t* bits;  // pass 2 in: [nwords, B] bit blk % 32 of word blk / 32 = the block is one of the team's K best blocks (blockmax_threshold_kernel), and
  const uint16_t* slot;  //            [B, nblk] for those blocks: the slot 0..K-1 among them (the other entries are not written; slot_blk[B, K] is the inverse)
  int nwords;
  float* cand;           // pass 2 out: [B, K, 32] the products a.w (no bias yet) of the team's K best blocks, slot by slot; every entry written
  int timing_pass;
  long long* timing;         // debug (NTF_IT_TIMING): per CTA 256 slots: [0] start clock, [1] end clock, [2] [3] the same in globaltimer ns; from 8, 4 per product q < 60: W tile landed (MMA warp),
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
  long long* const timing = g.timing_pass == PASS ? g.timing : nullptr;
  if (timing && threadIdx.x == 0) { timing[256 * blockIdx.x] = clock64(); timing[256 * blockIdx.x + 2] = (long long)globaltimer_ns(); }

  if (threadIdx.x == 0) {
    for (int hf = 0; hf < NH; ++hf) mbar_init(bar(BAR_A_FULL + hf), 1);
    for (int s = 0; s < W_STAGES; ++s) { mbar_init(bar(BAR_W_FULL + s), 1); mbar_init(bar(BAR_W_EMPTY + s), min(nh, N_MMA)); }  // (one release per issuing warp)
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
  griddep_launch();
  griddep_wait();  // (everything above is set-up; the activations, the select's outputs and the workspace belong to the kernels before)

  if (warp == WARP_TMA) {
    if (lane == 0) {
      for (int hf = 0; hf < nh; ++hf) {  // the first product needs batch tile 0 and W tile 0 only: one barrier per batch tile, W tile 0 right behind the first
        mbar_expect_tx(bar(BAR_A_FULL + hf), TILE_BYTES);
        for (int c = 0; c < 2; ++c) tma_load_2d(sbase + OFF_A + hf * TILE_BYTES + c * CHUNK, &map_a, c * 64, (bp * NH + hf) * TT, bar(BAR_A_FULL + hf));
        if (hf == 0 && ntiles > 0) {
          mbar_expect_tx(bar(BAR_W_FULL), TILE_BYTES);
          for (int c = 0; c < 2; ++c) tma_load_2d(sbase + OFF_W + c * CHUNK, &map_w, c * 64, t0 * TX, bar(BAR_W_FULL));
        }
      }
      for (int it = 1; it < ntiles; ++it) {
        const int s = it % W_STAGES;
        mbar_wait(bar(BAR_W_EMPTY + s), ((it / W_STAGES) & 1) ^ 1);
        mbar_expect_tx(bar(BAR_W_FULL + s), TILE_BYTES);
        for (int c = 0; c < 2; ++c) tma_load_2d(sbase + OFF_W + s * TILE_BYTES + c * CHUNK, &map_w, c * 64, (t0 + it) * TX, bar(BAR_W_FULL + s));
      }
    }
  } else if (warp >= WARP_MMA) {
    // N_MMA issuing warps, warp m takes the batch tiles hf = m, m + N_MMA, ... of every W tile (a single warp sharing its scheduler with four
    // epilogue warps fell behind the tensor pipe: scripts/topk_timing.py).  Each runs converged; one elected lane issues the MMAs and the commits.
    const int m = warp - WARP_MMA;
    if (m < nh) {
      for (int it = 0; it < ntiles; ++it) {
        const int s = it % W_STAGES;
        mbar_wait(bar(BAR_W_FULL + s), (it / W_STAGES) & 1);
        const uint64_t d_w = smem_desc(sbase + OFF_W + s * TILE_BYTES, 16, 1024);
        for (int hf = m; hf < nh; hf += N_MMA) {
          const int q = it * nh + hf, zs = q % Z_STAGES;  // accumulator stages are handed out per product (nh = 4: stage = batch tile)
          if (it == 0) mbar_wait(bar(BAR_A_FULL + hf), 0);
          if (timing && lane == 0 && q < 60) timing[256 * blockIdx.x + 8 + 4 * q] = clock64();
          mbar_wait(bar(BAR_Z_EMPTY + zs), ((q / Z_STAGES) & 1) ^ 1);
          tc_fence_after();
          const uint64_t d_a = smem_desc(sbase + OFF_A + hf * TILE_BYTES, 16, 1024);
#pragma unroll
          for (int i = 0; i < HK / 16; ++i) {  // 8 k-steps of 16 halfs: chunk i/4, 32-byte slice i%4 of the 128-byte swizzle row
            const uint64_t koff = (uint64_t)(((i >> 2) * CHUNK + (i & 3) * 32) >> 4);
            if (elect_one()) mma_f16(tmem + zs * TX, d_a + koff, d_w + koff, IDESC, i > 0);
          }
          if (elect_one()) tc_commit(bar(BAR_Z_FULL + zs));
          if (timing && lane == 0 && q < 60) timing[256 * blockIdx.x + 8 + 4 * q + 1] = clock64();
        }
        if (elect_one()) tc_commit(bar(BAR_W_EMPTY + s));
      }
    }
  } else {
    // ---- epilogue: thread = (team = TMEM lane, column block cb of 32 experts) ----
    const int qd = warp & 3, cb = warp >> 2;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    const int team0 = bp * NH * TT + qd * 32 + lane;  // this thread's team of batch tile hf: team0 + hf * TT
    const float NEG_INF = __int_as_float(0xff800000);
    uint32_t win[NH];  // pass 2: per batch tile the team's block bits of 8 expert tiles (bit 4 * (it % 8) + cb = this thread's block of tile it)
    for (int it = 0; it < ntiles; ++it) {
      const int e_base = (t0 + it) * TX + cb * BLK;  // first expert of this thread's block
      const int blk = (t0 + it) * 4 + cb;
      if (PASS == 2 && (it & 7) == 0) {  // (once per 8 expert tiles = 32 products)
        const int bit0 = (t0 + it) * 4, wi = bit0 >> 5;
#pragma unroll
        for (int hf = 0; hf < NH; ++hf) {
          const int team = team0 + hf * TT;
          win[hf] = 0u;
          if (hf < nh && team < g.B) {
            const uint32_t* wp = g.bits + (size_t)wi * g.B + team;  // word-major: the 32 teams of a warp read 128 consecutive bytes
            win[hf] = __funnelshift_r(__ldg(wp), wi + 1 < g.nwords ? __ldg(wp + g.B) : 0u, bit0 & 31);
          }
        }
      }
      float bv[BLK];       // pass 1: the block's biases (the same 32 addresses in every lane: broadcast loads); past E: -inf, the column can never win
      uint32_t slotv[NH];  // pass 2: the slot of this thread's block per batch tile, fetched here, used 1..4 products later
      if (PASS == 1) {
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
      } else {
#pragma unroll
        for (int hf = 0; hf < NH; ++hf) {
          slotv[hf] = 0u;
          if ((win[hf] >> (4 * (it & 7) + cb)) & 1u) slotv[hf] = __ldg(g.slot + (size_t)(team0 + hf * TT) * g.nblk + blk);
        }
      }
#pragma unroll
      for (int hf = 0; hf < NH; ++hf) {
        if (hf >= nh) break;
        const int q = it * nh + hf, zs = q % Z_STAGES;
        const int team = team0 + hf * TT;
        const bool team_ok = team < g.B;
        const bool mine = PASS == 2 && ((win[hf] >> (4 * (it & 7) + cb)) & 1u) != 0u;  // (0 for teams past B)
        mbar_wait(bar(BAR_Z_FULL + zs), (q / Z_STAGES) & 1);
        if (timing && threadIdx.x == 0 && q < 60) timing[256 * blockIdx.x + 8 + 4 * q + 2] = clock64();
        // The TMEM read port (64 B per cycle: 1024 cycles for a product's 64 KB of fp32 logits against 512 of tensor time) paces pass 1.
        // Pass 2 reads a warp's [32 teams x 32 experts] only when one of its 32 (team, block) pairs is among the K best: ~K/44 of the warps.
        if (PASS == 2 && !__any_sync(0xffffffffu, mine)) {
          mbar_arrive(bar(BAR_Z_EMPTY + zs));
          continue;
        }
        tc_fence_after();
        float z[BLK];
        tmem_ld32(tmem + lane_base + zs * TX + cb * BLK, z);
        tc_fence_before();
        mbar_arrive(bar(BAR_Z_EMPTY + zs));
        if (PASS == 1) {
          float m = NEG_INF;
#pragma unroll
          for (int i = 0; i < BLK; ++i) m = fmaxf(m, z[i] + bv[i]);
          if (team_ok) g.bm[(size_t)team * g.nblk + blk] = m;
        } else if (mine) {  // K of the E/32 blocks of a team: the block's 32 raw products, 128 contiguous bytes (the final kernel adds the biases)
          float4* dst = reinterpret_cast<float4*>(g.cand + ((size_t)team * g.K + slotv[hf]) * BLK);
#pragma unroll
          for (int i = 0; i < BLK; i += 4) dst[i >> 2] = make_float4(z[i], z[i + 1], z[i + 2], z[i + 3]);
        }
        if (timing && threadIdx.x == 0 && q < 60) timing[256 * blockIdx.x + 8 + 4 * q + 3] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (timing && threadIdx.x == 0) { timing[256 * blockIdx.x + 1] = clock64(); timing[256 * blockIdx.x + 3] = (long long)globaltimer_ns(); }
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// Candidate i of a team: slot i / 32, expert slot_blk[slot] * 32 + i % 32; its logit = the stored product + the expert's bias (the same fp32 add
// as pass 1's); it counts if the logit is >= the team's Tz and the expert exists; composite = ordered logit << 32 | ~expert (0 = none).

// per team (K > 32): the candidates >= Tz in rank order -> the first K as (probability, global expert id).  One CTA per team.  A warp takes a
// slot (a block of 32 candidates) at a time; the survivors (a few per block) are appended to shared memory, one atomic per warp and slot.
// `raise_bar`: an adaptive radix select over the UNIQUE composites -- 256-bin histograms of the current range, one sweep per round -- that
// finds a bar with K <= #(composites >= bar) <= limit.  It is used twice: over the candidates in global memory when more than SEL_CAP
// survive (K near 1000 on a flat score distribution, a plateau of equal scores: rare), and over the survivors in shared memory to cut them
// down to the next power of two above K before the bitonic sort (sorting 2 048 survivors to emit 1 000 cost more than everything else).
constexpr int FIN_THREADS = 256, SEL_CAP = 4096, KMAX = 1024;
__global__ void __launch_bounds__(FIN_THREADS) infer_topk_final_kernel(const float* __restrict__ cand, const int32_t* __restrict__ slot_blk,
                                                                       const float* __restrict__ bias, const float* __restrict__ thr_val, int cap, int K, int E,
                                                                       int e_lo, float* __restrict__ vals, int32_t* __restrict__ idx) {
  __shared__ unsigned long long sel[SEL_CAP];
  __shared__ unsigned long long top[KMAX];
  __shared__ int sblk[KMAX];  // the block of every slot
  __shared__ uint32_t hist[256];
  __shared__ unsigned long long bar_s;
  __shared__ int nsel, need_s, done_s;
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = FIN_THREADS / 32;
  if (tid == 0) nsel = 0;
  griddep_launch();
  griddep_wait();
  for (int i = tid; i < K; i += FIN_THREADS) sblk[i] = __ldg(slot_blk + (size_t)n * K + i);
  const uint32_t kz = ordered_key(thr_val[n]);
  __syncthreads();
  // one sweep over the stored candidates: fn(composite), 0 for a candidate that does not count; a warp calls fn together (collectives allowed)
  auto sweep_global = [&](auto fn) {
    constexpr int U = 8;  // slots per warp and round: their 2 U loads are in flight together (a chain of memory latencies otherwise)
    for (int s0 = warp; s0 < K; s0 += NW * U) {
      float zz[U], bb[U];
      int ee[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int sv = min(s0 + u * NW, K - 1);
        ee[u] = sblk[sv] * BLK + lane;
        zz[u] = __ldg(cand + ((size_t)n * K + sv) * BLK + lane);
        bb[u] = __ldg(bias + min(ee[u], E - 1));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t key = ordered_key(zz[u] + bb[u]);
        fn((s0 + u * NW < K && ee[u] < E && key >= kz) ? ((unsigned long long)key << 32) | (uint32_t)(~(uint32_t)ee[u]) : 0ull);
      }
    }
  };
  auto sweep_sel = [&](int c, auto fn) {  // the same over the survivors in shared memory
    for (int i0 = 0; i0 < c; i0 += FIN_THREADS) fn(i0 + tid < c ? sel[i0 + tid] : 0ull);
  };
  // append the composites that pass `keep` to dst: one shared-memory atomic per warp and call
  auto append = [&](unsigned long long* dst, int limit, unsigned long long x, bool keep) {
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m == 0u) return;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&nsel, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    const int pos = base + __popc(m & ((1u << lane) - 1u));
    if (keep && pos < limit) dst[pos] = x;
  };
  // -> the largest bar found with K <= #(composites >= bar) <= limit (limit >= K; the composites are unique, so a bin of width 1 ends it)
  auto raise_bar = [&](auto sweep, int limit) {
    unsigned long long lo = (unsigned long long)kz << 32, hi = ~0ull;  // the K-th largest composite = the need-th largest inside [lo, hi]
    int need = K;
    for (;;) {
      const unsigned long long width = hi - lo;
      const int shift = width < 256ull ? 0 : 56 - __clzll((long long)width);  // (width >> shift) < 256
      hist[tid] = 0u;  // (FIN_THREADS = 256 bins)
      __syncthreads();
      sweep([&](unsigned long long x) {
        if (x != 0ull && x >= lo && x <= hi) atomicAdd(&hist[(int)((x - lo) >> shift)], 1u);
      });
      __syncthreads();
      if (warp == 0) {  // from the top bin down to the one that holds the need-th largest: lane l owns bins 8l..8l+7, suffix sums over the lanes
        int cb[8], mine = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { cb[j] = (int)hist[8 * lane + j]; mine += cb[j]; }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_down_sync(0xffffffffu, incl, o);
          if (lane + o < 32) incl += t;
        }
        int cum = incl - mine;  // composites in the bins of the higher lanes
        const bool holds = (cum < need && need <= incl) || (lane == 0 && incl < need);  // (lane 0 takes it if the counts fell short: cannot happen)
        if (holds) {
          int d = 0;
#pragma unroll
          for (int j = 7; j >= 0; --j) {
            if (need <= cum + cb[j] || j == 0) { d = j; break; }
            cum += cb[j];
          }
          done_s = ((K - need) + cum + cb[d] <= limit) || shift == 0;  // (K - need composites lie above hi: all of them in the top K)
          bar_s = lo + ((unsigned long long)(8 * lane + d) << shift); need_s = need - cum;
        }
      }
      __syncthreads();
      const unsigned long long blo = bar_s;
      const bool done = done_s != 0;
      const int nn = need_s;
      __syncthreads();
      if (done) return blo;
      const unsigned long long t = blo + ((1ull << shift) - 1ull);
      hi = t < hi ? t : hi; lo = blo; need = nn;
    }
  };
  sweep_global([&](unsigned long long x) { append(sel, SEL_CAP, x, x != 0ull); });
  __syncthreads();
  int c = nsel;
  if (c > SEL_CAP) {  // (block-uniform) rare
    const unsigned long long bar = raise_bar(sweep_global, SEL_CAP);
    if (tid == 0) nsel = 0;
    __syncthreads();
    sweep_global([&](unsigned long long x) { append(sel, SEL_CAP, x, x != 0ull && x >= bar); });
    __syncthreads();
    c = min(nsel, SEL_CAP);
  }
  int kpad = 32;
  while (kpad < K) kpad <<= 1;
  unsigned long long* srt = sel;
  if (c > kpad) {  // more survivors than the sort needs: keep those at or above a bar with K <= count <= kpad
    const unsigned long long bar = raise_bar([&](auto fn) { sweep_sel(c, fn); }, kpad);
    if (tid == 0) nsel = 0;
    __syncthreads();
    sweep_sel(c, [&](unsigned long long x) { append(top, kpad, x, x != 0ull && x >= bar); });
    __syncthreads();
    c = min(nsel, kpad);
    srt = top;
  }
  int npad = 32;
  while (npad < c) npad <<= 1;
  for (int i = c + tid; i < npad; i += FIN_THREADS) srt[i] = 0ull;
  __syncthreads();
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npad; i += FIN_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = srt[i], y = srt[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (x < y) : (x > y)) { srt[i] = y; srt[ixj] = x; }
        }
      }
      __syncthreads();
    }
  constexpr float LOG2E = 1.4426950408889634f;
  for (int i = tid; i < K; i += FIN_THREADS) {
    const unsigned long long cmp = i < c ? srt[i] : 0ull;
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

// per team (one CTA of 4 warps): Tz = the K-th largest of the row's block maxima under the order (value descending, block ascending), and for
// every block whether it is one of the K at or above it (a bit, bits[blk / 32][team]) and, if so, its slot 0..K-1 (16 bits, slot[team][blk];
// slot_blk[team][slot] = the block).
// Adaptive radix select on the order-preserving integer image of the maxima: the keys stay in registers (PL per thread); a round histograms
// the ones inside the current range [lo, hi] into 256 equal bins of that range (shared-memory atomics -- the range starts at the row's
// min..max, so the keys spread over the bins), every warp scans the bins from the top for the one that holds the wanted rank, and the
// range shrinks to that bin.  As soon as the bin holds <= 128 keys, their (key, block) composites are gathered and ranked by counting; on
// logits one round is the rule.  More than 128 EQUAL keys (a bin of width 1): the select restarts on the block numbers of those keys.
// K <= 1024 (ntf_infer_topk_supported).
template <int PL>
__global__ void __launch_bounds__(128) blockmax_threshold_kernel(const float* __restrict__ bm, int nblk, int nwords, int K, float* __restrict__ thr_val,
                                                                 uint16_t* __restrict__ slot, uint32_t* __restrict__ bits, int32_t* __restrict__ slot_blk) {
  __shared__ __align__(16) uint32_t hist[3][256];
  __shared__ uint32_t red[2][4];
  __shared__ unsigned long long list[128];
  __shared__ unsigned long long kth;
  __shared__ int nlist, nslot;
  const int team = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* row = bm + (size_t)team * nblk;
  griddep_launch();
  griddep_wait();
  uint32_t key[PL];  // 0 = no such block (< every real key)
  uint32_t v[PL];    // what the select runs on: the key, later (ties) 4096 - block of the keys equal to Tz; 0 = not taking part
  uint32_t lo = ~0u, hi = 0u;
#pragma unroll
  for (int i = 0; i < PL; ++i) {
    const int blk = i * 128 + tid;
    key[i] = blk < nblk ? ordered_key(__ldg(row + blk)) : 0u;
    v[i] = key[i];
    if (blk < nblk) { lo = min(lo, key[i]); hi = max(hi, key[i]); }
  }
  for (int i = tid; i < 3 * 256; i += 128) (&hist[0][0])[i] = 0u;
  if (tid == 0) { nlist = 0; nslot = 0; }
  lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
  if (lane == 0) { red[0][w] = lo; red[1][w] = hi; }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) { lo = min(lo, red[0][q]); hi = max(hi, red[1][q]); }
  int need = K;        // the wanted element is the need-th largest of the v inside [lo, hi]
  uint32_t tie_key = 0u;  // != 0: the select is on block numbers among the keys equal to this one
#pragma unroll 1
  for (int r = 0;; ++r) {
    const uint32_t width = hi - lo;
    const int shift = width < 256u ? 0 : 24 - __clz((int)width);  // (width >> shift) < 256
    uint32_t* h = hist[r % 3];
    uint32_t* hn = hist[(r + 1) % 3];  // last read in round r-2: every warp is past that scan (it passed round r-1's barrier)
    hn[tid] = 0u; hn[tid + 128] = 0u;
    const uint32_t hbase = smem_u32(h);
#pragma unroll
    for (int i = 0; i < PL; ++i) {  // (a predicated red.shared: no branch per key)
      const uint32_t d = v[i] - lo;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.le.u32 p, %1, %2;\n\t@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(hbase + ((d >> shift) << 2)), "r"(d), "r"(width) : "memory");
    }
    __syncthreads();
    // every warp: lane l owns bins 8l..8l+7; suffix sums from the top bin down
    const uint4 h0 = reinterpret_cast<const uint4*>(h)[2 * lane], h1 = reinterpret_cast<const uint4*>(h)[2 * lane + 1];
    const int c[8] = {(int)h0.x, (int)h0.y, (int)h0.z, (int)h0.w, (int)h1.x, (int)h1.y, (int)h1.z, (int)h1.w};
    const int mine = c[0] + c[1] + c[2] + c[3] + c[4] + c[5] + c[6] + c[7];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += t;
    }
    int above = incl - mine;  // elements in the bins of the higher lanes
    const bool holds = above < need && need <= incl;
    int bin = 0, cbin = 0;
    if (holds) {
#pragma unroll
      for (int j = 7; j >= 0; --j) {
        if (above < need && need <= above + c[j]) { bin = 8 * lane + j; cbin = c[j]; break; }
        above += c[j];
      }
    }
    const int src = __ffs(__ballot_sync(0xffffffffu, holds)) - 1;  // (exactly one lane: K <= the number of blocks)
    bin = __shfl_sync(0xffffffffu, bin, src); cbin = __shfl_sync(0xffffffffu, cbin, src); above = __shfl_sync(0xffffffffu, above, src);
    need -= above;
    lo += (uint32_t)bin << shift;
    hi = min(hi, lo + ((1u << shift) - 1u));
    if (cbin <= 128) break;
    if (shift == 0) {  // > 128 keys equal to lo: the need-th of them in block order, i.e. the need-th largest of 4096 - block
      tie_key = lo;
#pragma unroll
      for (int i = 0; i < PL; ++i) v[i] = key[i] == tie_key ? (uint32_t)(4096 - (i * 128 + tid)) : 0u;
      lo = 1u; hi = 4096u;
    }
  }
#pragma unroll
  for (int i = 0; i < PL; ++i)
    if (v[i] - lo <= hi - lo) list[atomicAdd(&nlist, 1)] = ((unsigned long long)key[i] << 12) | (unsigned)(4095 - (i * 128 + tid));
  __syncthreads();
  const int n = nlist;
  if (tid < n) {
    const unsigned long long x = list[tid];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += list[j] > x;
    if (rank == need - 1) {
      thr_val[team] = key_to_float((uint32_t)(x >> 12));
      kth = x;
    }
  }
  __syncthreads();
  const unsigned long long x = kth;  // exactly K composites are >= it: they get the slots 0..K-1 (in any order)
#pragma unroll
  for (int i = 0; i < PL; ++i) {
    const int blk = i * 128 + tid;
    const bool in = blk < nblk && (((unsigned long long)key[i] << 12) | (unsigned)(4095 - blk)) >= x;
    const uint32_t word = __ballot_sync(0xffffffffu, in);  // blocks 32 (4 i + w) .. + 31
    if (lane == 0 && 4 * i + w < nwords) bits[(size_t)(4 * i + w) * gridDim.x + team] = word;  // [nwords][B]: pass 2 reads it by 32 teams
    if (in) {
      const int sl = atomicAdd(&nslot, 1);
      slot[(size_t)team * nblk + blk] = (uint16_t)sl;
      slot_blk[(size_t)team * K + sl] = blk;
    }
  }
}

// K <= 32: one WARP per team, no shared memory, no barriers.  The warp keeps the K best composites, rank r in lane r (topk.cu's small-K
// list); the team's K candidate blocks are offered 32 candidates at a time, one above the current K-th is inserted by one shuffle-shift.
template <int NG>
__global__ void __launch_bounds__(128) infer_topk_final_warp_kernel(const float* __restrict__ cand, const int32_t* __restrict__ slot_blk, const float* __restrict__ bias,
                                                                     const float* __restrict__ thr_val, int cap, int K, int E, int e_lo, int B,
                                                                     float* __restrict__ vals, int32_t* __restrict__ idx) {
  const int n = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  griddep_launch();
  griddep_wait();
  if (n >= B) return;
  const uint32_t kz = ordered_key(thr_val[n]);
  const int sb = lane < K ? __ldg(slot_blk + (size_t)n * K + lane) : 0;  // lane s: the block of slot s
  unsigned long long mine = 0ull, kth = 0ull;  // (0 = empty: a real composite is never 0)
  // all 2 * 8 NG loads of the team's K <= 8 NG candidate blocks are issued before the first is used (the kernel is a chain of memory latencies otherwise)
  unsigned long long xs[8 * NG];
#pragma unroll
  for (int u = 0; u < 8 * NG; ++u) {
    const int s = min(u, K - 1);
    const int e = __shfl_sync(0xffffffffu, sb, s) * BLK + lane;
    const float zz = __ldg(cand + ((size_t)n * K + s) * BLK + lane), bb = __ldg(bias + min(e, E - 1));
    const uint32_t key = ordered_key(zz + bb);
    xs[u] = (u < K && e < E && key >= kz) ? ((unsigned long long)key << 32) | (uint32_t)(~(uint32_t)e) : 0ull;
  }
#pragma unroll
  for (int u = 0; u < 8 * NG; ++u) {
    const unsigned long long x = xs[u];
    unsigned hit = __ballot_sync(0xffffffffu, x > kth);
    while (hit) {
      const int L = __ffs(hit) - 1;
      hit &= hit - 1;
      const unsigned long long v = __shfl_sync(0xffffffffu, x, L);
      if (v > kth) {  // (an earlier insert of this round may have raised the bar)
        const unsigned long long up = __shfl_up_sync(0xffffffffu, mine, 1);
        const unsigned long long prev = lane == 0 ? ~0ull : up;
        if (v > mine) mine = v > prev ? prev : v;  // ranks below the insertion point shift down by one, the point takes v
        if (lane >= K) mine = 0ull;
        kth = __shfl_sync(0xffffffffu, mine, K - 1);
      }
    }
  }
  if (lane < K) {
    constexpr float LOG2E = 1.4426950408889634f;
    float p = 0.f;
    int32_t e = -1;
    if (mine) {
      const float zz = key_to_float((uint32_t)(mine >> 32));
      const float x = fmaxf(zz, NTF_LRELU_SLOPE * zz);           // the arithmetic of ntf_infer_scores' tensor-core epilogue (out_tc.cu, MODE 1)
      const float ex = ex2_approx(fabsf(x) * -LOG2E);
      const float r = rcp_approx(1.f + ex);
      p = zz > 0.f ? r : ex * r;
      e = e_lo + (int32_t)(~(uint32_t)mine);
    }
    vals[(size_t)n * K + lane] = p;
    idx[(size_t)n * K + lane] = e;
  }
}

__global__ void to_half_kernel(const float* __restrict__ x, size_t n, __half* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = __float2half_rn(x[i]);
}

// <<<>>> with the programmatic-stream-serialization attribute (see griddep_wait)
template <typename... KArgs, typename... Args>
cudaError_t launch_chained(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

struct ItWs { size_t a16, bm, tv, slot, bits, sblk, cand, total; };
ItWs it_ws(int B, int h, int E, int K) {
  const size_t nblk = (size_t)cdiv(E, TX) * 4;
  ItWs w;
  w.a16 = 0;
  w.bm = w.a16 + align_up((size_t)B * h * sizeof(__half), 1024);
  w.tv = w.bm + align_up((size_t)B * nblk * sizeof(float), 1024);
  w.slot = w.tv + align_up((size_t)B * sizeof(float), 1024);
  w.bits = w.slot + align_up((size_t)B * nblk * sizeof(uint16_t), 1024);
  w.sblk = w.bits + align_up((size_t)B * cdiv((int)nblk, 32) * sizeof(uint32_t), 1024);
  w.cand = w.sblk + align_up((size_t)B * K * sizeof(int32_t), 1024);
  w.total = w.cand + align_up((size_t)B * (size_t)(BLK * K) * sizeof(float), 1024);
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
  return (h == HK && B >= 1 && K >= 1 && K <= KMAX && (long long)K * BLK <= (long long)E && cdiv(cdiv(E, TX) * 4, 128) <= 32) ? 1 : 0;  // (E <= 131072 per shard)
}

extern "C" size_t ntf_infer_topk_workspace_bytes(int B, int h, int E, int K) { return it_ws(B, h, E, K).total; }

extern "C" int ntf_infer_topk(ntf_ctx* ctx, void* stream, const ntf_infer_topk_args* a, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ctx && a && a->W16 && a->b && a->vals && a->idx && (a->A || a->A16), NTF_ERR_BAD_ARG, "infer_topk: null pointer");
  NTF_REQUIRE(ntf_infer_topk_supported(a->B, a->h, a->E, a->K), NTF_ERR_UNSUPPORTED, "infer_topk: B=%d h=%d E=%d K=%d (needs h=%d, K <= 1024, 32*K <= E <= 131072)", a->B, a->h, a->E,
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
  uint16_t* slot = (uint16_t*)(ws + w.slot);
  uint32_t* bits = (uint32_t*)(ws + w.bits);
  g.slot = slot; g.bits = bits; g.nwords = cdiv(g.nblk, 32);
  g.cand = (float*)(ws + w.cand);
  int32_t* sblk = (int32_t*)(ws + w.sblk);
  const int cap = BLK * a->K;
  const char* tp = getenv("NTF_IT_TIMING_PASS");
  g.timing_pass = tp ? atoi(tp) : 2;
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
  static const int stop_after = getenv("NTF_IT_STOP") ? atoi(getenv("NTF_IT_STOP")) : 0;  // debug: leave after stage 1..3 (scripts/topk_stages.sh times the prefixes)
  static const bool pdl = !getenv("NTF_IT_PDL") || atoi(getenv("NTF_IT_PDL")) != 0;
  NTF_COUNT_LAUNCH; NTF_CUDA(launch_chained(infer_topk_kernel<1>, grid, NT, SMEM_BYTES, st, pdl, ma, mw, g));
  if (stop_after == 1) return NTF_OK;
  {
    const int per = cdiv(g.nblk, 128);  // <= 32: ntf_infer_topk_supported
    NTF_COUNT_LAUNCH;
    auto* sel = per <= 4 ? blockmax_threshold_kernel<4> : per <= 12 ? blockmax_threshold_kernel<12> : blockmax_threshold_kernel<32>;
    NTF_CUDA(launch_chained(sel, a->B, 128, 0, st, pdl, (const float*)g.bm, g.nblk, g.nwords, a->K, tv, slot, bits, sblk));
  }
  if (stop_after == 2) return NTF_OK;
  NTF_COUNT_LAUNCH; NTF_CUDA(launch_chained(infer_topk_kernel<2>, grid, NT, SMEM_BYTES, st, pdl, ma, mw, g));
  if (stop_after == 3) return NTF_OK;
  if (a->K <= 32) {
    auto* fin = a->K <= 8 ? infer_topk_final_warp_kernel<1> : a->K <= 16 ? infer_topk_final_warp_kernel<2> : a->K <= 24 ? infer_topk_final_warp_kernel<3> : infer_topk_final_warp_kernel<4>;
    NTF_COUNT_LAUNCH; NTF_CUDA(launch_chained(fin, cdiv(a->B, 4), 128, 0, st, pdl, (const float*)g.cand, (const int32_t*)sblk, a->b, (const float*)tv, cap, a->K, a->E, a->e_lo, a->B, a->vals, a->idx));
  } else {
    NTF_COUNT_LAUNCH; NTF_CUDA(launch_chained(infer_topk_final_kernel, a->B, FIN_THREADS, 0, st, pdl, (const float*)g.cand, (const int32_t*)sblk, a->b, (const float*)tv, cap, a->K, a->E, a->e_lo, a->vals, a->idx));
  }
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
