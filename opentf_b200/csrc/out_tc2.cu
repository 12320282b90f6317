// out_tc2.cu -- NTF_TF32 training / validation step of the Fnn output layer, PERSISTENT version (round 2).
//
// Same mathematics and the same per-tile pipeline as out_tc.cu's MODE 0 (fnn.py:25 last layer, 32-46, 135 and its autograd in one
// kernel; logits, losses and dz live in TMEM / registers / shared memory only) -- what changes is how the work is laid over the machine.
// Round 1 launched one CTA per expert tile: 313 CTAs over 148 SMs = 2.1 waves, each CTA paying its own prologue (fp32 W tile in, fp16
// conversion: ~5000 cycles), pipeline fill and dW drain (~4500 cycles) for 8 batch tiles of work, and every dA tile left through a
// single 16 KB staging buffer whose TMA reduce-add had to finish READING before the next chunk could be staged (measured: 4600 cycles
// per tile for that drain alone = the kernel's tile period; scripts/mb/mb_reduce.cu).  Here:
//   * one CTA per SM walks a contiguous range of the (expert tile, batch tile) sequence -- no waves, no split last wave;
//   * the W tile arrives as fp16 straight from an fp16 image of the weight (TMA, no conversion pass); the caller keeps the image
//     (ntf_out_train_args.W16, rewritten by the optimiser) or this file makes it per call;
//   * the dW drain of an expert tile is staged in the idle dz ring and handed to the TMA unit (tensor store, or reduce-add for a tile two
//     CTAs share), so the NEXT tile's W / activation loads, first forward product and epilogue run underneath it;
//   * dA is produced TRANSPOSED (hidden unit on the TMEM lane), so its warps add it into dA[B,128] with coalesced red.global.add from
//     registers -- no staging buffer, no barriers, no proxy fences -- and the 48 KB of staging go to a third activation stage, which takes
//     tile n+1's activation load off the critical path of tile n-1's backward products.
//   * the kernel is purely DENSE: every (team, expert) pair is treated as (target 0, weight tnw).  The few pairs that are not -- a team's
//     members and its sampled negatives, ~8 per team -- are put right afterwards by out_fix_kernel (one warp per team: logit by a 128-term
//     dot product of the same fp16 operands, rank-1 corrections of dW / dA / db / loss).  Round 1 fixed them up inside the epilogue from
//     bit planes; with Zipf-distributed experts that made a few warps (popular experts) several times slower than the rest, and the
//     slowest warp gates every tile.  No planes, no plane loads, no clearing stores are left in the hot loop.
// An expert tile whose batch range is cut between two CTAs adds its dW / db partials into rows zeroed beforehand; with at most two
// contributions per element (0 + a + b) the sums are order-independent, so dW / db are run-to-run deterministic whenever
// E >= 128 * (number of SMs); dA is still summed by L2 in arrival order.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

#include "tc_common.cuh"

int ntf_loss_reduce_impl(cudaStream_t st, const float* part, int n, float scale, float* loss_out);

namespace {

constexpr int TE = 128, TB = 128, HK = 128;
constexpr uint32_t CHUNK = 128 * 128;
constexpr uint32_t W16_BYTES = TE * HK * 2, A16_BYTES = TB * HK * 2, DZ_BYTES = TE * TB * 2;
constexpr uint32_t OFF_W16 = 0;
constexpr int AST = 3;                                  // activation stages: tile n+1's load must not wait for tile n-1's backward products
constexpr uint32_t OFF_A16 = OFF_W16 + W16_BYTES;
constexpr uint32_t OFF_DZ = OFF_A16 + AST * A16_BYTES;  // 2 stages; at the end of an expert tile: staging of the dW drain (64 KB)
constexpr uint32_t OFF_BAR = OFF_DZ + 2 * DZ_BYTES;
constexpr uint32_t OFF_DBS = OFF_BAR + 512;              // [4][128] fp32 scratch of the db combine
constexpr uint32_t SMEM_BYTES = OFF_DBS + 4 * TE * 4;
static_assert(SMEM_BYTES <= 232448, "over the 227 KB shared memory limit");

constexpr uint32_t TM_Z = 0, TM_DW = 256, TM_DA = 384, TM_COLS = 512;

enum { BAR_W_FULL = 0, BAR_W_EMPTY = 1, BAR_A_FULL = 2, BAR_A_EMPTY = 5, BAR_Z_FULL = 8, BAR_Z_EMPTY = 10, BAR_DZ_FULL = 12, BAR_DZ_EMPTY = 14,
       BAR_DA_FULL = 16, BAR_DA_EMPTY = 17, BAR_DW_FULL = 18, BAR_DW_EMPTY = 19, NUM_BARS = 20 };

struct Tc2Args {
  const float* bias;
  int B, E;
  float tpw, tnw, scale;
  float* dW;        // [E,128] or NULL (validation: forward + loss only)
  float* db;
  float* dA;        // [B,128], zeroed by the caller; every CTA adds its expert tiles' contribution (red.global.add)
  float* loss_part; // [gridDim.x]
  float* Zdbg;      // debug: raw logits [B,E]
  long long* timing;  // debug: per CTA TSLOTS slots: start ns, end ns, smid, start / end clock; from slot 8, 8 per tile n < 60: Z ready, tile done, drain start, drain end, bwd issue start / end, dA staged, dW products complete
  int nct, nbt;     // expert tiles, batch tiles
  unsigned* done;   // block counter of the correction pass that follows: cleared here (CTA 0), so that pass needs no launch of its own for it
  int exp;          // debug (NTF_TC2_EXP, wrong results): 1 = the dA partial sums are read from TMEM but not added to global memory
                    // (profiles/r02h_tc2_no_dA_atomics_experiment.txt: the kernel goes from 57.8 to 48.3 us)
  int split;        // 0: CTA c owns tiles [c*T/G, (c+1)*T/G) of the expert-major sequence; s > 0: expert tile c/s, batch part c%s of s
};

constexpr int TSLOTS = 512;  // debug stamps per CTA (Tc2Args::timing)
constexpr int EPI_WARPS = 16, NT = 704, WARP_DA = 16, WARP_TMA = 20, WARP_MMA = 21, EPI_THREADS = EPI_WARPS * 32;
constexpr uint32_t IDESC_FWD = instr_desc(0, 0, 0, TE, TB);
constexpr uint32_t IDESC_DW = instr_desc(0, 0, 1, TE, HK);
constexpr uint32_t IDESC_DA = instr_desc(0, 1, 1, TB, HK);

__device__ __forceinline__ void tile_range(const Tc2Args& g, int c, int G, int* L0, int* L1) {
  if (g.split == 0) {
    const long long T = (long long)g.nct * g.nbt;
    *L0 = (int)(T * c / G); *L1 = (int)(T * (c + 1) / G);
  } else {
    const int e = c / g.split, part = c % g.split;
    *L0 = e * g.nbt + part * g.nbt / g.split; *L1 = e * g.nbt + (part + 1) * g.nbt / g.split;
  }
}

// rows of dW / db that receive partial sums from two (or, split > 2, more) CTAs: zeroed here, one CTA per cut expert tile
__global__ void __launch_bounds__(256) zero_cut_tiles_kernel(Tc2Args g) {
  int L0, L1;
  tile_range(g, blockIdx.x, gridDim.x, &L0, &L1);
  // range mode: a tile is cut iff a CTA's range starts inside it (ranges are at least one expert tile long: at most one cut per tile);
  // split mode: every tile is cut `split` ways, part 0 zeroes it
  const bool mine = g.split == 0 ? (L0 < L1 && L0 % g.nbt != 0) : (g.split > 1 && blockIdx.x % g.split == 0);
  if (!mine) return;
  const int e0 = (L0 / g.nbt) * TE;
  const int rows = min(TE, g.E - e0);
  float4* w = reinterpret_cast<float4*>(g.dW + (size_t)e0 * HK);
  for (int i = threadIdx.x; i < rows * (HK / 4); i += 256) w[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < rows; i += 256) g.db[e0 + i] = 0.f;
}

__global__ void __launch_bounds__(NT, 1) out_tc2_kernel(const __grid_constant__ CUtensorMap map_w16, const __grid_constant__ CUtensorMap map_a16,
                                                        const __grid_constant__ CUtensorMap map_dw, Tc2Args g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if ((sbase & 1023u) != 0u) __trap();
  uint8_t* sgen = smem_raw;
  const uint32_t bars = sbase + OFF_BAR;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + OFF_BAR + NUM_BARS * 8);
  auto bar = [&](int i) { return bars + 8u * i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (g.timing && threadIdx.x == 0) {
    unsigned long long gt; unsigned smid;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    g.timing[TSLOTS * blockIdx.x] = (long long)gt; g.timing[TSLOTS * blockIdx.x + 2] = smid; g.timing[TSLOTS * blockIdx.x + 3] = clock64();
  }
  int L0, L1;
  tile_range(g, blockIdx.x, gridDim.x, &L0, &L1);
  if (blockIdx.x == 0 && threadIdx.x == 0 && g.done) g.done[0] = 0u;
  const bool train = g.dW != nullptr;
  const int nbt = g.nbt;

  if (threadIdx.x == 0) {
    mbar_init(bar(BAR_W_FULL), 1); mbar_init(bar(BAR_W_EMPTY), 1);
    mbar_init(bar(BAR_DA_FULL), 1); mbar_init(bar(BAR_DA_EMPTY), 128);
    mbar_init(bar(BAR_DW_FULL), 1); mbar_init(bar(BAR_DW_EMPTY), EPI_THREADS);
    for (int s = 0; s < AST; ++s) { mbar_init(bar(BAR_A_FULL + s), 1); mbar_init(bar(BAR_A_EMPTY + s), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(BAR_Z_FULL + s), 1); mbar_init(bar(BAR_Z_EMPTY + s), EPI_THREADS);
      mbar_init(bar(BAR_DZ_FULL + s), EPI_THREADS); mbar_init(bar(BAR_DZ_EMPTY + s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + NUM_BARS * 8), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Every role walks the same sequence: tile n = L - L0 is (expert tile e, batch tile t) = (L / nbt, L % nbt); an ITEM is a maximal run of
  // tiles with the same e (index j).  Ring stages / phases follow n (A, Z, dz, planes: 2 stages) or j (W image, dW accumulator: 1 stage).
  if (warp == WARP_TMA) {
    if (lane == 0) {
      int j = -1;
      for (int L = L0, n = 0; L < L1; ++L, ++n) {
        const int e = L / nbt, t = L - e * nbt;
        if (L == L0 || t == 0) {
          ++j;
          mbar_wait(bar(BAR_W_EMPTY), (j & 1) ^ 1);  // the previous item's last products have read the image
          mbar_expect_tx(bar(BAR_W_FULL), W16_BYTES);
          for (int c = 0; c < 2; ++c) tma_load_2d(sbase + OFF_W16 + c * CHUNK, &map_w16, c * 64, e * TE, bar(BAR_W_FULL));
        }
        const int s = n % AST;
        mbar_wait(bar(BAR_A_EMPTY + s), ((n / AST) & 1) ^ 1);
        mbar_expect_tx(bar(BAR_A_FULL + s), A16_BYTES);
        for (int c = 0; c < 2; ++c) tma_load_2d(sbase + OFF_A16 + s * A16_BYTES + c * CHUNK, &map_a16, c * 64, t * TB, bar(BAR_A_FULL + s));
      }
    }
  } else if (warp == WARP_MMA) {
    // the whole warp runs this loop converged; one elected lane issues the MMAs and the commits (tc_common.cuh: elect_one)
    const uint64_t d_w16_k = smem_desc(sbase + OFF_W16, 16, 1024);      // W16 as K-major operand (forward), + chunk / 32-byte slice per k-step
    const uint64_t d_w16_mn = smem_desc(sbase + OFF_W16, CHUNK, 1024);  // W16 as MN-major operand (dA^T: M = hidden), + 2048 bytes per k-step
    auto issue_fwd = [&](int n, bool item_last) {
      const int s = n & 1, sa = n % AST;
      const uint32_t ph = (n >> 1) & 1;
      mbar_wait(bar(BAR_A_FULL + sa), (n / AST) & 1);
      mbar_wait(bar(BAR_Z_EMPTY + s), ph ^ 1);
      tc_fence_after();
      const uint64_t d_a = smem_desc(sbase + OFF_A16 + sa * A16_BYTES, 16, 1024);
#pragma unroll
      for (int i = 0; i < HK / 16; ++i) {
        const uint64_t koff = (uint64_t)(((i >> 2) * CHUNK + (i & 3) * 32) >> 4);  // (the address field holds bytes >> 4)
        if (elect_one()) mma_f16(tmem + TM_Z + s * TB, d_w16_k + koff, d_a + koff, IDESC_FWD, i > 0);
      }
      if (elect_one()) {
        tc_commit(bar(BAR_Z_FULL + s));
        if (!train) {  // forward only: the operands are free once this product has read them
          tc_commit(bar(BAR_A_EMPTY + sa));
          if (item_last) tc_commit(bar(BAR_W_EMPTY));
        }
      }
    };
    int j = -1;
    for (int L = L0, n = 0; L < L1; ++L, ++n) {
      const int e = L / nbt, t = L - e * nbt;
      const bool first = (L == L0 || t == 0), last = (L == L1 - 1 || t == nbt - 1);
      if (first) {
        ++j;
        mbar_wait(bar(BAR_W_FULL), j & 1);
        issue_fwd(n, last);  // (no product of this item could be issued ahead: the image was the previous item's until now)
      }
      if (!last) issue_fwd(n + 1, (L + 1 == L1 - 1) || (t + 1 == nbt - 1));  // keeps the tensor pipe busy while the epilogue works on tile n
      if (!train) continue;
      const int s = n & 1;
      mbar_wait(bar(BAR_DZ_FULL + s), (n >> 1) & 1);
      if (first && j > 0) mbar_wait(bar(BAR_DW_EMPTY), (j - 1) & 1);  // the previous item's dW has left TMEM
      tc_fence_after();
      if (g.timing && lane == 0 && n < 60) g.timing[TSLOTS * blockIdx.x + 8 + 8 * n + 4] = clock64();
      const uint64_t d_dz_k = smem_desc(sbase + OFF_DZ + s * DZ_BYTES, 16, 1024);       // dz^T as K-major operand (K = teams)
      const uint64_t d_dz_mn = smem_desc(sbase + OFF_DZ + s * DZ_BYTES, CHUNK, 1024);   // dz as MN-major operand (M = teams)
      const int sa = n % AST;
      const uint64_t d_a_mn = smem_desc(sbase + OFF_A16 + sa * A16_BYTES, CHUNK, 1024);  // A16 as MN-major operand (N = hidden)
#pragma unroll
      for (int i = 0; i < TB / 16; ++i) {  // dW[128 j x 128 k] += dz^T[j, n] . A16[n, k]
        const uint64_t koff = (uint64_t)(((i >> 2) * CHUNK + (i & 3) * 32) >> 4);
        if (elect_one()) mma_f16(tmem + TM_DW, d_dz_k + koff, d_a_mn + (uint64_t)((i * 2048) >> 4), IDESC_DW, (!first || i > 0));
      }
      mbar_wait(bar(BAR_DA_EMPTY), (n & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < TE / 16; ++i) {  // dA^T[128 k x 128 n] = W16[j, k]^T . dz[j, n]: hidden on the TMEM lanes (see the dA warps)
        const uint64_t koff = (uint64_t)((i * 2048) >> 4);
        if (elect_one()) mma_f16(tmem + TM_DA, d_w16_mn + koff, d_dz_mn + koff, IDESC_DA, i > 0);
      }
      if (elect_one()) {
        tc_commit(bar(BAR_DA_FULL));
        tc_commit(bar(BAR_DZ_EMPTY + s));
        tc_commit(bar(BAR_A_EMPTY + sa));
        if (last) { tc_commit(bar(BAR_W_EMPTY)); tc_commit(bar(BAR_DW_FULL)); }
      }
      if (g.timing && lane == 0 && n < 60) g.timing[TSLOTS * blockIdx.x + 8 + 8 * n + 5] = clock64();
    }
  } else if (warp < EPI_WARPS) {
    // ====================== logits / loss epilogue: thread = (expert jl, block cb of 32 of the tile's 128 teams) ======================
    const int jl = threadIdx.x & 127, cb = threadIdx.x >> 7;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    float loss_total = 0.f;
    bool drain_pending = false;
    // per item (expert tile):
    int j = -1, e0 = 0, e = 0;
    bool e_ok = false, whole = false;
    float bj = 0.f, c_pos = 0.f, c_neg = 0.f, kb1 = 0.f, kb2 = 0.f;
    float acc_lin = 0.f, acc_lg = 0.f, db_acc = 0.f;
    float2 acc_lin2 = make_float2(0.f, 0.f), db_acc2 = make_float2(0.f, 0.f);
    for (int L = L0, n = 0; L < L1; ++L, ++n) {
      const int et = L / nbt, t = L - et * nbt;
      const bool first = (L == L0 || t == 0), last = (L == L1 - 1 || t == nbt - 1);
      if (first) {
        ++j;
        e0 = et * TE; e = e0 + jl; e_ok = e < g.E;
        bj = e_ok ? __ldg(g.bias + e) : 0.f;
        c_pos = e_ok ? g.tnw : 0.f; c_neg = e_ok ? g.tnw * NTF_LRELU_SLOPE : 0.f;
        kb1 = -LOG2E * bj; kb2 = -LOG2E * NTF_LRELU_SLOPE * bj;
        acc_lin = acc_lg = db_acc = 0.f;
        acc_lin2 = make_float2(0.f, 0.f); db_acc2 = make_float2(0.f, 0.f);
        whole = (t == 0) && (L + nbt <= L1);  // this CTA runs every batch tile of the expert tile: plain stores, no partial sums
      }
      const int s = n & 1;
      const uint32_t ph = (n >> 1) & 1;
      const int n0 = t * TB + cb * 32;
      mbar_wait(bar(BAR_Z_FULL + s), ph);
      if (g.timing && threadIdx.x == 0 && n < 60) g.timing[TSLOTS * blockIdx.x + 8 + 8 * n] = clock64();
      tc_fence_after();
      float z[32];
      tmem_ld32(tmem + lane_base + TM_Z + s * TB + cb * 32, z);
      tc_fence_before();
      mbar_arrive(bar(BAR_Z_EMPTY + s));
      const int nrem = g.B - n0;
      if (g.Zdbg) {
#pragma unroll
        for (int q = 0; q < 32; ++q)
          if (e_ok && q < nrem) g.Zdbg[(size_t)(n0 + q) * g.E + e] = z[q] + bj;
      }
      if (train) mbar_wait(bar(BAR_DZ_EMPTY + s), ph ^ 1);
      if (drain_pending) {  // (warp-uniform) the previous expert tile's dW is leaving through the dz ring: its staging must have been read
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, 512;" ::: "memory");
        drain_pending = false;
      }
      const uint32_t dzrow = sbase + OFF_DZ + s * DZ_BYTES + (cb >> 1) * CHUNK + jl * 128;
      const int unit0 = 4 * (cb & 1);
      const uint32_t kx = (uint32_t)(unit0 ^ (jl & 7));
      // the pass is dense (out_tc.cu has the derivation): every element as (target 0, weight tnw); t = -log2e*lrelu(z+b) by two FFMA2 + FMNMX,
      // e = 2^t, one reciprocal and one logarithm per 8 elements through the product of the 8 denominators
      auto dense8 = [&](int u, auto masked) {
        const float2 k1 = make_float2(-LOG2E, -LOG2E), k2 = make_float2(-LOG2E * NTF_LRELU_SLOPE, -LOG2E * NTF_LRELU_SLOPE);
        const float2 kb1v = make_float2(kb1, kb1), kb2v = make_float2(kb2, kb2), one2 = make_float2(1.f, 1.f);
        float2 D[4], G[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float2 zz = make_float2(z[u * 8 + 2 * p], z[u * 8 + 2 * p + 1]);
          const float2 t1 = __ffma2_rn(zz, k1, kb1v), t2 = __ffma2_rn(zz, k2, kb2v);
          float2 tt = make_float2(fminf(t1.x, t2.x), fminf(t1.y, t2.y));
          G[p] = make_float2(t1.x < 0.f ? c_pos : c_neg, t1.y < 0.f ? c_pos : c_neg);
          D[p] = __fadd2_rn(make_float2(ex2_approx(tt.x), ex2_approx(tt.y)), one2);
          if (decltype(masked)::value) {
            if (u * 8 + 2 * p >= nrem) { G[p].x = 0.f; tt.x = 0.f; D[p].x = 1.f; }
            if (u * 8 + 2 * p + 1 >= nrem) { G[p].y = 0.f; tt.y = 0.f; D[p].y = 1.f; }
          }
          acc_lin2 = __fadd2_rn(acc_lin2, tt);
        }
        const float2 M01 = __fmul2_rn(D[0], D[1]), M23 = __fmul2_rn(D[2], D[3]);
        const float2 Q = __fmul2_rn(M01, M23);
        const float prod = Q.x * Q.y;
        if (prod < 1.0e30f) {
          acc_lg += lg2_approx(prod);
          const float r = rcp_approx(prod);
          const float2 RQ = __fmul2_rn(make_float2(r, r), make_float2(Q.y, Q.x));
          const float2 R01 = __fmul2_rn(RQ, M23), R23 = __fmul2_rn(RQ, M01);
          G[0] = __fmul2_rn(G[0], __fmul2_rn(R01, D[1])); G[1] = __fmul2_rn(G[1], __fmul2_rn(R01, D[0]));
          G[2] = __fmul2_rn(G[2], __fmul2_rn(R23, D[3])); G[3] = __fmul2_rn(G[3], __fmul2_rn(R23, D[2]));
        } else {
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            acc_lg += lg2_approx(D[p].x) + lg2_approx(D[p].y);
            G[p].x *= rcp_approx(D[p].x); G[p].y *= rcp_approx(D[p].y);
          }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) db_acc2 = __fadd2_rn(db_acc2, G[p]);
        if (train) {
          const __half2 h0 = __floats2half2_rn(G[0].x, G[0].y), h1 = __floats2half2_rn(G[1].x, G[1].y);
          const __half2 h2 = __floats2half2_rn(G[2].x, G[2].y), h3 = __floats2half2_rn(G[3].x, G[3].y);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dzrow + ((kx ^ (uint32_t)u) << 4)), "r"(*reinterpret_cast<const uint32_t*>(&h0)),
                       "r"(*reinterpret_cast<const uint32_t*>(&h1)), "r"(*reinterpret_cast<const uint32_t*>(&h2)), "r"(*reinterpret_cast<const uint32_t*>(&h3)) : "memory");
        }
      };
      if (nrem >= 32) {
#pragma unroll
        for (int u = 0; u < 4; ++u) dense8(u, std::false_type{});
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) dense8(u, std::true_type{});
      }
      if (train) {
        fence_proxy_async();
        mbar_arrive(bar(BAR_DZ_FULL + s));
      }
      if (g.timing && lane == 0 && n < 60) atomicMax((unsigned long long*)(g.timing + TSLOTS * blockIdx.x + 8 + 8 * n + 1), (unsigned long long)clock64());  // slowest warp
      if (!last) continue;
      // ---- end of the expert tile: fold its loss terms, drain dW / db ----
      acc_lin += acc_lin2.x + acc_lin2.y;
      db_acc += db_acc2.x + db_acc2.y;
      loss_total += e_ok ? g.tnw * 0.6931471805599453f * (acc_lg - acc_lin) : 0.f;
      if (train) {
        if (g.timing && threadIdx.x == 0 && n < 60) g.timing[TSLOTS * blockIdx.x + 8 + 8 * n + 2] = clock64();
        mbar_wait(bar(BAR_DW_FULL), j & 1);  // every product of the item has completed: the dz ring is free for staging
        if (g.timing && threadIdx.x == 0 && n < 60) g.timing[TSLOTS * blockIdx.x + 8 + 8 * n + 7] = clock64();
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem + lane_base + TM_DW + cb * 32, v);
        tc_fence_before();
        mbar_arrive(bar(BAR_DW_EMPTY));
        // staging (the 64 KB of the dz ring) as four TMA boxes: chunk cb = columns [32cb, 32cb+32) of the tile, [128 experts][128 B], 128-byte
        // swizzled.  One thread then hands the tile to the TMA unit -- a plain tensor store when this CTA ran the whole expert tile, a
        // reduce-add into zeroed rows when two CTAs share it (two contributions: order-independent) -- and the warps go straight on to the
        // next expert tile; rows past E are clipped by the tensor map.
        uint8_t* srow = sgen + OFF_DZ + cb * CHUNK + jl * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(srow + ((q ^ (jl & 7)) << 4)) = make_float4(v[4 * q] * g.scale, v[4 * q + 1] * g.scale, v[4 * q + 2] * g.scale, v[4 * q + 3] * g.scale);
        float* dbsum = reinterpret_cast<float*>(sgen + OFF_DBS);  // [4][128] db partials of the four team blocks, combined in a fixed order
        dbsum[cb * TE + jl] = db_acc;
        fence_proxy_async();
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (threadIdx.x == 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (whole)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                           ::"l"(reinterpret_cast<uint64_t>(&map_dw)), "r"(c * 32), "r"(e0), "r"(sbase + OFF_DZ + c * CHUNK) : "memory");
            else
              asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                           ::"l"(reinterpret_cast<uint64_t>(&map_dw)), "r"(c * 32), "r"(e0), "r"(sbase + OFF_DZ + c * CHUNK) : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (cb == 0 && e_ok) {
          const float dbv = (((dbsum[jl] + dbsum[TE + jl]) + dbsum[2 * TE + jl]) + dbsum[3 * TE + jl]) * g.scale;
          if (whole) g.db[e] = dbv; else atomicAdd(g.db + e, dbv);
        }
        drain_pending = true;  // the TMA unit is still reading the staging: see the first dz write of the next expert tile
        if (g.timing && threadIdx.x == 0 && n < 60) g.timing[TSLOTS * blockIdx.x + 8 + 8 * n + 3] = clock64();
      }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the last tile has left
    // one loss partial per CTA: fixed-order combine (shuffle tree, then warps in order)
    float* red = reinterpret_cast<float*>(sgen + OFF_BAR + NUM_BARS * 8 + 16);  // [16]
    const float tot = warp_sum(loss_total);
    if (lane == 0) red[warp] = tot;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (threadIdx.x == 0) {
      float l = 0.f;
#pragma unroll
      for (int q = 0; q < EPI_WARPS; ++q) l += red[q];
      g.loss_part[blockIdx.x] = l;
    }
  } else if (warp >= WARP_DA && warp < WARP_DA + 4) {
    // ====== dA epilogue.  The product is taken TRANSPOSED (dA^T = W^T dz: hidden unit on the TMEM lane, teams along the columns), so a
    // warp's 32 lanes hold 32 consecutive hidden units of ONE team: red.global.add.f32 straight from registers, 128 contiguous bytes per
    // instruction -- the access pattern the L2 reduction units take at full rate (scripts/mb/mb_reduce.cu: as fast as the TMA reduce-add,
    // without its shared-memory staging, barriers and proxy fences).  The sum over expert tiles (= across CTAs) happens in L2. ======
    if (train) {
      const int k = threadIdx.x - WARP_DA * 32;  // hidden unit = TMEM lane
      const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
      for (int L = L0, n = 0; L < L1; ++L, ++n) {
        const int t = L % nbt;
        mbar_wait(bar(BAR_DA_FULL), n & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < TB / 32; ++c) {
          float v[32];
          tmem_ld32(tmem + lane_base + TM_DA + c * 32, v);
          if (c == TB / 32 - 1) {
            tc_fence_before();
            mbar_arrive(bar(BAR_DA_EMPTY));
          }
          const int team0 = t * TB + c * 32;
          float* dst = g.dA + (size_t)team0 * HK + k;
          const int nrem = g.B - team0;
          if (nrem >= 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (!(g.exp & 1)) atomicAdd(dst + (size_t)i * HK, v[i] * g.scale);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nrem) atomicAdd(dst + (size_t)i * HK, v[i] * g.scale);
          }
        }
        if (g.timing && k == 0 && n < 60) g.timing[TSLOTS * blockIdx.x + 8 + 8 * n + 6] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (g.timing && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    g.timing[TSLOTS * blockIdx.x + 1] = (long long)gt; g.timing[TSLOTS * blockIdx.x + 4] = clock64();
  }
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
  }
}

// ---- sparse correction pass: the (team, expert) pairs that are NOT (target 0, weight tnw) -----------------------------------------------
// fnn.py:33-43: a team's members carry target 1 and weight tpw, its sampled negatives target 0 (unless they are members) and weight tpw.
// One CTA per team, one warp per candidate (members, then negatives; duplicates and -1 dropped; experts outside this shard skipped), all
// candidates of a team in flight at once (the pass is latency-bound: a handful of dependent global loads per candidate).  A warp computes
// z = a.w + b from the SAME fp16 operand images the tensor-core pass read (fp32 accumulation), then the difference between what the pair
// should contribute and what the dense pass already contributed for it:
//   loss += tpw*bce(x, y) - tnw*softplus(x);  dg = fp16(g_true) - fp16(g_dense)  (the dense pass fed fp16(g_dense) to its products)
//   dW[e,:] += scale*dg*a16[n,:]  (atomic: several teams may hit one expert);  db[e] += scale*(g_true - g_dense);
//   dA[n,:] += scale*dg*w16[e,:]  (summed over the team's candidates in shared memory, fixed order; the CTA owns the row)
// The CTA then finishes its row: dA + corrections, or -- fused ntf_act_bwd of the layer below -- dz_prev = that times lrelu'(act_prev).
// The loss partials of both passes (and, when fused, the column sums of dz_prev) are added up by out_tail_kernel -- they gate nothing but
// the optimiser, so ntf_fnn_step runs that kernel on a side stream.  Runs after out_tc2_kernel.
struct FixArgs {
  const __half* A16;   // [B,128]
  const __half* W16;   // [E,128]
  const float* bias;   // [E]
  const int32_t* m_indptr; const int32_t* m_indices;  // member CSR of the batch (absolute offsets, GLOBAL expert ids, ascending per team)
  const int32_t* neg; int ns;                          // [B,ns] global expert ids, -1 = none (may be NULL: ns = 0)
  int B, E, e_lo;
  float tpw, tnw, scale;
  float* dW; float* db; float* dA;  // NULL on a validation step
  float* loss_part;                   // [n_dense + B]: the dense pass's loss partials, then one per team
  int n_dense;
  const float* act_prev; float* dz_prev;  // optional fusion of ntf_act_bwd for the layer below (ntf_out_train_args.act_prev)
  int exp;                                // debug (NTF_FIX_EXP): 1 = no dW / db atomics, 2 = no loss finalisation, 4 = no candidate work at all
};
constexpr int FIX_WARPS = 8;

__global__ void __launch_bounds__(FIX_WARPS * 32) out_fix_kernel(FixArgs g) {
  __shared__ float sd[FIX_WARPS][HK];
  __shared__ float sloss[FIX_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x;
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  const bool train = g.dW != nullptr;
  // the row this CTA finishes at the end: in flight while the candidates are worked on
  float row0 = 0.f, rslope = 1.f;
  if (train && threadIdx.x < HK) {
    row0 = g.dA[(size_t)n * HK + threadIdx.x];
    if (g.dz_prev) rslope = g.act_prev[(size_t)n * HK + threadIdx.x] > 0.f ? 1.f : NTF_LRELU_SLOPE;
  }
  const int mb = g.m_indptr[n], me = g.m_indptr[n + 1], nm = me - mb;
  // this lane's slice of the team's activations: hidden units lane, lane+32, lane+64, lane+96 (coalesced 64-byte loads / 128-byte adds)
  float a[4], dacc[4] = {0.f, 0.f, 0.f, 0.f}, loss = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) a[q] = __half2float(g.A16[(size_t)n * HK + q * 32 + lane]);
  const int ncand = nm + g.ns;
  int kth = 0;  // running index of the valid candidates: warp w takes those with kth % FIX_WARPS == w
  for (int c0 = 0; c0 < ncand; c0 += 32) {
    // candidate c0 + lane (the same in every warp): a member (ascending, unique) or a negative; a negative is dropped if it is -1, a member,
    // or a repeat of an earlier negative
    const int c = c0 + lane;
    int e = -1;
    bool member = false;
    if (c < nm) { e = g.m_indices[mb + c]; member = true; }
    else if (c < ncand) e = g.neg[(size_t)n * g.ns + (c - nm)];
    const unsigned same = __match_any_sync(0xffffffffu, e);  // duplicates inside this group: the lowest lane keeps the value (members sit lowest)
    if (e >= 0 && (same & ((1u << lane) - 1u)) != 0u) e = -1;
    if (c0 > 0 && e >= 0) {  // (more than 32 candidates: rare) also against the earlier groups, from memory
      for (int p = mb; p < me && p - mb < c0; ++p) if (g.m_indices[p] == e) { e = -1; break; }
      for (int p = max(0, c0 - nm) - 1; p >= 0 && e >= 0; --p) if (g.neg[(size_t)n * g.ns + p] == e) e = -1;
    }
    const int el = e - g.e_lo;  // column of this shard
    unsigned todo = __ballot_sync(0xffffffffu, e >= 0 && el >= 0 && el < g.E);
    if (g.exp & 4) todo = 0u;
    while (todo) {
      const int L = __ffs(todo) - 1;
      todo &= todo - 1;
      if ((kth++ % FIX_WARPS) != warp) continue;
      const int j = __shfl_sync(0xffffffffu, el, L);
      const bool y = __shfl_sync(0xffffffffu, (int)member, L) != 0;
      float w[4], dot = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) { w[q] = __half2float(g.W16[(size_t)j * HK + q * 32 + lane]); dot = fmaf(a[q], w[q], dot); }
      const float bj = __ldg(g.bias + j);
      dot = warp_sum(dot);
      const float zz = dot + bj;
      const float x = fmaxf(zz, NTF_LRELU_SLOPE * zz);
      const float slope = zz > 0.f ? 1.f : NTF_LRELU_SLOPE;
      // what the dense pass computed for this pair (its formulas: t = -log2e*x, e = 2^t, sigmoid = 1/(1+e), softplus = (lg2(1+e) - t)*ln2)
      const float tt = -LOG2E * x;
      const float dens = 1.f + ex2_approx(tt);
      const float g_dense = g.tnw * slope * rcp_approx(dens);
      const float l_dense = g.tnw * LN2 * (lg2_approx(dens) - tt);
      // what it should contribute (stable form)
      const float ex = ex2_approx(fabsf(x) * -LOG2E), den = 1.f + ex, r = rcp_approx(den);
      const float sig = x > 0.f ? r : ex * r, yf = y ? 1.f : 0.f;
      const float g_true = g.tpw * (sig - yf) * slope;
      const float l_true = g.tpw * ((1.f - yf) * x + fmaxf(-x, 0.f) + LN2 * lg2_approx(den));
      loss += l_true - l_dense;  // (the same value in every lane; lane 0's copy is used)
      if (train) {
        const float dg = (__half2float(__float2half_rn(g_true)) - __half2float(__float2half_rn(g_dense))) * g.scale;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (!(g.exp & 1)) atomicAdd(g.dW + (size_t)j * HK + q * 32 + lane, dg * a[q]);
          dacc[q] = fmaf(dg, w[q], dacc[q]);
        }
        if (lane == 0 && !(g.exp & 1)) atomicAdd(g.db + j, (g_true - g_dense) * g.scale);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) sd[warp][q * 32 + lane] = dacc[q];
  if (lane == 0) sloss[warp] = loss;
  __syncthreads();
  if (train && threadIdx.x < HK) {
    float d = row0;
#pragma unroll
    for (int q = 0; q < FIX_WARPS; ++q) d += sd[q][threadIdx.x];
    if (g.dz_prev) g.dz_prev[(size_t)n * HK + threadIdx.x] = d * rslope;  // the layer below's activation derivative (fnn.py:25 lrelu), fused
    else g.dA[(size_t)n * HK + threadIdx.x] = d;
  }
  if (threadIdx.x == 0) {
    float l = 0.f;
#pragma unroll
    for (int q = 0; q < FIX_WARPS; ++q) l += sloss[q];
    g.loss_part[g.n_dense + n] = l;
  }
}

// final reductions of a call, fixed order (deterministic): loss_out = scale * (sum of the loss partials of both passes) and, after the fused
// activation backward, db_prev = colsum(dz_prev): 8 teams per CTA -> per-CTA column partials -> the last CTA to finish adds them in CTA
// order (thread = 4 columns x one of 8 interleaved slices of the CTAs: 16-byte loads, four independent chains, then the slices in order)
constexpr int TAIL_ROWS = 8;
__global__ void __launch_bounds__(256) out_tail_kernel(const float* __restrict__ dz, int B, float* __restrict__ col_part, float* __restrict__ db_prev,
                                                       const float* __restrict__ loss_part, int n_loss, float scale, float* __restrict__ loss_out,
                                                       unsigned* __restrict__ done) {
  __shared__ float cs[8][HK];
  __shared__ double dred[256];
  __shared__ int last_s;
  if (dz) {
    const int c = threadIdx.x & (HK - 1), half = threadIdx.x >> 7;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < TAIL_ROWS / 2; ++i) {
      const int r = blockIdx.x * TAIL_ROWS + half * (TAIL_ROWS / 2) + i;
      if (r < B) s += dz[(size_t)r * HK + c];
    }
    cs[half][c] = s;
    __syncthreads();
    if (threadIdx.x < HK) col_part[(size_t)blockIdx.x * HK + threadIdx.x] = cs[0][threadIdx.x] + cs[1][threadIdx.x];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last_s = (atomicAdd(done, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!last_s) return;
    __threadfence();
    const int cg = threadIdx.x & 31, sl8 = threadIdx.x >> 5, nb = (int)gridDim.x;
    const float4* cp = reinterpret_cast<const float4*>(col_part);
    float4 acc[4] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    for (int b0 = sl8; b0 < nb; b0 += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int b = b0 + 8 * u;
        if (b < nb) {
          const float4 v = __ldcg(cp + (size_t)b * (HK / 4) + cg);
          acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
        }
      }
    }
    const float4 t = make_float4((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y),
                                 (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z), (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w));
    __syncthreads();
    reinterpret_cast<float4*>(&cs[sl8][0])[cg] = t;
    __syncthreads();
    if (threadIdx.x < HK) {
      float c = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) c += cs[q][threadIdx.x];
      db_prev[threadIdx.x] = c;
    }
    if (threadIdx.x == 0) *done = 0u;
  }
  // (one CTA gets here: the last one, or the only one when there are no column sums to take)
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_loss; i += 256) acc += (double)__ldcg(loss_part + i);
  dred[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) dred[threadIdx.x] += dred[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = (float)(dred[0] * (double)scale);
}

__global__ void to_half_kernel2(const float* __restrict__ x, size_t n, __half* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = __float2half_rn(x[i]);
}
}  // namespace

// workspace: fp16 A | loss partials | fp16 W image (when the caller keeps none)
// workspace: fp16 A | loss partials (dense pass: one per CTA; correction pass: one per block) + block counter | column-sum partials | fp16 W image
static size_t tc2_loss_slots(int B) { return (size_t)1024 + (size_t)B + 8; }  // dense pass: one per CTA; correction pass: one per team; 2 counters
static size_t tc2_off_loss(int B, int h) { return align_up((size_t)B * h * sizeof(__half), 256); }
static size_t tc2_off_col(int B, int h) { return tc2_off_loss(B, h) + align_up(tc2_loss_slots(B) * sizeof(float), 1024); }
static size_t tc2_off_w16(int B, int h) { return tc2_off_col(B, h) + align_up((size_t)cdiv(B, TAIL_ROWS) * HK * sizeof(float), 1024); }
size_t ntf_out_train_tc2_workspace_bytes(int B, int h, int E) { return tc2_off_w16(B, h) + align_up((size_t)E * h * sizeof(__half), 1024); }

static void tc2_plan(const ntf_ctx* ctx, int B, int E, Tc2Args* g, int* grid) {
  g->nct = cdiv(E, TE); g->nbt = cdiv(B, TB);
  const int sm = ctx->sm_count > 0 ? ctx->sm_count : 148;
  if (g->nct >= sm) { g->split = 0; *grid = sm; return; }
  int s = sm / g->nct;
  if (s > g->nbt) s = g->nbt;
  if (s < 1) s = 1;
  g->split = s; *grid = g->nct * s;
}

extern "C" int ntf_out_train_prepare(ntf_ctx* ctx, void* stream, int B, int h, int E, float* dW, float* db, float* dA) {
  NTF_REQUIRE(ctx && dW && db && dA && h == HK && B > 0 && E > 0, NTF_ERR_BAD_ARG, "out_train_prepare: bad argument");
  cudaStream_t st = as_stream(stream);
  NTF_CUDA(cudaMemsetAsync(dA, 0, (size_t)B * h * sizeof(float), st));
  Tc2Args g{};
  g.B = B; g.E = E; g.dW = dW; g.db = db;
  int grid;
  tc2_plan(ctx, B, E, &g, &grid);
  if (g.split != 1) { NTF_COUNT_LAUNCH; zero_cut_tiles_kernel<<<grid, 256, 0, st>>>(g); NTF_LAUNCH_CHECK(); }
  return NTF_OK;
}


int ntf_out_train_tc2(ntf_ctx* ctx, cudaStream_t st, const ntf_out_train_args* a, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(a->h == HK, NTF_ERR_UNSUPPORTED, "out_train(tf32): hidden width %d (kernel is built for %d)", a->h, HK);
  NTF_REQUIRE(!a->W_delta, NTF_ERR_UNSUPPORTED, "out_train(tf32, persistent): the Flipout layer runs on out_tc.cu");
  NTF_REQUIRE(workspace_bytes >= ntf_out_train_tc2_workspace_bytes(a->B, a->h, a->E), NTF_ERR_WORKSPACE, "out_train(tf32): workspace too small");
  NTF_REQUIRE(((uintptr_t)workspace & 255) == 0, NTF_ERR_BAD_ARG, "out_train(tf32): the workspace must be 256-byte aligned");
  NTF_REQUIRE((((uintptr_t)a->A | (uintptr_t)a->W) & 15) == 0, NTF_ERR_BAD_ARG, "out_train(tf32): A and W must be 16-byte aligned");
  const bool train = a->dW != nullptr;
  NTF_REQUIRE(!train || (a->db && a->dA), NTF_ERR_BAD_ARG, "out_train(tf32): training needs dW, db and dA");
  NTF_REQUIRE(!train || (((uintptr_t)a->dA | (uintptr_t)a->dW) & 15) == 0, NTF_ERR_BAD_ARG, "out_train(tf32): dA and dW must be 16-byte aligned");
  NTF_REQUIRE(!a->special && !a->special_t, NTF_ERR_BAD_ARG, "out_train(tf32, persistent): takes the member CSR and `neg` (the sparse correction pass), not bit planes");
  NTF_REQUIRE(a->ns == 0 || a->neg, NTF_ERR_BAD_ARG, "out_train(tf32): ns=%d without neg", a->ns);
  NTF_REQUIRE(a->m_indptr && a->m_indices, NTF_ERR_BAD_ARG, "out_train(tf32): the member CSR is missing");
  char* ws = (char*)workspace;
  __half* A16 = a->A16 ? (__half*)const_cast<void*>(a->A16) : (__half*)ws;
  float* loss_part = (float*)(ws + tc2_off_loss(a->B, a->h));
  unsigned* done = (unsigned*)(loss_part + tc2_loss_slots(a->B) - 2);  // (zeroed once per context below; the kernel leaves it zero)
  const __half* W16 = (const __half*)a->W16;
  NTF_REQUIRE((a->act_prev != nullptr) == (a->dz_prev != nullptr) && (a->act_prev != nullptr) == (a->db_prev != nullptr), NTF_ERR_BAD_ARG,
              "out_train(tf32): act_prev, dz_prev and db_prev come together");
  NTF_REQUIRE(!a->act_prev || train, NTF_ERR_BAD_ARG, "out_train(tf32): the fused activation backward needs a training step");
  const int blocks_h = ctx->sm_count * 8;
  if (!W16) {  // no image kept by the caller: made here (one pass over W: 6 bytes per weight)
    __half* w = (__half*)(ws + tc2_off_w16(a->B, a->h));
    const size_t nw = (size_t)a->E * HK;
    NTF_COUNT_LAUNCH; to_half_kernel2<<<(unsigned)(cdiv((int)((nw + 255) / 256), 1) < blocks_h ? (nw + 255) / 256 : blocks_h), 256, 0, st>>>(a->W, nw, w);
    NTF_LAUNCH_CHECK();
    W16 = w;
  }
  NTF_REQUIRE((((uintptr_t)A16 | (uintptr_t)W16) & 15) == 0, NTF_ERR_BAD_ARG, "out_train(tf32): fp16 operands must be 16-byte aligned");
  if (!a->A16) {
    const size_t na = (size_t)a->B * HK;
    NTF_COUNT_LAUNCH; to_half_kernel2<<<(unsigned)((na + 255) / 256 < (size_t)blocks_h ? (na + 255) / 256 : blocks_h), 256, 0, st>>>(a->A, na, A16);
    NTF_LAUNCH_CHECK();
  }
  CUtensorMap mw, mh, mdw;
  int rc;
  if ((rc = make_map(ctx, &mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, W16, (uint64_t)a->E, HK, TE, 64))) return rc;
  if ((rc = make_map(ctx, &mh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, A16, (uint64_t)a->B, HK, TB, 64))) return rc;
  if ((rc = make_map(ctx, &mdw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, train ? (const void*)a->dW : (const void*)a->W, (uint64_t)a->E, HK, TE, 32))) return rc;
  if (train && !a->prepared) NTF_CUDA(cudaMemsetAsync(a->dA, 0, (size_t)a->B * a->h * sizeof(float), st));
  Tc2Args g{};
  g.bias = a->b;
  g.B = a->B; g.E = a->E; g.tpw = a->tpw; g.tnw = a->tnw; g.scale = a->loss_scale;
  g.dW = a->dW; g.db = a->db; g.dA = a->dA; g.loss_part = loss_part; g.done = done;
  const char* dbg = getenv("NTF_TC_ZDBG");
  g.Zdbg = dbg ? (float*)(uintptr_t)strtoull(dbg, nullptr, 0) : nullptr;
  const char* tim = getenv("NTF_TC_TIMING");
  g.timing = tim ? (long long*)(uintptr_t)strtoull(tim, nullptr, 0) : nullptr;
  { static const int ex = getenv("NTF_TC2_EXP") ? atoi(getenv("NTF_TC2_EXP")) : 0; g.exp = ex; }
  int grid;
  tc2_plan(ctx, a->B, a->E, &g, &grid);
  NTF_REQUIRE(grid <= 1024, NTF_ERR_UNSUPPORTED, "out_train(tf32): %d CTAs", grid);
  if (train && !a->prepared && g.split != 1) { NTF_COUNT_LAUNCH; zero_cut_tiles_kernel<<<grid, 256, 0, st>>>(g); NTF_LAUNCH_CHECK(); }
  NTF_CUDA(cudaFuncSetAttribute(out_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  NTF_COUNT_LAUNCH; out_tc2_kernel<<<grid, NT, SMEM_BYTES, st>>>(mw, mh, mdw, g);
  NTF_LAUNCH_CHECK();
  FixArgs f{};
  f.A16 = A16; f.W16 = W16; f.bias = a->b; f.m_indptr = a->m_indptr; f.m_indices = a->m_indices; f.neg = a->neg; f.ns = a->neg ? a->ns : 0;
  f.B = a->B; f.E = a->E; f.e_lo = a->e_lo; f.tpw = a->tpw; f.tnw = a->tnw; f.scale = a->loss_scale;
  f.dW = a->dW; f.db = a->db; f.dA = a->dA; f.loss_part = loss_part; f.n_dense = grid;
  f.act_prev = a->act_prev; f.dz_prev = a->dz_prev;
  { const char* fx = getenv("NTF_FIX_EXP"); f.exp = fx ? atoi(fx) : 0; }
  NTF_COUNT_LAUNCH; out_fix_kernel<<<a->B, FIX_WARPS * 32, 0, st>>>(f);
  NTF_LAUNCH_CHECK();
  if (a->defer_finish) return NTF_OK;
  return ntf_out_train_finish(ctx, (void*)st, a, workspace, workspace_bytes);
}

// the call's final reductions (out_tail_kernel); see ntf_out_train_args.defer_finish
extern "C" int ntf_out_train_finish(ntf_ctx* ctx, void* stream, const ntf_out_train_args* a, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ctx && a && workspace && workspace_bytes >= ntf_out_train_tc2_workspace_bytes(a->B, a->h, a->E), NTF_ERR_BAD_ARG, "out_train_finish: bad argument");
  char* ws = (char*)workspace;
  float* loss_part = (float*)(ws + tc2_off_loss(a->B, a->h));
  unsigned* done = (unsigned*)(loss_part + tc2_loss_slots(a->B) - 2);
  Tc2Args g{};
  int grid;
  tc2_plan(ctx, a->B, a->E, &g, &grid);
  const bool fused = a->dz_prev != nullptr;
  NTF_COUNT_LAUNCH;
  out_tail_kernel<<<fused ? cdiv(a->B, TAIL_ROWS) : 1, 256, 0, as_stream(stream)>>>(fused ? a->dz_prev : nullptr, a->B, (float*)(ws + tc2_off_col(a->B, a->h)), a->db_prev,
                                                                                   loss_part, grid + a->B, a->loss_scale, a->loss_out, done);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
