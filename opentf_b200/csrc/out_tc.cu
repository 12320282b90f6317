// out_tc.cu -- NTF_TF32 mode: the output layer on tcgen05 tensor cores with TMEM accumulators and TMA-fed operands.
//
// One fused kernel per training step of the hidden->all-experts layer (fnn.py:25 last layer, 32-46, 135 and its
// autograd): logits, weighted BCE, dz, dW, db and dA are produced without the [B,E] logits or gradients ever
// touching HBM.  Structure = FlashAttention backward with Q<->A (hidden activations), K<->W, dS<->dz:
//
//   CTA e owns experts [128e, 128e+128): an fp16 image of the W tile resident in shared memory, the dW accumulator
//   resident in TMEM for the whole kernel, db in registers.  It streams over the batch in tiles of 128 teams:
//     TMA warp       : W tile (fp32, once) and the fp16 activation tiles (2-stage ring)        [cp.async.bulk.tensor]
//     MMA thread     : Z^T[128e x 128n] = W . A^T                           kind::f16, fp32 accumulate in TMEM (2 stages)
//     16 epilogue warps (thread = expert x 32 of the tile's teams): TMEM -> regs, +b, lrelu, weighted BCE, dz; loss and db
//                      accumulate in registers; dz (scaled by 1/loss_scale, fp16) -> shared memory
//     special warp   : bulk-copies the tile's slice of the tile-transposed `condition` / member bit planes (ntf_special_tiles)
//                      into shared memory, one tile ahead of the epilogue
//     MMA thread     : dW[128e x 128k] += dz^T . A        accumulates across all batch tiles in TMEM
//                      dA[128n x 128k]   = dz . W
//     4 dA warps     : TMEM -> regs -> shared memory -> TMA reduce-add (cp.reduce.async.bulk.tensor) into dA[B,128]: the sum over
//                      expert tiles (= across CTAs) is done by L2, with no SM-side atomics
//   At the end the epilogue warps drain dW from TMEM, write db, and one loss partial per CTA.
//
// Operand precision: all three products read fp16 operands (RN from fp32: 10-bit mantissa, the same as TF32) and accumulate
// in fp32 -- this is what the ABI calls NTF_TF32: TF32-class tolerance (2^-10 relative per product term), stated in
// tests/test_gpu_tc.py.  16-bit operands are what lets ONE 128B-swizzled image serve as K-major operand of the forward
// product and as MN-major operand of the backward products (32-bit types need a different swizzle atom for MN-major), halve
// shared-memory traffic (the binding resource next to the MUFU pipe: every SS-mode MMA re-reads both operand tiles) and run
// at twice the tf32 MMA rate.  dz is scaled into fp16's normal range.  Values beyond fp16 range (|x| > 65504) would saturate:
// hidden activations and weights of this model family are O(1).
//
// Roofline: tensor pipe by the contract's accounting (algorithmic flops per team = 6*h*E); the resources that
// actually bind are the MUFU pipe (2 transcendentals per logit: ex2, rcp), shared-memory bandwidth and issue slots.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

#include "tc_common.cuh"

int ntf_loss_reduce_impl(cudaStream_t st, const float* part, int n, float scale, float* loss_out);

namespace {

constexpr int TE = 128;  // experts per CTA tile (UMMA M of the forward / dW products)
constexpr int TB = 128;  // teams per batch tile (UMMA N of the forward / dA products)
constexpr int HK = 128;  // hidden width this kernel is built for

constexpr uint32_t CHUNK = 128 * 128;            // one swizzle chunk: 128 rows x 128 bytes (64 halfs)
constexpr uint32_t W16_BYTES = TE * HK * 2;      // 32 KB : 2 chunks (hidden 0-63 | 64-127) of [128 experts][128 B]
constexpr uint32_t A16_BYTES = TB * HK * 2;      // 32 KB : 2 chunks (hidden 0-63 | 64-127) of [128 teams  ][128 B]
constexpr uint32_t DZ_BYTES = TE * TB * 2;       // 32 KB : 2 chunks (teams  0-63 | 64-127) of [128 experts][128 B]
constexpr uint32_t W32_BYTES = TE * HK * 4;      // 64 KB : 4 chunks (32 hidden each) of [128 experts][128 B], staged in the dz ring
constexpr uint32_t OFF_W16 = 0;
constexpr int A_STAGES = 3;                                 // tile t+2's activations load while tile t's backward products still read theirs
constexpr uint32_t OFF_A16 = OFF_W16 + W16_BYTES;
constexpr uint32_t OFF_Q = OFF_A16 + 2 * A16_BYTES;          // Flipout: the third activation slot holds the perturbation-term tile instead
constexpr uint32_t Q_BYTES = TE * TB * 2;
constexpr uint32_t OFF_DZ = OFF_A16 + A_STAGES * A16_BYTES; // 2 stages (first use: landing zone of the fp32 W tile)
constexpr uint32_t OFF_PLANE = OFF_DZ + 2 * DZ_BYTES;       // 2 stages x (special | member) bit planes of a tile, [128 experts][4 words] each
constexpr uint32_t PLANE_BYTES = TE * 4 * 4;                // 2 KB each
constexpr uint32_t OFF_DAST = OFF_PLANE + 4 * PLANE_BYTES;  // dA staging: one [128 teams][128 B] fp32 chunk on its way to the TMA reduce-add
constexpr uint32_t OFF_BAR = OFF_DAST + CHUNK;              // mbarriers + tmem base + small reduction scratch
constexpr uint32_t SMEM_BYTES = OFF_BAR + 512;              // the dynamic window is declared 1024-byte aligned (checked at run time)
static_assert(W32_BYTES <= 2 * DZ_BYTES, "the fp32 W tile is staged in the dz ring");
static_assert(SMEM_BYTES <= 232448, "over the 227 KB shared memory limit");

// TMEM columns (fp32 accumulators, 128 lanes each)
constexpr uint32_t TM_Z = 0;      // 2 x 128
constexpr uint32_t TM_DW = 256;   // 128
constexpr uint32_t TM_DA = 384;   // 128
constexpr uint32_t TM_COLS = 512;

enum { BAR_W32 = 0, BAR_W16 = 1, BAR_A_FULL = 2, BAR_A_EMPTY = 5, BAR_Z_FULL = 8, BAR_Z_EMPTY = 10, BAR_DZ_FULL = 12, BAR_DZ_EMPTY = 14,
       BAR_DA_FULL = 16, BAR_DA_EMPTY = 17, BAR_DW_FULL = 18, BAR_SP_FULL = 19, BAR_SP_EMPTY = 21, BAR_Q_FULL = 23, BAR_Q_EMPTY = 24, NUM_BARS = 25 };
// A_*  : fp16 activation tile ring, 3 stages (operand of the forward and of the dW product; released after the tile's last product)
// SP_* : special / member bit planes of a tile (2-stage ring, filled by 1-D bulk copies from the tile-transposed planes in HBM)
// Q_*  : Flipout only: the tile of the perturbation term (one stage, living in the third activation slot; the activation ring has 2 stages then)

struct TcArgs {
  const float* bias;          // [E]
  uint32_t* special_t;        // tile-transposed planes (ntf_special_tiles), or NULL: every weight tnw, every target 0.
  uint32_t* member_t;         // The kernel zeroes the words it consumed: the planes are clean again when it returns.
  int Epad;                   // E rounded up to 128: row count of one tile slab of the planes
  int B, E;
  float tpw, tnw, scale;      // scale = loss_scale (1/B_global)
  float* dW;                  // [E,128] or NULL (validation step: forward + loss only)
  float* db;                  // [E]
  float* dA;                  // [B,128], zeroed by the caller, accumulated with red.add
  float* loss_part;           // [gridDim.x]
  float* P;                   // inference: [B,E] probabilities
  float* Zdbg;                // debug: raw logits z [B,E] (NULL in production)
  long long* timing;          // debug: clock64 stamps of CTA 0, [tile][8] (NULL in production)
  int exp;                    // debug: experiment bits (NTF_TC_EXP): 1 = skip the dA reduction, 2 = skip the special-bit path
  // work split: CTAs [0, n_full) own a whole expert tile (all batch tiles); the remaining expert tiles -- the partial last wave --
  // are cut `split` ways along the batch so that the last wave is short: those CTAs ADD their dW / db partials (pre-zeroed rows)
  int n_full, split;
  // Flipout (Bnn): z += s_out * (A_s W_delta^T + b_delta).  The perturbation term is a product of its own (MODE 2) whose
  // signed result crosses HBM once as fp16 tiles; its backward products (MODE 3) read s_out*dz tiles written by the MODE 0 epilogue.
  const uint32_t* sign_t;     // s_out as a tile-transposed plane (bit = 1: -1), word layout of special_t; read-only
  __half* Q;                  // [tiles of 128 teams][Epad/128][16 units][128 experts][8 halfs]: s_out*(Z2 + b_delta) (MODE 2 writes, MODE 0/1 add)
  __half* DZS;                // [tiles of 128 teams * Epad][128 teams] row-major: s_out*dz/loss_scale (MODE 0 writes, MODE 3 loads by TMA)
  float* db_delta;            // [E]: colsum(s_out*dz)
};

constexpr int EPI_WARPS = 16;
constexpr int NT = 736;  // warps 0-15: logits/loss epilogue, 16-19: dA epilogue, 20: TMA, 21: MMA issuer, 22: special planes (23 warps: 88 registers per thread)
constexpr int WARP_DA = 16, WARP_TMA = 20, WARP_MMA = 21, WARP_SP = 22;
constexpr int EPI_THREADS = EPI_WARPS * 32;

constexpr uint32_t IDESC_FWD = instr_desc(0, 0, 0, TE, TB);   // Z^T  = W16(K-major) . A16(K-major)^T                     K = hidden
constexpr uint32_t IDESC_DW = instr_desc(0, 0, 1, TE, HK);    // dW  += dz^T(K-major: K = teams) . A16(MN-major: N = hidden)
constexpr uint32_t IDESC_DA = instr_desc(0, 1, 1, TB, HK);    // dA   = dz(MN-major: M = teams) . W16(MN-major: N = hidden), K = experts

// MODE 0: training / validation step.  MODE 1: inference scores P = sigmoid(lrelu(z)).
// Flipout (FLIP): MODE 2: forward product of the perturbation path only, Q = s_out*(A_s W_delta^T + b_delta) -> HBM (fp16 tiles);
//                 MODE 0/1 with FLIP add the Q tile to the logits; MODE 0 also writes s_out*dz tiles (DZS) and db_delta;
//                 MODE 3: backward products of the perturbation path only: dW_delta += DZS^T A_s, dA_s = DZS W_delta (dz tiles come by TMA).
template <int MODE, bool FLIP>
__global__ void __launch_bounds__(NT, 1) out_tc_kernel(const __grid_constant__ CUtensorMap map_w32, const __grid_constant__ CUtensorMap map_a16,
                                                       const __grid_constant__ CUtensorMap map_da, const __grid_constant__ CUtensorMap map_dzs, TcArgs g) {
  constexpr int AST = (FLIP && MODE < 2) ? 2 : A_STAGES;  // activation ring stages in use (the Q tile takes the third slot)
  constexpr bool QIN = FLIP && MODE < 2;                  // this instance adds the perturbation-term tile to its logits
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if ((sbase & 1023u) != 0u) __trap();  // the 128-byte swizzle atoms need a 1024-byte aligned base
  uint8_t* sgen = smem_raw;
  const uint32_t bars = sbase + OFF_BAR;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + OFF_BAR + NUM_BARS * 8);
  auto bar = [&](int i) { return bars + 8u * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (g.timing && blockIdx.x == 0 && threadIdx.x == 0) g.timing[15 * 8 + 0] = clock64();  // kernel entry
  if (g.timing && threadIdx.x == 0) {  // every CTA: start / end on the global timer (ns) and its SM
    unsigned long long gt; unsigned smid;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    g.timing[128 + 3 * blockIdx.x] = (long long)gt; g.timing[128 + 3 * blockIdx.x + 2] = smid;
  }
  const int nt_all = (g.B + TB - 1) / TB;
  int etile = blockIdx.x, t_begin = 0, t_end = nt_all;
  const bool shared_tile = (int)blockIdx.x >= g.n_full;  // this expert tile's batch range is cut across g.split CTAs
  if (shared_tile) {
    const int k = blockIdx.x - g.n_full, part = k % g.split;
    etile = g.n_full + k / g.split;
    t_begin = part * nt_all / g.split;
    t_end = (part + 1) * nt_all / g.split;
  }
  const int e0 = etile * TE;
  const int ntiles = t_end - t_begin;  // batch tiles of this CTA: local index it, global tile t_begin + it
  const bool train = (MODE == 0 && g.dW != nullptr) || MODE == 3;  // backward products run

  if (threadIdx.x == 0) {
    mbar_init(bar(BAR_W32), 1);
    mbar_init(bar(BAR_W16), EPI_THREADS);
    mbar_init(bar(BAR_DW_FULL), 1);
    mbar_init(bar(BAR_DA_FULL), 1);
    mbar_init(bar(BAR_DA_EMPTY), 128);
    mbar_init(bar(BAR_Q_FULL), 1);
    mbar_init(bar(BAR_Q_EMPTY), EPI_THREADS);
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(bar(BAR_A_FULL + s), 1);
      mbar_init(bar(BAR_A_EMPTY + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(BAR_Z_FULL + s), 1);
      mbar_init(bar(BAR_Z_EMPTY + s), EPI_THREADS);
      mbar_init(bar(BAR_DZ_FULL + s), MODE == 3 ? 1 : EPI_THREADS);  // MODE 3: the dz tiles arrive by TMA
      mbar_init(bar(BAR_DZ_EMPTY + s), 1);
      mbar_init(bar(BAR_SP_FULL + s), 1);
      mbar_init(bar(BAR_SP_EMPTY + s), EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WARP_MMA) {  // TMEM allocation is owned by the MMA warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + NUM_BARS * 8), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (g.timing && blockIdx.x == 0 && threadIdx.x == 0) g.timing[15 * 8 + 1] = clock64();  // barriers + TMEM ready

  if (warp == WARP_TMA) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      mbar_expect_tx(bar(BAR_W32), W32_BYTES);
      for (int c = 0; c < 4; ++c) tma_load_2d(sbase + OFF_DZ + c * CHUNK, &map_w32, c * 32, e0, bar(BAR_W32));
      for (int it = 0; it < ntiles; ++it) {
        const int s = it % AST;
        const uint32_t ph = (it / AST) & 1;
        mbar_wait(bar(BAR_A_EMPTY + s), ph ^ 1);
        mbar_expect_tx(bar(BAR_A_FULL + s), A16_BYTES);
        for (int c = 0; c < 2; ++c)
          tma_load_2d(sbase + OFF_A16 + s * A16_BYTES + c * CHUNK, &map_a16, c * 64, (t_begin + it) * TB, bar(BAR_A_FULL + s));
        if (MODE == 3) {  // the s_out*dz tile of (batch tile, expert tile): rows = experts, 2 chunks of 64 teams, same swizzled image the epilogue writes
          const int s2 = it & 1;
          if (it == 0) mbar_wait(bar(BAR_W16), 0);  // the dz ring was the landing zone of the fp32 W tile
          mbar_wait(bar(BAR_DZ_EMPTY + s2), ((it >> 1) & 1) ^ 1);
          mbar_expect_tx(bar(BAR_DZ_FULL + s2), DZ_BYTES);
          for (int c = 0; c < 2; ++c)
            tma_load_2d(sbase + OFF_DZ + s2 * DZ_BYTES + c * CHUNK, &map_dzs, c * 64, (t_begin + it) * g.Epad + e0, bar(BAR_DZ_FULL + s2));
        }
      }
    }
  } else if (warp == WARP_SP) {
    // ====== special / member planes of the tile: two 2 KB bulk copies per tile from the tile-transposed planes (ntf_special_tiles) ======
    const bool planes = MODE == 0 && g.special_t && !(g.exp & 2);
    if (lane == 0 && (planes || QIN)) {
      for (int it = 0; it < ntiles; ++it) {
        if (planes) {
          const int s = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          mbar_wait(bar(BAR_SP_EMPTY + s), ph ^ 1);
          mbar_expect_tx(bar(BAR_SP_FULL + s), 2 * PLANE_BYTES);
          const size_t off = ((size_t)(t_begin + it) * g.Epad + e0) * 4;
          bulk_load(sbase + OFF_PLANE + s * 2 * PLANE_BYTES, g.special_t + off, PLANE_BYTES, bar(BAR_SP_FULL + s));
          bulk_load(sbase + OFF_PLANE + s * 2 * PLANE_BYTES + PLANE_BYTES, g.member_t + off, PLANE_BYTES, bar(BAR_SP_FULL + s));
        }
        if (QIN) {  // the 32 KB perturbation-term tile (contiguous in HBM), one tile ahead of the epilogue
          mbar_wait(bar(BAR_Q_EMPTY), (it & 1) ^ 1);
          mbar_expect_tx(bar(BAR_Q_FULL), Q_BYTES);
          bulk_load(sbase + OFF_Q, g.Q + ((size_t)(t_begin + it) * g.Epad + e0) * TB, Q_BYTES, bar(BAR_Q_FULL));
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // =========================================== MMA issuer ===========================================
    // the whole warp runs this loop converged; one elected lane issues the MMAs and the commits (tc_common.cuh: elect_one)
    const uint64_t d_w16_k = smem_desc(sbase + OFF_W16, 16, 1024);      // W16 as K-major operand (forward)
    const uint64_t d_w16_mn = smem_desc(sbase + OFF_W16, CHUNK, 1024);  // W16 as MN-major operand (dA^T: M = hidden)
    auto issue_fwd = [&](int it) {
      const int sa = it % AST, s = it & 1;
      const uint32_t pha = (it / AST) & 1, ph = (it >> 1) & 1;
      mbar_wait(bar(BAR_A_FULL + sa), pha);
      if (g.timing && blockIdx.x == 0 && lane == 0) g.timing[it * 8 + 1] = clock64();
      mbar_wait(bar(BAR_Z_EMPTY + s), ph ^ 1);
      if (g.timing && blockIdx.x == 0 && lane == 0) g.timing[it * 8 + 2] = clock64();
      tc_fence_after();
      const uint64_t d_a = smem_desc(sbase + OFF_A16 + sa * A16_BYTES, 16, 1024);
#pragma unroll
      for (int i = 0; i < HK / 16; ++i) {  // 8 k-steps of 16 halfs (32 bytes): chunk i/4, 32-byte slice i%4 of the 128-byte swizzle row
        const uint64_t koff = (uint64_t)(((i >> 2) * CHUNK + (i & 3) * 32) >> 4);
        if (elect_one()) mma_f16(tmem + TM_Z + s * TB, d_w16_k + koff, d_a + koff, IDESC_FWD, i > 0);
      }
      if (elect_one()) {
        tc_commit(bar(BAR_Z_FULL + s));
        if (!train) tc_commit(bar(BAR_A_EMPTY + sa));  // forward-only: the stage is free once the forward product has read it
      }
    };
    mbar_wait(bar(BAR_W16), 0);
    if (g.timing && blockIdx.x == 0 && lane == 0) g.timing[15 * 8 + 2] = clock64();  // W image ready
    if (MODE != 3 && ntiles > 0) issue_fwd(0);
    for (int it = 0; it < ntiles; ++it) {
      if (MODE != 3 && it + 1 < ntiles) issue_fwd(it + 1);  // keep the tensor pipe busy while the epilogue works on tile it
      if (!train) continue;
      const int sa = it % AST, s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      if (MODE == 3) mbar_wait(bar(BAR_A_FULL + sa), (it / AST) & 1);  // (no forward product waited for the activation tile)
      mbar_wait(bar(BAR_DZ_FULL + s), ph);
      tc_fence_after();
      if (g.timing && blockIdx.x == 0 && lane == 0) g.timing[it * 8 + 0] = clock64();  // backward products: issue starts
      const uint64_t d_dz_k = smem_desc(sbase + OFF_DZ + s * DZ_BYTES, 16, 1024);       // dz^T as K-major operand (K = teams)
      const uint64_t d_dz_mn = smem_desc(sbase + OFF_DZ + s * DZ_BYTES, CHUNK, 1024);   // dz as MN-major operand (N = teams)
      const uint64_t d_a_mn = smem_desc(sbase + OFF_A16 + sa * A16_BYTES, CHUNK, 1024);  // A16 as MN-major operand (N = hidden)
      // dW[128 j x 128 k] += dz^T[j, n] . A16[n, k] : K = 128 teams = 8 steps of 16
#pragma unroll
      for (int i = 0; i < TB / 16; ++i) {
        const uint64_t koff = (uint64_t)(((i >> 2) * CHUNK + (i & 3) * 32) >> 4);
        if (elect_one()) mma_f16(tmem + TM_DW, d_dz_k + koff, d_a_mn + (uint64_t)((i * 2048) >> 4), IDESC_DW, (it > 0 || i > 0));
      }
      // dA^T[128 k x 128 n] = W16[j, k]^T . dz[j, n] : K = 128 experts = 8 steps of 16; hidden on the TMEM lanes (see the dA warps)
      mbar_wait(bar(BAR_DA_EMPTY), (it & 1) ^ 1);
      tc_fence_after();
      if (g.timing && blockIdx.x == 0 && lane == 0) g.timing[it * 8 + 7] = clock64();  // dW issued, dA accumulator free
#pragma unroll
      for (int i = 0; i < TE / 16; ++i) {
        const uint64_t koff = (uint64_t)((i * 2048) >> 4);
        if (elect_one()) mma_f16(tmem + TM_DA, d_w16_mn + koff, d_dz_mn + koff, IDESC_DA, i > 0);
      }
      if (elect_one()) {
        tc_commit(bar(BAR_DA_FULL));
        tc_commit(bar(BAR_DZ_EMPTY + s));
        tc_commit(bar(BAR_A_EMPTY + sa));
        if (it == ntiles - 1) tc_commit(bar(BAR_DW_FULL));
      }
    }
  } else if (warp < EPI_WARPS) {
    // ====================== logits / loss epilogue: thread = (expert jl, block cb of 32 of the tile's 128 teams) ======================
    const int jl = threadIdx.x & 127;  // TMEM lane = local expert
    const int cb = threadIdx.x >> 7;   // teams [32*cb, 32*cb+32) of the tile
    const int e = e0 + jl;
    const bool e_ok = e < g.E;
    const float bj = e_ok ? __ldg(g.bias + e) : 0.f;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    {
      // fp16 image of the W tile: rows = experts, 64 hidden units per 128-byte row, 2 chunks.  This thread converts hidden units
      // [32*cb, 32*cb+32) of its expert: chunk cb of the fp32 landing zone -> half of a row of chunk cb/2 of the image.
      mbar_wait(bar(BAR_W32), 0);
      const uint8_t* src = sgen + OFF_DZ + cb * CHUNK + jl * 128;
      uint8_t* dst = sgen + OFF_W16 + (cb >> 1) * CHUNK + jl * 128;
#pragma unroll
      for (int m = 0; m < 4; ++m) {  // 8 hidden values -> one 16-byte unit
        const float4 f0 = *reinterpret_cast<const float4*>(src + (((2 * m) ^ (jl & 7)) << 4));
        const float4 f1 = *reinterpret_cast<const float4*>(src + (((2 * m + 1) ^ (jl & 7)) << 4));
        __half2 h0 = __floats2half2_rn(f0.x, f0.y), h1 = __floats2half2_rn(f0.z, f0.w);
        __half2 h2 = __floats2half2_rn(f1.x, f1.y), h3 = __floats2half2_rn(f1.z, f1.w);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + (((4 * (cb & 1) + m) ^ (jl & 7)) << 4)) = pk;
      }
      fence_proxy_async();
      mbar_arrive(bar(BAR_W16));
    }
    float acc_lin = 0.f, acc_lg = 0.f, loss_sp = 0.f, db_acc = 0.f;  // dense loss = tnw*ln2*(acc_lg - acc_lin): sums of lg2(1+e) and of t = -c*x
    float2 acc_lin2 = make_float2(0.f, 0.f), db_acc2 = make_float2(0.f, 0.f);  // the dense pass's packed halves of acc_lin / db_acc
    float2 dbd2 = make_float2(0.f, 0.f);                                        // Flipout: colsum(s_out*dz) of this thread's teams
    // Flipout: bit n of Sg = s_out(team n0+n, my expert) is -1.  One word per tile from the tile-transposed plane, fetched one tile ahead.
    auto sign_word = [&](int it) -> uint32_t {
      return (FLIP && MODE != 3 && g.sign_t && it < ntiles) ? __ldg(g.sign_t + ((size_t)(t_begin + it) * g.Epad + e) * 4 + (threadIdx.x >> 7)) : 0u;
    };
    uint32_t Sg_next = sign_word(0);
    // dz is kept as w*(sigmoid-y)*slope, i.e. true dz / loss_scale; experts past E (last tile) get zero gradient and their loss is dropped below
    const float c_pos = e_ok ? g.tnw : 0.f, c_neg = e_ok ? g.tnw * NTF_LRELU_SLOPE : 0.f;
    const bool has_sp = MODE == 0 && g.special_t && !(g.exp & 2);
    constexpr float LOG2E = 1.4426950408889634f;
    const float kb1 = -LOG2E * bj, kb2 = -LOG2E * NTF_LRELU_SLOPE * bj;
    for (int it = 0; it < (MODE == 3 ? 0 : ntiles); ++it) {  // (MODE 3 has no logits: these warps convert the W tile and drain dW)
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int n0 = (t_begin + it) * TB + cb * 32;  // first team of this thread's block
      const uint32_t Sg = Sg_next;
      Sg_next = sign_word(it + 1);
      uint4 qv[4];
      if (QIN) {  // this thread's 32 values of the perturbation-term tile: unit (4*cb+u) of expert jl (unit-major tile: conflict-free 16-byte reads)
        mbar_wait(bar(BAR_Q_FULL), it & 1);
#pragma unroll
        for (int u = 0; u < 4; ++u) qv[u] = *reinterpret_cast<const uint4*>(sgen + OFF_Q + (((cb * 4 + u) * TE + jl) << 4));
#ifdef NTF_FLIP_EARLY_RELEASE  // the round-1 hand-over, kept to reproduce its race (scripts/flip_variants.sh): the slot is released right after the
        mbar_arrive(bar(BAR_Q_EMPTY));  // loads were ISSUED -- see the release below
#endif
      }
      // bit n of S / Y: (team n0+n, my expert) carries weight tpw / target 1 -- two words from the tile's planes, then the stage is free
      uint32_t S = 0, Y = 0;
      if (has_sp) {
        const uint32_t* plane = reinterpret_cast<const uint32_t*>(sgen + OFF_PLANE + s * 2 * PLANE_BYTES);
        mbar_wait(bar(BAR_SP_FULL + s), ph);
        S = plane[jl * 4 + cb];
        Y = plane[TE * 4 + jl * 4 + cb];
        if (S) {  // consumed: clear the words in HBM so that the caller's planes are clean for the next batch (no second pass over them)
          const size_t wofs = ((size_t)(t_begin + it) * g.Epad + e) * 4 + cb;
          g.special_t[wofs] = 0u;
          if (Y) g.member_t[wofs] = 0u;
        }
        mbar_arrive(bar(BAR_SP_EMPTY + s));  // (after the branch above consumed S and Y: the loads have completed, not merely issued)
        if (!e_ok) S = 0;
      }
      mbar_wait(bar(BAR_Z_FULL + s), ph);
      if (g.timing && blockIdx.x == 0 && threadIdx.x == 0) g.timing[it * 8 + 4] = clock64();
      tc_fence_after();
      float z[32];
      tmem_ld32(tmem + lane_base + TM_Z + s * TB + cb * 32, z);
      tc_fence_before();
      mbar_arrive(bar(BAR_Z_EMPTY + s));
      if (g.timing && blockIdx.x == 0 && threadIdx.x == 0) g.timing[it * 8 + 5] = clock64();
      const int nrem = g.B - n0;  // teams of this block that exist
      if (QIN) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t w[4] = {qv[u].x, qv[u].y, qv[u].z, qv[u].w};
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[p]));
            z[u * 8 + 2 * p] += f.x; z[u * 8 + 2 * p + 1] += f.y;
          }
        }
#ifndef NTF_FLIP_EARLY_RELEASE
        // The single-stage Q slot goes back to its producer only HERE.  Round 1 arrived on Q_EMPTY right after issuing the four LDS.128 above;
        // the producer thread spins on that barrier and re-arms the slot with the NEXT tile's bulk copy (async proxy) the moment the 512th
        // arrival lands, and on a near-empty machine that copy could overtake shared-memory reads that were issued but not yet performed:
        // tile 0's logits then carried tile 1's perturbation term (bisected in profiles/r02a_flipout_bisect_variants.txt).
        // The adds above consumed the loaded registers, so the reads have completed; the proxy fence orders them before the async-proxy write.
        fence_proxy_async();
        mbar_arrive(bar(BAR_Q_EMPTY));
#endif
      }
      if (MODE == 2) {
        // Q = s_out * (Z2 + b_delta) as fp16: unit (4*cb+u) = teams [8u, 8u+8) of this thread's block; lanes write consecutive 16-byte units
        uint4* qt = reinterpret_cast<uint4*>(g.Q + ((size_t)(t_begin + it) * g.Epad + e0) * TB);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t w[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int k = u * 8 + 2 * p;
            const float a = __uint_as_float(__float_as_uint(z[k] + bj) ^ ((Sg << (31 - k)) & 0x80000000u));
            const float b = __uint_as_float(__float_as_uint(z[k + 1] + bj) ^ ((Sg << (30 - k)) & 0x80000000u));
            const __half2 h = __floats2half2_rn(a, b);
            w[p] = *reinterpret_cast<const uint32_t*>(&h);
          }
          qt[(cb * 4 + u) * TE + jl] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        continue;
      }
      if (MODE == 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float p[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float zz = z[u * 8 + q] + bj;
            const float x = fmaxf(zz, NTF_LRELU_SLOPE * zz);
            const float ex = ex2_approx(fabsf(x) * -LOG2E);
            const float r = rcp_approx(1.f + ex);
            p[q] = zz > 0.f ? r : ex * r;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (e_ok && u * 8 + q < nrem) g.P[(size_t)(n0 + u * 8 + q) * g.E + e] = p[q];
        }
        if (g.timing && blockIdx.x == 0 && threadIdx.x == 0) g.timing[it * 8 + 6] = clock64();
        continue;
      }
      if (g.Zdbg) {
#pragma unroll
        for (int n = 0; n < 32; ++n)
          if (e_ok && n < nrem) g.Zdbg[(size_t)(n0 + n) * g.E + e] = z[n] + bj;
      }
      if (train) {
        if (it == 0) mbar_wait(bar(BAR_W16), 0);  // every thread is done with the fp32 landing zone before the dz ring is written
        mbar_wait(bar(BAR_DZ_EMPTY + s), ph ^ 1);
      }
      // this thread's 128-byte row of the dz^T tile (shared-state-space address: no generic -> shared conversion per store); the 16-byte
      // unit of group u is (unit0 + u) ^ (jl & 7) = kx ^ u
      const uint32_t dzrow = sbase + OFF_DZ + s * DZ_BYTES + (cb >> 1) * CHUNK + jl * 128;
      const int unit0 = 4 * (cb & 1);
      const uint32_t kx = (uint32_t)(unit0 ^ (jl & 7));
      // Dense pass: every element as (target 0, weight tnw):  loss = softplus(x) = x + ln(1 + exp(-x)),  dz = tnw*sigmoid(x)*slope,
      // x = lrelu(z + b).  With c = log2(e):  t = -c*x = min(-c*(z+b), -0.01c*(z+b))  (two FFMAs with the bias folded in, one FMNMX),
      // e = 2^t, den = 1 + e, sigmoid(x) = 1/den, softplus(x) = (-t + lg2(den))*ln2.
      // The loop is bound by issue slots, so: (1) the fp32 arithmetic runs as PACKED f32x2 instructions (sm_100: FFMA2 / FADD2 / FMUL2,
      // two elements per issue slot); (2) ONE reciprocal and ONE logarithm per 8 elements: the product tree of the 8 denominators
      // gives lg2 of their product, and walking it back down from 1/product gives every 1/den (1.25 instead of 2.125 MUFU ops per
      // logit).  sum(t), sum(lg2) and db are accumulated in packed registers and combined once per CTA.  8 teams -> one 16-byte unit
      // of the dz^T row.  A product beyond 1e30 (logits below about -70) falls back to per-element logarithms / reciprocals.
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
          if (decltype(masked)::value) {  // teams past the end of the batch
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
          const float2 RQ = __fmul2_rn(make_float2(r, r), make_float2(Q.y, Q.x));  // (1/Q.x, 1/Q.y)
          const float2 R01 = __fmul2_rn(RQ, M23), R23 = __fmul2_rn(RQ, M01);      // 1/M01, 1/M23
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
      if (nrem >= 32) {  // (warp-uniform)
#pragma unroll
        for (int u = 0; u < 4; ++u) dense8(u, std::false_type{});
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) dense8(u, std::true_type{});
      }
      // Sparse fix-up (rare): members of the team and sampled negatives carry weight tpw and the member target.  Their dense
      // contribution is taken back out and the fp16 gradient already staged in shared memory is overwritten.  Popular experts
      // collect many such teams (Zipf), so the work is done warp-wide: for every lane (expert) that has any, its 32 logits are
      // transposed across the warp by shuffles, lane i handles team i, and the corrections return by fixed-order shuffle sums.
      unsigned todo = __ballot_sync(0xffffffffu, S != 0u);
      while (todo) {
        const int L = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t S_L = __shfl_sync(0xffffffffu, S, L), Y_L = __shfl_sync(0xffffffffu, Y, L);
        const float bj_L = __shfl_sync(0xffffffffu, bj, L);
        float zi = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float v = __shfl_sync(0xffffffffu, z[k], L);
          zi = (k == lane) ? v : zi;
        }
        const bool mine = (S_L >> lane) & 1u;  // (team n0 + lane, expert of lane L); bits past the batch end / past E are never set
        const float zz = zi + bj_L;
        const float x = fmaxf(zz, NTF_LRELU_SLOPE * zz);
        const float tt = -LOG2E * x;
        const float exs = ex2_approx(tt), dens = 1.f + exs;          // the dense pass's (signed) terms, taken back out
        const float g_dense = rcp_approx(dens) * (zz > 0.f ? g.tnw : g.tnw * NTF_LRELU_SLOPE);
        const float ex = ex2_approx(fabsf(x) * -LOG2E);               // stable form for the real term
        const float den = 1.f + ex, lg = lg2_approx(den), r = rcp_approx(den);
        const float sig = x > 0.f ? r : ex * r;
        const float yf = ((Y_L >> lane) & 1u) ? 1.f : 0.f;
        const float g_sp = g.tpw * (sig - yf) * (zz > 0.f ? 1.f : NTF_LRELU_SLOPE);
        float d_lin = mine ? -tt : 0.f, d_lg = mine ? -lg2_approx(dens) : 0.f;
        float d_sp = mine ? g.tpw * ((1.f - yf) * x + fmaxf(-x, 0.f) + 0.6931471805599453f * lg) : 0.f;
        float d_db = mine ? g_sp - g_dense : 0.f;
        __syncwarp();
        if (mine && train) {
          const int jl_L = jl - lane + L;
          uint8_t* rowL = sgen + OFF_DZ + s * DZ_BYTES + (cb >> 1) * CHUNK + jl_L * 128;
          *reinterpret_cast<__half*>(rowL + (((unit0 + (lane >> 3)) ^ (jl_L & 7)) << 4) + (lane & 7) * 2) = __float2half_rn(g_sp);
        }
        d_lin = warp_sum(d_lin); d_lg = warp_sum(d_lg); d_sp = warp_sum(d_sp); d_db = warp_sum(d_db);
        if (lane == L) { acc_lin += d_lin; acc_lg += d_lg; loss_sp += d_sp; db_acc += d_db; }
      }
      if (FLIP && MODE == 0 && train) {
        // Flipout: s_out*dz of this thread's (expert, 32 teams) -> the DZS tile in HBM (operand of the perturbation path's backward
        // products, MODE 3) and db_delta.  Read back from the staged fp16 row so that the fix-up's corrections are included.
        __syncwarp();
        __half* drow = g.DZS + ((size_t)(t_begin + it) * g.Epad + e) * TB + cb * 32;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t w[4];
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(dzrow + ((kx ^ (uint32_t)u) << 4)) : "memory");
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int k = u * 8 + 2 * p;
            w[p] ^= ((Sg >> k) & 1u) << 15 | ((Sg >> (k + 1)) & 1u) << 31;
            dbd2 = __fadd2_rn(dbd2, __half22float2(*reinterpret_cast<const __half2*>(&w[p])));
          }
          *reinterpret_cast<uint4*>(drow + u * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      if (train) {
        fence_proxy_async();
        mbar_arrive(bar(BAR_DZ_FULL + s));
      }
      if (g.timing && blockIdx.x == 0 && threadIdx.x == 0) g.timing[it * 8 + 6] = clock64();
    }
    acc_lin += acc_lin2.x + acc_lin2.y;
    db_acc += db_acc2.x + db_acc2.y;
    if (MODE == 0 || MODE == 3) {
      // loss partial of this CTA and db: fixed-order combines (shuffle tree, then warps / team blocks in order)
      float* red = reinterpret_cast<float*>(sgen + OFF_BAR + NUM_BARS * 8 + 16);   // [16]
      float* dbs = reinterpret_cast<float*>(sgen + OFF_DZ);                          // [3][128], the dz stages are idle by now ...
      float* dbs2 = dbs + 3 * TE;                                                    // [3][128] Flipout: the db_delta partials
      if (train) mbar_wait(bar(BAR_DW_FULL), 0);                                     // ... once every MMA that read them has completed
      if (g.timing && blockIdx.x == 0 && threadIdx.x == 0) g.timing[15 * 8 + 3] = clock64();  // all products complete
      if (MODE == 0) {
        const float tot = warp_sum(e_ok ? g.tnw * 0.6931471805599453f * (acc_lg - acc_lin) + loss_sp : 0.f);
        const float dbd = dbd2.x + dbd2.y;
        if (lane == 0) red[warp] = tot;
        if (train && cb > 0) dbs[(cb - 1) * TE + jl] = db_acc;
        if (FLIP && train && cb > 0) dbs2[(cb - 1) * TE + jl] = dbd;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (threadIdx.x == 0) {
          float l = 0.f;
#pragma unroll
          for (int q = 0; q < EPI_WARPS; ++q) l += red[q];
          g.loss_part[blockIdx.x] = l;
        }
        if (train && cb == 0 && e_ok) {
          const float dbv = (((db_acc + dbs[jl]) + dbs[TE + jl]) + dbs[2 * TE + jl]) * g.scale;
          if (shared_tile) atomicAdd(g.db + e, dbv); else g.db[e] = dbv;
          if (FLIP) {
            const float dv = (((dbd + dbs2[jl]) + dbs2[TE + jl]) + dbs2[2 * TE + jl]) * g.scale;
            if (shared_tile) atomicAdd(g.db_delta + e, dv); else g.db_delta[e] = dv;
          }
        }
      }
      if (train) {
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem + lane_base + TM_DW + cb * 32, v);  // this thread's 32 of the 128 dW columns of its expert
        // through shared memory (the idle activation ring; 16-byte units XOR-swizzled by row) so that the 64 KB tile goes out as
        // fully coalesced 512-byte rows
        float4* stage = reinterpret_cast<float4*>(sgen + OFF_A16);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          stage[jl * 32 + ((cb * 8 + q) ^ (jl & 31))] = make_float4(v[4 * q] * g.scale, v[4 * q + 1] * g.scale, v[4 * q + 2] * g.scale, v[4 * q + 3] * g.scale);
        asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll
        for (int i = 0; i < TE / EPI_WARPS; ++i) {
          const int r = warp * (TE / EPI_WARPS) + i;
          if (e0 + r < g.E) {
            const float4 v4 = stage[r * 32 + (lane ^ (r & 31))];
            float* dst = g.dW + (size_t)(e0 + r) * HK + 4 * lane;
            if (shared_tile) { atomicAdd(dst, v4.x); atomicAdd(dst + 1, v4.y); atomicAdd(dst + 2, v4.z); atomicAdd(dst + 3, v4.w); }
            else *reinterpret_cast<float4*>(dst) = v4;
          }
        }
      }
    }
  } else if (warp >= WARP_DA && warp < WARP_DA + 4) {
    // ====== dA epilogue.  The product is taken TRANSPOSED (hidden unit on the TMEM lane, teams along the columns): a warp's 32 lanes hold 32
    // consecutive hidden units of one team, so the tile is added into dA[B,128] by red.global.add.f32 straight from registers, 128
    // contiguous bytes per instruction (round 1 staged 16 KB chunks for a TMA reduce-add through ONE buffer: 4600 cycles per tile, the
    // tile period of the whole kernel; scripts/mb/mb_reduce.cu).  The sum over expert tiles (= across CTAs) happens in L2. ======
    if (train) {
      const int k = threadIdx.x - WARP_DA * 32;  // hidden unit = TMEM lane
      const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
      for (int it = 0; it < ntiles; ++it) {
        mbar_wait(bar(BAR_DA_FULL), it & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < TB / 32; ++c) {
          float v[32];
          tmem_ld32(tmem + lane_base + TM_DA + c * 32, v);
          if (c == TB / 32 - 1) {
            tc_fence_before();
            mbar_arrive(bar(BAR_DA_EMPTY));
          }
          const int team0 = (t_begin + it) * TB + c * 32;
          float* dst = g.dA + (size_t)team0 * HK + k;
          const int nrem = g.B - team0;
          if (!(g.exp & 1)) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nrem) atomicAdd(dst + (size_t)i * HK, v[i] * g.scale);
          }
        }
        if (g.timing && blockIdx.x == 0 && k == 0) g.timing[it * 8 + 3] = clock64();  // dA tile added
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (g.timing && blockIdx.x == 0 && threadIdx.x == 0) g.timing[15 * 8 + 4] = clock64();  // drained
  if (g.timing && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    g.timing[128 + 3 * blockIdx.x + 1] = (long long)gt;
  }
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
  }
}

__global__ void to_half_kernel(const float* __restrict__ x, size_t n, __half* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = __float2half_rn(x[i]);
}

// Flipout: the row-major s_out bit plane [B, pitch] (bit j of word (n, j/32) = 1: s_out[n,j] = -1) -> the tile-transposed plane the
// epilogue threads read (word ((n/128)*Epad + j)*4 + (n%128)/32, bit n%32).  One warp per (tile, block of 32 teams, word column):
// lane i holds the word of team i, 32 ballots transpose the 32x32 bit block, lane j keeps expert j's word.
__global__ void sign_tiles_kernel(const uint32_t* __restrict__ sign, int B, int pitch, int Epad, int ntiles, uint32_t* __restrict__ sign_t) {
  const int wcols = Epad / 32;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (long long)ntiles * 4 * wcols) return;
  const int wc = (int)(gw % wcols), cb = (int)((gw / wcols) & 3), t = (int)(gw / (4LL * wcols));
  const int n = t * TB + cb * 32 + lane;
  const uint32_t w = (n < B && wc < pitch) ? __ldg(sign + (size_t)n * pitch + wc) : 0u;
  uint32_t mine = 0u;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const uint32_t r = __ballot_sync(0xffffffffu, (w >> j) & 1u);
    if (j == lane) mine = r;
  }
  sign_t[((size_t)t * Epad + wc * 32 + lane) * 4 + cb] = mine;
}

}  // namespace

int ntf_out_tc_supported(int B, int h, int E, int flipout) { (void)flipout; return (h == HK && B >= 1 && E >= 1) ? 1 : 0; }

// Flipout workspace behind the Fnn part: fp16 A_s | s_out tile plane | Q tiles | DZS tiles
struct FlipWs { size_t as16, sign_t, q, dzs, total; };
static FlipWs flip_ws(int B, int h, int E, bool train) {
  const size_t Epad = (size_t)cdiv(E, TE) * TE, nt = (size_t)cdiv(B, TB);
  FlipWs w;
  w.as16 = 0;
  w.sign_t = w.as16 + align_up((size_t)B * h * sizeof(__half), 1024);
  w.q = w.sign_t + align_up(nt * Epad * 4 * sizeof(uint32_t), 1024);
  w.dzs = w.q + align_up(nt * Epad * TB * sizeof(__half), 1024);
  w.total = w.dzs + (train ? align_up(nt * Epad * TB * sizeof(__half), 1024) : 0);
  return w;
}

static size_t tc_base_ws(int B, int h, int E) {
  return align_up((size_t)B * h * sizeof(__half), 256) + align_up((size_t)(cdiv(E, TE) + 1024) * sizeof(float), 1024);  // + one loss partial per CTA
}

size_t ntf_out_train_tc_workspace_bytes(const ntf_ctx*, int B, int h, int E, int flipout) {
  return tc_base_ws(B, h, E) + (flipout ? flip_ws(B, h, E, true).total : 0);
}

// Work split (TcArgs::n_full / split): the expert tiles that fill whole waves of the machine get one CTA each; the tiles of the
// partial last wave are cut along the batch so that the last wave is short instead of leaving most SMs idle for a full tile.
static void plan_split(const ntf_ctx* ctx, int nct, int nt_all, int* n_full, int* split, int* grid) {
  const int sm = ctx->sm_count > 0 ? ctx->sm_count : 148;
  const int rem = nct % sm;
  int sp = (rem > 0 && nct > sm) ? sm / rem : 1;   // (a grid smaller than one wave is left alone)
  if (sp > nt_all) sp = nt_all;
  if (sp < 2) { *n_full = nct; *split = 1; *grid = nct; return; }
  *n_full = nct - rem; *split = sp; *grid = *n_full + rem * sp;
}

static void tc_debug_hooks(TcArgs& g) {
  const char* dbg = getenv("NTF_TC_ZDBG");  // debug hook used by tests/test_gpu_tc.py: address of a [B,E] device buffer for the raw logits
  g.Zdbg = dbg ? (float*)(uintptr_t)strtoull(dbg, nullptr, 0) : nullptr;
  const char* tim = getenv("NTF_TC_TIMING");
  g.timing = tim ? (long long*)(uintptr_t)strtoull(tim, nullptr, 0) : nullptr;
  const char* ex = getenv("NTF_TC_EXP");
  g.exp = ex ? atoi(ex) : 0;
}

template <int MODE>
static int launch_flip(cudaStream_t st, int grid, const CUtensorMap& mw, const CUtensorMap& ma, const CUtensorMap& mda, const CUtensorMap& mdz, const TcArgs& g) {
  NTF_CUDA(cudaFuncSetAttribute(out_tc_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  NTF_COUNT_LAUNCH; out_tc_kernel<MODE, true><<<grid, NT, SMEM_BYTES, st>>>(mw, ma, mda, mdz, g);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

// fp16 copies + the s_out tile plane + MODE 2 (Q = s_out*(A_s W_delta^T + b_delta)); shared by the training and the inference entry points
static int flip_forward(ntf_ctx* ctx, cudaStream_t st, const float* A_s, const float* W_delta, const float* b_delta, const uint32_t* sign_out, int pitch_words,
                        int B, int E, char* fws, const FlipWs& fw, TcArgs* gq, CUtensorMap* mwd, CUtensorMap* mas) {
  const int Epad = cdiv(E, TE) * TE, nt = cdiv(B, TB);
  __half* As16 = (__half*)(fws + fw.as16);
  uint32_t* sign_t = (uint32_t*)(fws + fw.sign_t);
  int rc;
  if ((rc = make_map(ctx, mwd, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, W_delta, (uint64_t)E, HK, TE, 32))) return rc;
  if ((rc = make_map(ctx, mas, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, As16, (uint64_t)B, HK, TB, 64))) return rc;
  NTF_COUNT_LAUNCH; to_half_kernel<<<min(cdiv(B * HK, 256), ctx->sm_count * 4), 256, 0, st>>>(A_s, (size_t)B * HK, As16);
  const long long warps = (long long)nt * 4 * (Epad / 32);
  NTF_COUNT_LAUNCH; sign_tiles_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(sign_out, B, pitch_words, Epad, nt, sign_t);
  TcArgs g{};
  g.bias = b_delta; g.Epad = Epad; g.B = B; g.E = E;
  g.sign_t = sign_t; g.Q = (__half*)(fws + fw.q);
  tc_debug_hooks(g);
  g.Zdbg = nullptr; g.timing = nullptr;
  int grid;
  plan_split(ctx, cdiv(E, TE), nt, &g.n_full, &g.split, &grid);
  *gq = g;
  return launch_flip<2>(st, grid, *mwd, *mas, *mwd, *mwd, g);
}

// Flipout training step of the output layer: three launches of the same pipeline (out_tc_kernel modes 2, 0, 3)
static int out_train_tc_flip(ntf_ctx* ctx, cudaStream_t st, const ntf_out_train_args* a, __half* A16, float* loss_part, char* fws) {
  const bool train = a->dW != nullptr;
  const FlipWs fw = flip_ws(a->B, a->h, a->E, true);
  const int Epad = cdiv(a->E, TE) * TE, nt = cdiv(a->B, TB), nct = cdiv(a->E, TE);
  CUtensorMap mw, mh, mda, mwd, mas, mdas, mdz;
  TcArgs gq;
  int rc;
  if ((rc = flip_forward(ctx, st, a->A_s, a->W_delta, a->b_delta, a->sign_out, a->pitch_words, a->B, a->E, fws, fw, &gq, &mwd, &mas))) return rc;
  __half* DZS = (__half*)(fws + fw.dzs);
  if ((rc = make_map(ctx, &mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->W, (uint64_t)a->E, HK, TE, 32))) return rc;
  if ((rc = make_map(ctx, &mh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, A16, (uint64_t)a->B, HK, TB, 64))) return rc;
  if ((rc = make_map(ctx, &mda, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, train ? (const void*)a->dA : (const void*)a->W, (uint64_t)(train ? a->B : a->E), HK, TB, 32))) return rc;
  if ((rc = make_map(ctx, &mdas, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, train ? (const void*)a->dA_s : (const void*)a->W, (uint64_t)(train ? a->B : a->E), HK, TB, 32))) return rc;
  if ((rc = make_map(ctx, &mdz, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, DZS, (uint64_t)nt * Epad, TB, TE, 64))) return rc;
  if (!a->A16) { NTF_COUNT_LAUNCH; to_half_kernel<<<min(cdiv(a->B * a->h, 256), ctx->sm_count * 4), 256, 0, st>>>(a->A, (size_t)a->B * a->h, A16); }
  if (train) {
    NTF_CUDA(cudaMemsetAsync(a->dA, 0, (size_t)a->B * a->h * sizeof(float), st));
    NTF_CUDA(cudaMemsetAsync(a->dA_s, 0, (size_t)a->B * a->h * sizeof(float), st));
  }
  TcArgs g = gq;  // same tiling / split as the Q launch
  g.bias = a->b; g.special_t = const_cast<uint32_t*>(a->special_t); g.member_t = const_cast<uint32_t*>(a->member_t);
  g.tpw = a->tpw; g.tnw = a->tnw; g.scale = a->loss_scale;
  g.dW = a->dW; g.db = a->db; g.dA = a->dA; g.loss_part = loss_part; g.DZS = DZS; g.db_delta = a->db_delta;
  tc_debug_hooks(g);
  const int grid = g.n_full + (nct - g.n_full) * g.split;
  if (train && g.split > 1) {  // the shared tiles' CTAs add their partials
    const size_t e_first = (size_t)g.n_full * TE, rows = (size_t)a->E - e_first;
    NTF_CUDA(cudaMemsetAsync(a->dW + e_first * HK, 0, rows * HK * sizeof(float), st));
    NTF_CUDA(cudaMemsetAsync(a->db + e_first, 0, rows * sizeof(float), st));
    NTF_CUDA(cudaMemsetAsync(a->dW_delta + e_first * HK, 0, rows * HK * sizeof(float), st));
    NTF_CUDA(cudaMemsetAsync(a->db_delta + e_first, 0, rows * sizeof(float), st));
  }
  if ((rc = launch_flip<0>(st, grid, mw, mh, mda, mda, g))) return rc;
  if (train) {
    TcArgs g3 = gq;
    g3.bias = a->b_delta; g3.scale = a->loss_scale; g3.dW = a->dW_delta; g3.dA = a->dA_s; g3.Q = nullptr; g3.sign_t = nullptr;
    if ((rc = launch_flip<3>(st, grid, mwd, mas, mdas, mdz, g3))) return rc;
  }
  return ntf_loss_reduce_impl(st, loss_part, grid, a->loss_scale, a->loss_out);
}

int ntf_out_train_tc(ntf_ctx* ctx, cudaStream_t st, const ntf_out_train_args* a, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(a->h == HK, NTF_ERR_UNSUPPORTED, "out_train(tf32): hidden width %d (kernel is built for %d)", a->h, HK);
  const bool flip = a->W_delta != nullptr;
  NTF_REQUIRE(workspace_bytes >= ntf_out_train_tc_workspace_bytes(ctx, a->B, a->h, a->E, flip), NTF_ERR_WORKSPACE, "out_train(tf32): workspace too small");
  NTF_REQUIRE(((uintptr_t)workspace & 255) == 0, NTF_ERR_BAD_ARG, "out_train(tf32): the workspace must be 256-byte aligned");
  NTF_REQUIRE((((uintptr_t)a->A | (uintptr_t)a->W) & 15) == 0, NTF_ERR_BAD_ARG, "out_train(tf32): A and W must be 16-byte aligned");
  const bool train = a->dW != nullptr;
  NTF_REQUIRE(!train || (a->db && a->dA), NTF_ERR_BAD_ARG, "out_train(tf32): training needs dW, db and dA");
  NTF_REQUIRE(!train || (((uintptr_t)a->dA | (uintptr_t)a->dW) & 15) == 0, NTF_ERR_BAD_ARG, "out_train(tf32): dA and dW must be 16-byte aligned");
  if (flip) {
    NTF_REQUIRE(a->A_s && a->b_delta && a->sign_out, NTF_ERR_BAD_ARG, "out_train(tf32): incomplete Flipout arguments");
    NTF_REQUIRE(!train || (a->dW_delta && a->db_delta && a->dA_s), NTF_ERR_BAD_ARG, "out_train(tf32): Flipout training needs dW_delta, db_delta and dA_s");
    NTF_REQUIRE((((uintptr_t)a->A_s | (uintptr_t)a->W_delta | (uintptr_t)a->dW_delta | (uintptr_t)a->dA_s) & 15) == 0, NTF_ERR_BAD_ARG, "out_train(tf32): Flipout tensors must be 16-byte aligned");
    NTF_REQUIRE(a->e_lo == 0, NTF_ERR_UNSUPPORTED, "out_train(tf32): expert-sharded Flipout layer");
  }
  NTF_REQUIRE((a->special_t == nullptr) == (a->member_t == nullptr), NTF_ERR_BAD_ARG, "out_train(tf32): special_t and member_t come together");
  NTF_REQUIRE(a->special_t || !a->special, NTF_ERR_BAD_ARG, "out_train(tf32): the tensor-core kernel reads the tile-transposed planes (ntf_special_tiles), not `special`");
  __half* A16 = a->A16 ? (__half*)const_cast<void*>(a->A16) : (__half*)workspace;
  NTF_REQUIRE(((uintptr_t)A16 & 15) == 0, NTF_ERR_BAD_ARG, "out_train(tf32): A16 must be 16-byte aligned");
  float* loss_part = (float*)((char*)workspace + align_up((size_t)a->B * a->h * sizeof(__half), 256));
  const int nct = cdiv(a->E, TE);
  CUtensorMap mw, mh, mda;
  int rc;
  if (flip) return out_train_tc_flip(ctx, st, a, A16, loss_part, (char*)workspace + tc_base_ws(a->B, a->h, a->E));
  if ((rc = make_map(ctx, &mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->W, (uint64_t)a->E, HK, TE, 32))) return rc;
  if ((rc = make_map(ctx, &mh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, A16, (uint64_t)a->B, HK, TB, 64))) return rc;
  if ((rc = make_map(ctx, &mda, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, train ? (const void*)a->dA : (const void*)a->W, (uint64_t)(train ? a->B : a->E), HK, TB, 32))) return rc;
  if (!a->A16) { NTF_COUNT_LAUNCH; to_half_kernel<<<min(cdiv(a->B * a->h, 256), ctx->sm_count * 4), 256, 0, st>>>(a->A, (size_t)a->B * a->h, A16); }
  if (train) NTF_CUDA(cudaMemsetAsync(a->dA, 0, (size_t)a->B * a->h * sizeof(float), st));
  TcArgs g{};
  g.bias = a->b; g.special_t = const_cast<uint32_t*>(a->special_t); g.member_t = const_cast<uint32_t*>(a->member_t); g.Epad = cdiv(a->E, TE) * TE;
  g.B = a->B; g.E = a->E; g.tpw = a->tpw; g.tnw = a->tnw; g.scale = a->loss_scale;
  g.dW = a->dW; g.db = a->db; g.dA = a->dA; g.loss_part = loss_part; g.P = nullptr;
  tc_debug_hooks(g);
  int grid;
  plan_split(ctx, nct, cdiv(a->B, TB), &g.n_full, &g.split, &grid);
  if (g.exp & 8) { g.n_full = nct; g.split = 1; grid = nct; }  // debug: no split
  if (train && g.split > 1) {  // the shared tiles' CTAs add their partial dW / db
    const size_t e_first = (size_t)g.n_full * TE;
    NTF_CUDA(cudaMemsetAsync(a->dW + e_first * HK, 0, ((size_t)a->E - e_first) * HK * sizeof(float), st));
    NTF_CUDA(cudaMemsetAsync(a->db + e_first, 0, ((size_t)a->E - e_first) * sizeof(float), st));
  }
  NTF_CUDA(cudaFuncSetAttribute(out_tc_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  NTF_COUNT_LAUNCH; out_tc_kernel<0, false><<<grid, NT, SMEM_BYTES, st>>>(mw, mh, mda, mda, g);
  NTF_LAUNCH_CHECK();
  return ntf_loss_reduce_impl(st, loss_part, grid, a->loss_scale, a->loss_out);
}

size_t ntf_infer_scores_tc_workspace_bytes(int B, int h) { return align_up((size_t)B * h * sizeof(__half), 256); }

int ntf_infer_scores_tc(ntf_ctx* ctx, cudaStream_t st, const float* A, const float* W, const float* b, int B, int h, int E, float* P,
                        void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(h == HK, NTF_ERR_UNSUPPORTED, "infer_scores(tf32): hidden width %d (kernel is built for %d)", h, HK);
  NTF_REQUIRE(workspace && workspace_bytes >= ntf_infer_scores_tc_workspace_bytes(B, h), NTF_ERR_WORKSPACE, "infer_scores(tf32): workspace too small");
  __half* A16 = (__half*)workspace;
  CUtensorMap mw, mh;
  int rc;
  if ((rc = make_map(ctx, &mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, W, (uint64_t)E, HK, TE, 32))) return rc;
  if ((rc = make_map(ctx, &mh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, A16, (uint64_t)B, HK, TB, 64))) return rc;
  NTF_COUNT_LAUNCH; to_half_kernel<<<min(cdiv(B * h, 256), ctx->sm_count * 4), 256, 0, st>>>(A, (size_t)B * h, A16);
  TcArgs g{};
  g.bias = b; g.B = B; g.E = E; g.P = P;
  tc_debug_hooks(g);
  g.Zdbg = nullptr;
  int grid;
  plan_split(ctx, cdiv(E, TE), cdiv(B, TB), &g.n_full, &g.split, &grid);
  NTF_CUDA(cudaFuncSetAttribute(out_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  NTF_COUNT_LAUNCH; out_tc_kernel<1, false><<<grid, NT, SMEM_BYTES, st>>>(mw, mh, mw, mw, g);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

// Flipout inference (fnn.py:202-209 with a Bnn): P = sigmoid(lrelu(A mu^T + mu_b + s_out*(A_s W_delta^T + b_delta))): MODE 2 then MODE 1
size_t ntf_infer_scores_tc_flip_workspace_bytes(int B, int h, int E) { return align_up((size_t)B * h * sizeof(__half), 1024) + flip_ws(B, h, E, false).total; }

int ntf_infer_scores_tc_flip(ntf_ctx* ctx, cudaStream_t st, const float* A, const float* W, const float* b, int B, int h, int E, const float* A_s,
                             const float* W_delta, const float* b_delta, const uint32_t* sign_out, int pitch_words, float* P, void* workspace,
                             size_t workspace_bytes) {
  NTF_REQUIRE(h == HK, NTF_ERR_UNSUPPORTED, "infer_scores(tf32): hidden width %d (kernel is built for %d)", h, HK);
  NTF_REQUIRE(workspace && workspace_bytes >= ntf_infer_scores_tc_flip_workspace_bytes(B, h, E), NTF_ERR_WORKSPACE, "infer_scores(tf32, Flipout): workspace too small");
  NTF_REQUIRE((((uintptr_t)workspace & 255) | (((uintptr_t)A | (uintptr_t)W | (uintptr_t)A_s | (uintptr_t)W_delta) & 15)) == 0, NTF_ERR_BAD_ARG, "infer_scores(tf32, Flipout): misaligned pointer");
  __half* A16 = (__half*)workspace;
  char* fws = (char*)workspace + align_up((size_t)B * h * sizeof(__half), 1024);
  const FlipWs fw = flip_ws(B, h, E, false);
  CUtensorMap mw, mh, mwd, mas;
  TcArgs gq;
  int rc;
  if ((rc = flip_forward(ctx, st, A_s, W_delta, b_delta, sign_out, pitch_words, B, E, fws, fw, &gq, &mwd, &mas))) return rc;
  if ((rc = make_map(ctx, &mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, W, (uint64_t)E, HK, TE, 32))) return rc;
  if ((rc = make_map(ctx, &mh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, A16, (uint64_t)B, HK, TB, 64))) return rc;
  NTF_COUNT_LAUNCH; to_half_kernel<<<min(cdiv(B * h, 256), ctx->sm_count * 4), 256, 0, st>>>(A, (size_t)B * h, A16);
  TcArgs g = gq;
  g.bias = b; g.P = P;
  const int grid = g.n_full + (cdiv(E, TE) - g.n_full) * g.split;
  return launch_flip<1>(st, grid, mw, mh, mw, mw, g);
}
