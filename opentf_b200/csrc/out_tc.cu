// out_tc.cu -- NTF_TF32 mode: the output layer on tcgen05 tensor cores with TMEM accumulators and TMA-fed operands.
//
// One fused kernel per training step of the hidden->all-experts layer (fnn.py:25 last layer, 32-46, 135 and its
// autograd): logits, weighted BCE, dz, dW, db and dA are produced without the [B,E] logits or gradients ever
// touching HBM.  Structure = FlashAttention backward with Q<->A (hidden activations), K<->W, dS<->dz:
//
//   CTA e owns experts [128e, 128e+128): W tile resident in shared memory (fp32 for the forward, an fp16 image for
//   the backward), dW accumulator resident in TMEM for the whole kernel, db in registers.
//   It streams over the batch in tiles of 64 teams (double buffered):
//     TMA warp      : A tile (fp32 + fp16 copy) -> smem                                     [cp.async.bulk.tensor]
//     MMA thread    : Z^T[128e x 64n] = W . A^T          kind::tf32, fp32 accumulate in TMEM  (forward, TF32 logits)
//     8 epilogue warps (thread = expert x half of the tile's teams): TMEM -> regs, +b, lrelu, weighted BCE, dz; loss and db accumulate in
//                     registers; dz (scaled by 1/loss_scale, fp16 = same 10-bit mantissa as tf32) -> smem
//     MMA thread    : dW[128e x 128k] += dz^T . A        kind::f16, accumulates across all batch tiles in TMEM
//                     dA^T[128k x 64n] = W^T . dz^T      kind::f16
//     4 dA warps    : TMEM -> regs -> red.global.add into dA[B,128]  (sum over expert tiles = across CTAs)
//   At the end the epilogue warps drain dW from TMEM, write db, and one loss partial per CTA.
//
// Why fp16 for the two backward products: their operands are needed in MN-major form (contraction over teams / over
// experts), which tcgen05 supports for 32-bit types only through a dedicated 32B-atom swizzle that cannot share a
// buffer with the K-major forward operands; 16-bit operands allow both majors on one 128B-swizzled image, halve
// shared memory and run at twice the tf32 rate.  dz is scaled into fp16's normal range, so the gradient products
// carry the same 10-bit mantissa as TF32.
//
// Roofline: tensor pipe.  Algorithmic flops per team = 6*h*E.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

int ntf_loss_reduce_impl(cudaStream_t st, const float* part, int n, float scale, float* loss_out);

namespace {

constexpr int TE = 128;  // experts per CTA tile (UMMA M of the forward / dW products)
constexpr int TB = 64;   // teams per batch tile (UMMA N of the forward / dA products)
constexpr int HK = 128;  // hidden width this kernel is built for

constexpr uint32_t W32_BYTES = TE * HK * 4;      // 64 KB : 4 k-chunks of [128 rows][128 B]
constexpr uint32_t A32_BYTES = TB * HK * 4;      // 32 KB : 4 k-chunks of [ 64 rows][128 B]
constexpr uint32_t W16_BYTES = TE * HK * 2;      // 32 KB : 2 k-chunks of [128 rows j][128 B]
constexpr uint32_t A16_BYTES = TB * HK * 2;      // 16 KB : 2 k-chunks of [ 64 rows n][128 B]
constexpr uint32_t DZ_BYTES = TE * TB * 2;       // 16 KB : [128 rows j][64 n halfs = 128 B]
constexpr uint32_t OFF_W32 = 0;
constexpr uint32_t OFF_A32 = OFF_W32 + W32_BYTES;           // 2 stages
constexpr uint32_t OFF_W16 = OFF_A32 + 2 * A32_BYTES;
constexpr uint32_t OFF_A16 = OFF_W16 + W16_BYTES;           // 2 stages
constexpr uint32_t OFF_DZ = OFF_A16 + 2 * A16_BYTES;        // 2 stages
constexpr uint32_t OFF_PLANE = OFF_DZ + 2 * DZ_BYTES;       // special / target bit planes of the current tile: 2 x [128 experts][2 words]
constexpr uint32_t PLANE_BYTES = TE * 2 * 4;                // 1 KB each
constexpr uint32_t OFF_BAR = OFF_PLANE + 2 * PLANE_BYTES;   // mbarriers + tmem base
constexpr uint32_t SMEM_TRAIN = OFF_BAR + 256;              // the dynamic window is declared 1024-byte aligned (checked at run time)
constexpr uint32_t SMEM_INFER = OFF_W16 + 256;              // forward only: W32 + 2 x A32
static_assert(SMEM_TRAIN <= 232448, "over the 227 KB shared memory limit");

// TMEM columns (fp32 accumulators, 128 lanes each)
constexpr uint32_t TM_Z = 0;      // 2 x 64
constexpr uint32_t TM_DW = 128;   // 128
constexpr uint32_t TM_DA = 256;   // 2 x 64
constexpr uint32_t TM_COLS = 512;

enum { BAR_W = 0, BAR_A_FULL = 1, BAR_A_EMPTY = 3, BAR_Z_FULL = 5, BAR_Z_EMPTY = 7, BAR_DZ_FULL = 9, BAR_DZ_EMPTY = 11,
       BAR_DA_FULL = 13, BAR_DA_EMPTY = 15, BAR_DW_FULL = 17, BAR_W16 = 18, BAR_H_FULL = 19, BAR_H_EMPTY = 21, BAR_SP_FULL = 23, BAR_SP_EMPTY = 24,
       NUM_BARS = 25 };
// A_* : fp32 activation tile ring (forward operand, released as soon as the forward product has read it)
// H_* : fp16 activation tile ring (dW operand, released after the backward products)
// SP_*: special / target bit planes of one tile (single stage: the epilogue copies its two words to registers and releases it at once)

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets TMEM lane (lane_base + t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors (cute/arch/mma_sm100_desc.hpp bit layout) ------------------------------------------------------
// shared memory matrix descriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | swizzle [61,64) (2 = 128B)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
         (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fmt [4,6) 1=f32 | A fmt [7,10) | B fmt [10,13) (0=f16, 2=tf32) | A major [15] | B major [16] (1 = MN) | N>>3 [17,23) | M>>4 [24,29)
constexpr uint32_t instr_desc(uint32_t fmt, uint32_t a_mn, uint32_t b_mn, uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
constexpr uint32_t IDESC_FWD = instr_desc(2, 0, 0, TE, TB);   // Z^T  = W(K-major) . A(K-major)^T            tf32, K = hidden
constexpr uint32_t IDESC_DW = instr_desc(0, 0, 1, TE, HK);    // dW  += dz^T(K-major: K = teams) . A16(MN-major)  f16
constexpr uint32_t IDESC_DA = instr_desc(0, 1, 1, HK, TB);    // dA^T = W16(MN-major: M = hidden) . dz^T(MN-major: N = teams)  f16, K = experts

struct TcArgs {
  const float* bias;          // [E]
  const uint32_t* special;    // [B, pitch] or NULL
  int pitch;
  const int32_t* m_indptr;
  const int32_t* m_indices;
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
};

constexpr int NT = 512;  // warps 0-7: logits/loss epilogue, 8-11: dA epilogue, 12: TMA (fp32 ring), 13: MMA issuer, 14: TMA (fp16 ring), 15: special planes
constexpr int WARP_TMA = 12, WARP_MMA = 13, WARP_TMA16 = 14, WARP_SP = 15;

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// MODE 0: training / validation step.  MODE 1: inference scores P = sigmoid(lrelu(z)).
template <int MODE>
__global__ void __launch_bounds__(NT, 1) out_tc_kernel(const __grid_constant__ CUtensorMap map_w32, const __grid_constant__ CUtensorMap map_a32,
                                                       const __grid_constant__ CUtensorMap map_a16, TcArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if ((sbase & 1023u) != 0u) __trap();  // the 128-byte swizzle atoms need a 1024-byte aligned base
  uint8_t* sgen = smem_raw;
  constexpr uint32_t BAR_OFF = MODE == 0 ? OFF_BAR : OFF_W16;
  const uint32_t bars = sbase + BAR_OFF;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sgen + BAR_OFF + NUM_BARS * 8);
  auto bar = [&](int i) { return bars + 8u * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e0 = blockIdx.x * TE;
  const int ntiles = (g.B + TB - 1) / TB;
  const bool train = MODE == 0 && g.dW != nullptr;

  if (threadIdx.x == 0) {
    mbar_init(bar(BAR_W), 1);
    mbar_init(bar(BAR_W16), 384);
    mbar_init(bar(BAR_DW_FULL), 1);
    mbar_init(bar(BAR_SP_FULL), 1);
    mbar_init(bar(BAR_SP_EMPTY), 256);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(BAR_A_FULL + s), 1);
      mbar_init(bar(BAR_A_EMPTY + s), 1);
      mbar_init(bar(BAR_H_FULL + s), 1);
      mbar_init(bar(BAR_H_EMPTY + s), 1);
      mbar_init(bar(BAR_Z_FULL + s), 1);
      mbar_init(bar(BAR_Z_EMPTY + s), 256);
      mbar_init(bar(BAR_DZ_FULL + s), 256);
      mbar_init(bar(BAR_DZ_EMPTY + s), 1);
      mbar_init(bar(BAR_DA_FULL + s), 1);
      mbar_init(bar(BAR_DA_EMPTY + s), 128);
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

  if (warp == WARP_TMA) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      mbar_expect_tx(bar(BAR_W), W32_BYTES);
      for (int c = 0; c < 4; ++c) tma_load_2d(sbase + OFF_W32 + c * (TE * 128), &map_w32, c * 32, e0, bar(BAR_W));
      for (int t = 0; t < ntiles; ++t) {
        const int s = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        mbar_wait(bar(BAR_A_EMPTY + s), ph ^ 1);
        mbar_expect_tx(bar(BAR_A_FULL + s), A32_BYTES);
        for (int c = 0; c < 4; ++c) tma_load_2d(sbase + OFF_A32 + s * A32_BYTES + c * (TB * 128), &map_a32, c * 32, t * TB, bar(BAR_A_FULL + s));
      }
    }
  } else if (warp == WARP_TMA16) {
    if (lane == 0 && train) {
      for (int t = 0; t < ntiles; ++t) {
        const int s = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        mbar_wait(bar(BAR_H_EMPTY + s), ph ^ 1);
        mbar_expect_tx(bar(BAR_H_FULL + s), A16_BYTES);
        for (int c = 0; c < 2; ++c) tma_load_2d(sbase + OFF_A16 + s * A16_BYTES + c * (TB * 128), &map_a16, c * 64, t * TB, bar(BAR_H_FULL + s));
      }
    }
  } else if (warp == WARP_SP) {
    // ====== special-plane producer: bit planes [expert][team] of the tile, transposed from the [team][expert] plane in HBM ======
    // plane_s bit = weight tpw (member or sampled negative), plane_y bit = target 1 (member).  Off the epilogue's critical path:
    // the words of the next tile are in registers before the current one is released.
    if (MODE == 0) {
      uint32_t* plane_s = reinterpret_cast<uint32_t*>(sgen + OFF_PLANE);
      uint32_t* plane_y = plane_s + TE * 2;
      const int wi0 = e0 >> 5;
      uint32_t w[8];
      auto fetch = [&](int t) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int n = t * TB + r * 32 + lane;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            w[r * 4 + c] = (g.special && !(g.exp & 2) && n < g.B && wi0 + c < g.pitch) ? __ldg(g.special + (size_t)n * g.pitch + wi0 + c) : 0u;
        }
      };
      if (ntiles > 0) fetch(0);
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(bar(BAR_SP_EMPTY), (t & 1) ^ 1);
        for (int i = lane; i < TE * 4; i += 32) plane_s[i] = 0u;  // both planes are contiguous
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int nl = r * 32 + lane;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t bits = w[r * 4 + c];
            while (bits) {
              const int jb = __ffs(bits) - 1;
              bits &= bits - 1;
              const int jl = c * 32 + jb;
              atomicOr(&plane_s[jl * 2 + (nl >> 5)], 1u << (nl & 31));
              if (is_member(g.m_indptr, g.m_indices, t * TB + nl, e0 + jl)) atomicOr(&plane_y[jl * 2 + (nl >> 5)], 1u << (nl & 31));
            }
          }
        }
        if (t + 1 < ntiles) fetch(t + 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_SP_FULL));
      }
    }
  } else if (warp == WARP_MMA) {
    // =========================================== MMA issuer ===========================================
    if (lane == 0) {
      auto issue_fwd = [&](int t) {
        const int s = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        mbar_wait(bar(BAR_A_FULL + s), ph);
        if (g.timing && blockIdx.x == 0) g.timing[t * 8 + 1] = clock64();
        mbar_wait(bar(BAR_Z_EMPTY + s), ph ^ 1);
        if (g.timing && blockIdx.x == 0) g.timing[t * 8 + 2] = clock64();
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < HK / 8; ++i) {  // 16 k-steps of 8 tf32 (32 bytes) each: chunk i/4, 32-byte slice i%4 of the 128-byte swizzle row
          const uint64_t da = smem_desc(sbase + OFF_W32 + (i >> 2) * (TE * 128) + (i & 3) * 32, 16, 1024);
          const uint64_t db = smem_desc(sbase + OFF_A32 + s * A32_BYTES + (i >> 2) * (TB * 128) + (i & 3) * 32, 16, 1024);
          mma_tf32(tmem + TM_Z + s * TB, da, db, IDESC_FWD, i > 0);
        }
        tc_commit(bar(BAR_Z_FULL + s));
        if (g.timing) {  // diagnostic mode serialises the pipeline: stamp the COMPLETION of the forward product
          mbar_wait(bar(BAR_Z_FULL + s), ph);
          if (blockIdx.x == 0) g.timing[t * 8 + 3] = clock64();
        }
        tc_commit(bar(BAR_A_EMPTY + s));  // the fp32 stage is free once the forward product has read it
      };
      mbar_wait(bar(BAR_W), 0);
      if (ntiles > 0) issue_fwd(0);
      if (train) mbar_wait(bar(BAR_W16), 0);
      for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) issue_fwd(t + 1);  // keep the tensor pipe busy while the epilogue works on tile t
        if (!train) continue;
        const int s = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        mbar_wait(bar(BAR_DZ_FULL + s), ph);
        mbar_wait(bar(BAR_H_FULL + s), ph);
        mbar_wait(bar(BAR_DA_EMPTY + s), ph ^ 1);
        tc_fence_after();
        if (g.timing && blockIdx.x == 0) g.timing[t * 8 + 0] = clock64();  // (overwrites the producer stamp) backward products: issue starts
        // dW[128 j x 128 k] += dz^T[j, n] . A16[n, k] : K = 64 teams = 4 steps of 16
#pragma unroll
        for (int i = 0; i < TB / 16; ++i) {
          const uint64_t da = smem_desc(sbase + OFF_DZ + s * DZ_BYTES + i * 32, 16, 1024);                        // K-major, K = teams
          const uint64_t db = smem_desc(sbase + OFF_A16 + s * A16_BYTES + i * 2048, TB * 128, 1024);               // MN-major, N = hidden
          mma_f16(tmem + TM_DW, da, db, IDESC_DW, (t > 0 || i > 0));
        }
        // dA^T[128 k x 64 n] = W16^T[k, j] . dz[j, n] : K = 128 experts = 8 steps of 16
#pragma unroll
        for (int i = 0; i < TE / 16; ++i) {
          const uint64_t da = smem_desc(sbase + OFF_W16 + i * 2048, TE * 128, 1024);                               // MN-major, M = hidden
          const uint64_t db = smem_desc(sbase + OFF_DZ + s * DZ_BYTES + i * 2048, 0, 1024);                        // MN-major, N = teams
          mma_f16(tmem + TM_DA + s * TB, da, db, IDESC_DA, i > 0);
        }
        tc_commit(bar(BAR_DA_FULL + s));
        if (g.timing) {
          mbar_wait(bar(BAR_DA_FULL + s), ph);
          if (blockIdx.x == 0) g.timing[t * 8 + 7] = clock64();  // backward products complete
        }
        tc_commit(bar(BAR_DZ_EMPTY + s));
        tc_commit(bar(BAR_H_EMPTY + s));
        if (t == ntiles - 1) tc_commit(bar(BAR_DW_FULL));
      }
    }
  } else if (warp < 8) {
    // ====================== logits / loss epilogue: thread = (expert jl, half hh of the tile's 64 teams) ======================
    const int jl = threadIdx.x & 127;  // TMEM lane = local expert
    const int hh = threadIdx.x >> 7;   // teams [32*hh, 32*hh+32) of the tile
    const int e = e0 + jl;
    const bool e_ok = e < g.E;
    const float bj = e_ok ? __ldg(g.bias + e) : 0.f;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    if (train) {
      // fp16 image of the W tile for the dA product (MN-major: rows = experts, 64 hidden units per 128-byte row)
      mbar_wait(bar(BAR_W), 0);
#pragma unroll 4
      for (int uu = 0; uu < HK / 16; ++uu) {  // this half's 8 units of 8 hidden values
        const int k = (hh * 8 + uu) * 8;
        const uint8_t* src = sgen + OFF_W32 + (k >> 5) * (TE * 128) + jl * 128;
        const int u32a = ((k & 31) >> 2), u32b = u32a + 1;
        const float4 f0 = *reinterpret_cast<const float4*>(src + ((u32a ^ (jl & 7)) << 4));
        const float4 f1 = *reinterpret_cast<const float4*>(src + ((u32b ^ (jl & 7)) << 4));
        __half2 h0 = __floats2half2_rn(f0.x, f0.y), h1 = __floats2half2_rn(f0.z, f0.w);
        __half2 h2 = __floats2half2_rn(f1.x, f1.y), h3 = __floats2half2_rn(f1.z, f1.w);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        uint8_t* dst = sgen + OFF_W16 + (k >> 6) * (TE * 128) + jl * 128 + ((((k & 63) >> 3) ^ (jl & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = pk;
      }
      fence_proxy_async();
    }
    if (MODE == 0) mbar_arrive(bar(BAR_W16));  // (also counted for validation steps; the MMA thread only waits when training)
    float acc_lin = 0.f, acc_lg = 0.f, loss_sp = 0.f, db_acc = 0.f;  // dense loss = tnw*ln2*(acc_lg - acc_lin): sums of lg2(1+e) and of t = -c*x
    // dz is kept as w*(sigmoid-y)*slope, i.e. true dz / loss_scale; experts past E (last tile) get zero gradient and their loss is dropped below
    const float c_pos = e_ok ? g.tnw : 0.f, c_neg = e_ok ? g.tnw * NTF_LRELU_SLOPE : 0.f;
    const uint32_t* plane_s = reinterpret_cast<const uint32_t*>(sgen + OFF_PLANE);
    const uint32_t* plane_y = plane_s + TE * 2;
    constexpr float LOG2E = 1.4426950408889634f;
    const float kb1 = -LOG2E * bj, kb2 = -LOG2E * NTF_LRELU_SLOPE * bj;
    for (int t = 0; t < ntiles; ++t) {
      const int s = t & 1;
      const uint32_t ph = (t >> 1) & 1;
      const int n0 = t * TB + hh * 32;  // first team of this thread's half tile
      // bit n of S / Y: (team n0+n, my expert) carries weight tpw / target 1 -- two words from the tile's planes, then the stage is free
      uint32_t S = 0, Y = 0;
      if (MODE == 0) {
        mbar_wait(bar(BAR_SP_FULL), t & 1);
        S = plane_s[jl * 2 + hh];
        Y = plane_y[jl * 2 + hh];
        mbar_arrive(bar(BAR_SP_EMPTY));
      }
      mbar_wait(bar(BAR_Z_FULL + s), ph);
      if (g.timing && blockIdx.x == 0 && threadIdx.x == 0) g.timing[t * 8 + 4] = clock64();
      tc_fence_after();
      float z[32];
      tmem_ld32(tmem + lane_base + TM_Z + s * TB + hh * 32, z);
      tc_fence_before();
      mbar_arrive(bar(BAR_Z_EMPTY + s));
      if (g.timing && blockIdx.x == 0 && threadIdx.x == 0) g.timing[t * 8 + 5] = clock64();
      const int nrem = g.B - n0;  // teams of this half tile that exist
      if (MODE == 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float p[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float zz = z[u * 8 + q] + bj;
            const float x = fmaxf(zz, NTF_LRELU_SLOPE * zz);
            const float ex = ex2_approx(fabsf(x) * -1.4426950408889634f);
            const float r = rcp_approx(1.f + ex);
            p[q] = zz > 0.f ? r : ex * r;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (e_ok && u * 8 + q < nrem) g.P[(size_t)(n0 + u * 8 + q) * g.E + e] = p[q];
        }
        if (g.timing && blockIdx.x == 0 && threadIdx.x == 0) g.timing[t * 8 + 6] = clock64();
        continue;
      }
      if (g.Zdbg) {
#pragma unroll
        for (int n = 0; n < 32; ++n)
          if (e_ok && n < nrem) g.Zdbg[(size_t)(n0 + n) * g.E + e] = z[n] + bj;
      }
      if (train) mbar_wait(bar(BAR_DZ_EMPTY + s), ph ^ 1);
      uint8_t* dzrow = sgen + OFF_DZ + s * DZ_BYTES + jl * 128;
      // Dense pass: every element as (target 0, weight tnw):  loss = softplus(x) = x + ln(1 + exp(-x)),  dz = tnw*sigmoid(x)*slope,
      // x = lrelu(z + b).  With c = log2(e):  t = -c*x = min(-c*(z+b), -0.01c*(z+b))  (two FFMAs with the bias folded in, one FMNMX),
      // e = 2^t, den = 1 + e, sigmoid(x) = 1/den, softplus(x) = (-t + lg2(den))*ln2.  lg2 is taken once per 8 elements, of the product
      // of the den factors; sum(t) and sum(lg2) are accumulated separately and combined once per CTA.  8 teams -> one 16-byte unit
      // of the dz^T row; 8 independent chains keep the MUFU pipe fed.  A product that overflows (logits below about -300: den > 2^16
      // each) falls back to per-element logarithms.
      auto dense8 = [&](int u, auto masked) {
        float gz[8], den[8];
        float prod = 1.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float t1 = fmaf(z[u * 8 + q], -LOG2E, kb1);
          const float t2 = fmaf(z[u * 8 + q], -LOG2E * NTF_LRELU_SLOPE, kb2);
          float t = fminf(t1, t2);
          const float ex = ex2_approx(t);
          den[q] = 1.f + ex;
          gz[q] = rcp_approx(den[q]) * (t1 < 0.f ? c_pos : c_neg);
          if (decltype(masked)::value && u * 8 + q >= nrem) { gz[q] = 0.f; t = 0.f; den[q] = 1.f; }  // teams past the end of the batch
          acc_lin += t;
          prod *= den[q];
          db_acc += gz[q];
        }
        if (prod < 3.0e38f) {
          acc_lg += lg2_approx(prod);
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) acc_lg += lg2_approx(den[q]);
        }
        if (train) {
          const __half2 h0 = __floats2half2_rn(gz[0], gz[1]), h1 = __floats2half2_rn(gz[2], gz[3]);
          const __half2 h2 = __floats2half2_rn(gz[4], gz[5]), h3 = __floats2half2_rn(gz[6], gz[7]);
          uint4 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&h0); pk.y = *reinterpret_cast<const uint32_t*>(&h1);
          pk.z = *reinterpret_cast<const uint32_t*>(&h2); pk.w = *reinterpret_cast<const uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(dzrow + (((hh * 4 + u) ^ (jl & 7)) << 4)) = pk;
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
      // contribution is taken back out and the fp16 gradient already staged in shared memory is overwritten.
      while (S) {
        const int i = __ffs(S) - 1;
        S &= S - 1;
        float zi = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) zi = (k == i) ? z[k] : zi;
        asm volatile("" : "+f"(zi));  // keep this path from pinning the dense pass's temporaries in registers
        const float zz = zi + bj;
        const float x = fmaxf(zz, NTF_LRELU_SLOPE * zz);
        const float t = -LOG2E * x;
        const float exs = ex2_approx(t), dens = 1.f + exs;          // the dense pass's (signed) terms, taken back out
        const float g_dense = rcp_approx(dens) * (zz > 0.f ? c_pos : c_neg);
        acc_lin -= t;
        acc_lg -= lg2_approx(dens);
        const float ex = ex2_approx(fabsf(x) * -LOG2E);               // stable form for the real term
        const float den = 1.f + ex, lg = lg2_approx(den), r = rcp_approx(den);
        const float sig = x > 0.f ? r : ex * r;
        const float yf = ((Y >> i) & 1u) ? 1.f : 0.f;
        const float g_sp = g.tpw * (sig - yf) * (zz > 0.f ? 1.f : NTF_LRELU_SLOPE);
        loss_sp += g.tpw * ((1.f - yf) * x + fmaxf(-x, 0.f) + 0.6931471805599453f * lg);
        db_acc += g_sp - g_dense;
        if (train) *reinterpret_cast<__half*>(dzrow + (((hh * 4 + (i >> 3)) ^ (jl & 7)) << 4) + (i & 7) * 2) = __float2half_rn(g_sp);
      }
      if (train) {
        fence_proxy_async();
        mbar_arrive(bar(BAR_DZ_FULL + s));
      }
      if (g.timing && blockIdx.x == 0 && threadIdx.x == 0) g.timing[t * 8 + 6] = clock64();
    }
    if (MODE == 0) {
      // loss partial of this CTA and db: fixed-order combines (shuffle tree, then warps / halves in order)
      float* red = reinterpret_cast<float*>(sgen + BAR_OFF + NUM_BARS * 8 + 16);   // [8]
      float* dbs = reinterpret_cast<float*>(sgen + OFF_DZ);                          // [128], the dz stages are idle by now ...
      if (train) mbar_wait(bar(BAR_DW_FULL), 0);                                     // ... once every MMA that read them has completed
      const float tot = warp_sum(e_ok ? g.tnw * 0.6931471805599453f * (acc_lg - acc_lin) + loss_sp : 0.f);
      if (lane == 0) red[warp] = tot;
      if (train && hh == 1) dbs[jl] = db_acc;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) g.loss_part[blockIdx.x] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
      if (train) {
        if (hh == 0 && e_ok) g.db[e] = (db_acc + dbs[jl]) * g.scale;
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {  // this half's 64 of the 128 dW columns
          float v[32];
          tmem_ld32(tmem + lane_base + TM_DW + hh * 64 + c * 32, v);
          if (e_ok) {
            float4* dst = reinterpret_cast<float4*>(g.dW + (size_t)e * HK + hh * 64 + c * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q] * g.scale, v[4 * q + 1] * g.scale, v[4 * q + 2] * g.scale, v[4 * q + 3] * g.scale);
          }
        }
      }
    }
  } else {
    // =========================================== dA epilogue: thread = hidden unit ===========================================
    if (train) {
      const int k = threadIdx.x - 256;  // 0..127 = TMEM lane = hidden unit
      const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
      mbar_arrive(bar(BAR_W16));  // these warps take the rest of the arrival count (the image is written by warps 0-7)
      for (int t = 0; t < ntiles; ++t) {
        const int s = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        const int n0 = t * TB;
        mbar_wait(bar(BAR_DA_FULL + s), ph);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < TB / 32; ++c) {
          float v[32];
          tmem_ld32(tmem + lane_base + TM_DA + s * TB + c * 32, v);
          if (c == TB / 32 - 1) {
            tc_fence_before();
            mbar_arrive(bar(BAR_DA_EMPTY + s));
          }
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int n = n0 + c * 32 + q;
            if (n < g.B && !(g.exp & 1)) atomicAdd(g.dA + (size_t)n * HK + k, v[q] * g.scale);  // red.global.add.f32, 128 B per warp
          }
        }
      }
    } else if (MODE == 0) {
      mbar_arrive(bar(BAR_W16));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
  }
}

__global__ void to_half_kernel(const float* __restrict__ x, size_t n, __half* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = __float2half_rn(x[i]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(const ntf_ctx* ctx, CUtensorMap* m, CUtensorMapDataType dt, int esize, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
             uint32_t box_cols) {
  NTF_REQUIRE(ctx->encode_tiled != nullptr, NTF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {cols * (uint64_t)esize};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = ((EncodeTiledFn)ctx->encode_tiled)(m, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NTF_REQUIRE(r == CUDA_SUCCESS, NTF_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return NTF_OK;
}
}  // namespace

int ntf_out_tc_supported(int B, int h, int E, int flipout) { return (h == HK && !flipout && B >= 1 && E >= 1) ? 1 : 0; }

size_t ntf_out_train_tc_workspace_bytes(const ntf_ctx*, int B, int h, int E, int) {
  return align_up((size_t)B * h * sizeof(__half), 256) + align_up((size_t)cdiv(E, TE) * sizeof(float), 256);
}

int ntf_out_train_tc(ntf_ctx* ctx, cudaStream_t st, const ntf_out_train_args* a, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(a->h == HK, NTF_ERR_UNSUPPORTED, "out_train(tf32): hidden width %d (kernel is built for %d)", a->h, HK);
  NTF_REQUIRE(workspace_bytes >= ntf_out_train_tc_workspace_bytes(ctx, a->B, a->h, a->E, 0), NTF_ERR_WORKSPACE, "out_train(tf32): workspace too small");
  NTF_REQUIRE((((uintptr_t)a->A | (uintptr_t)a->W) & 15) == 0, NTF_ERR_BAD_ARG, "out_train(tf32): A and W must be 16-byte aligned");
  const bool train = a->dW != nullptr;
  NTF_REQUIRE(!train || (a->db && a->dA), NTF_ERR_BAD_ARG, "out_train(tf32): training needs dW, db and dA");
  __half* A16 = (__half*)workspace;
  float* loss_part = (float*)((char*)workspace + align_up((size_t)a->B * a->h * sizeof(__half), 256));
  const int nct = cdiv(a->E, TE);
  CUtensorMap mw, ma, mh;
  int rc;
  if ((rc = make_map(ctx, &mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->W, (uint64_t)a->E, HK, TE, 32))) return rc;
  if ((rc = make_map(ctx, &ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->A, (uint64_t)a->B, HK, TB, 32))) return rc;
  if ((rc = make_map(ctx, &mh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, A16, (uint64_t)a->B, HK, TB, 64))) return rc;
  if (train) {
    NTF_COUNT_LAUNCH; to_half_kernel<<<min(cdiv(a->B * a->h, 256), ctx->sm_count * 4), 256, 0, st>>>(a->A, (size_t)a->B * a->h, A16);
    NTF_CUDA(cudaMemsetAsync(a->dA, 0, (size_t)a->B * a->h * sizeof(float), st));
  }
  TcArgs g{};
  g.bias = a->b; g.special = a->special; g.pitch = a->pitch_words; g.m_indptr = a->m_indptr; g.m_indices = a->m_indices;
  g.B = a->B; g.E = a->E; g.tpw = a->tpw; g.tnw = a->tnw; g.scale = a->loss_scale;
  g.dW = a->dW; g.db = a->db; g.dA = a->dA; g.loss_part = loss_part; g.P = nullptr;
  const char* dbg = getenv("NTF_TC_ZDBG");  // debug hook used by tests/test_gpu_tc.py: address of a [B,E] device buffer for the raw logits
  g.Zdbg = dbg ? (float*)(uintptr_t)strtoull(dbg, nullptr, 0) : nullptr;
  const char* tim = getenv("NTF_TC_TIMING");
  g.timing = tim ? (long long*)(uintptr_t)strtoull(tim, nullptr, 0) : nullptr;
  const char* ex = getenv("NTF_TC_EXP");
  g.exp = ex ? atoi(ex) : 0;
  NTF_CUDA(cudaFuncSetAttribute(out_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TRAIN));
  NTF_COUNT_LAUNCH; out_tc_kernel<0><<<nct, NT, SMEM_TRAIN, st>>>(mw, ma, mh, g);
  NTF_LAUNCH_CHECK();
  return ntf_loss_reduce_impl(st, loss_part, nct, a->loss_scale, a->loss_out);
}

int ntf_infer_scores_tc(ntf_ctx* ctx, cudaStream_t st, const float* A, const float* W, const float* b, int B, int h, int E, float* P) {
  NTF_REQUIRE(h == HK, NTF_ERR_UNSUPPORTED, "infer_scores(tf32): hidden width %d (kernel is built for %d)", h, HK);
  CUtensorMap mw, ma;
  int rc;
  if ((rc = make_map(ctx, &mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, W, (uint64_t)E, HK, TE, 32))) return rc;
  if ((rc = make_map(ctx, &ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, A, (uint64_t)B, HK, TB, 32))) return rc;
  TcArgs g{};
  g.bias = b; g.B = B; g.E = E; g.P = P;
  const char* tim = getenv("NTF_TC_TIMING");
  g.timing = tim ? (long long*)(uintptr_t)strtoull(tim, nullptr, 0) : nullptr;
  NTF_CUDA(cudaFuncSetAttribute(out_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_INFER));
  NTF_COUNT_LAUNCH; out_tc_kernel<1><<<cdiv(E, TE), NT, SMEM_INFER, st>>>(mw, ma, ma, g);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
