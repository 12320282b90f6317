// dense_simt.cu -- NTF_FP32 mode: CUDA-core fp32 FMA-chain GEMMs with fused epilogues.
//
// This is the bit-stable parity path (every dot product is one ascending-k fmaf chain, no split
// accumulators inside a tile) and the on-device cross-check for the tcgen05 kernels in out_tc.cu.  It serves
//   - hidden layers fwd/bwd (fnn.py:25 layers i>=1 and their autograd),
//   - the output layer training step in fp32 (forward+loss+dz, then dW, db, dA as three more contractions),
//   - inference scores in fp32.
// One generic kernel: C[m,n] = sum_k A(m,k)*B(n,k) with arbitrary element strides, 128x128x16 tiles,
// 256 threads x (8x8) outputs, optional split-K over blockIdx.z, epilogue as a functor.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

struct GemmArgs {
  int M, N, K;
  const float* A; long long a_sm, a_sk;
  const float* B; long long b_sn, b_sk;
  int kchunk;  // K range per blockIdx.z
};

template <class Epi>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g, Epi epi) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * g.kchunk, k_end = min(g.K, k_begin + g.kchunk);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const bool a_kcontig = (g.a_sk == 1), b_kcontig = (g.b_sk == 1);
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // ---- stage tiles (zero-filled outside the matrix) ----
#pragma unroll
    for (int i = 0; i < (BM * BK) / 256; ++i) {
      const int idx = t + 256 * i;
      int kk, mm;
      if (a_kcontig) { kk = idx % BK; mm = idx / BK; } else { mm = idx % BM; kk = idx / BM; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < g.M && k < k_end) ? __ldg(g.A + (long long)m * g.a_sm + (long long)k * g.a_sk) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / 256; ++i) {
      const int idx = t + 256 * i;
      int kk, nn;
      if (b_kcontig) { kk = idx % BK; nn = idx / BK; } else { nn = idx % BN; kk = idx / BN; }
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < g.N && k < k_end) ? __ldg(g.B + (long long)n * g.b_sn + (long long)k * g.b_sk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4 + 64]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue: rows ty*8+i, columns tx*4+q and 64+tx*4+q ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m < g.M) {
      const float v0[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
      const float v1[4] = {acc[i][4], acc[i][5], acc[i][6], acc[i][7]};
      epi.apply(m, n0 + tx * 4, v0, g.N);
      epi.apply(m, n0 + tx * 4 + 64, v1, g.N);
    }
  }
  epi.finish();
}

// ---- epilogues ------------------------------------------------------------------------------------------
__device__ __forceinline__ float sign_of(const uint32_t* __restrict__ bits, int pitch, int m, int n) {
  return ((__ldg(bits + (size_t)m * pitch + (n >> 5)) >> (n & 31)) & 1u) ? -1.f : 1.f;
}

// C[z][m][n] = acc  (split-K partials or plain products)
struct EpiStore {
  float* C; long long ldc; long long part_stride;
  __device__ __forceinline__ void apply(int m, int n, const float (&v)[4], int N) {
    float* dst = C + (long long)blockIdx.z * part_stride + (long long)m * ldc + n;
#pragma unroll
    for (int q = 0; q < 4; ++q) if (n + q < N) dst[q] = v[q];
  }
  __device__ __forceinline__ void finish() {}
};

// Y = act( (acc + bias[n]) * sign(m,n) + addend[m,n] ), optionally accumulated into Y
//   act 0: identity, 1: lrelu, 2: sigmoid(lrelu(.))
struct EpiBiasAct {
  float* Y; long long ldy; const float* bias; const uint32_t* sign; int sign_pitch; const float* addend; int act; int accumulate;
  __device__ __forceinline__ void apply(int m, int n, const float (&v)[4], int N) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (n + q >= N) continue;
      float z = v[q] + (bias ? __ldg(bias + n + q) : 0.f);
      if (sign) z *= sign_of(sign, sign_pitch, m, n + q);
      if (addend) z += addend[(long long)m * ldy + n + q];
      float o = act == 0 ? z : (act == 1 ? lrelu(z) : sigmoid_lrelu<false>(z));
      float* dst = Y + (long long)m * ldy + n + q;
      *dst = accumulate ? *dst + o : o;
    }
  }
  __device__ __forceinline__ void finish() {}
};

// output layer forward + weighted BCE + dz  (fnn.py:25,32-46,135 and the first step of its autograd)
struct EpiLoss {
  const float* bias; const float* addend; long long ld;  // addend: Flipout perturbation term T[m,n] (nullable)
  const uint32_t* special; int pitch; const int32_t* m_indptr; const int32_t* m_indices; int e_lo;
  float tpw, tnw, scale;
  float* dz;                                  // [B,E] or NULL (validation step)
  const uint32_t* sign_out; float* dzs;       // Flipout: dzs = dz * s_out
  float* part;                                // one partial loss per CTA
  float loss_acc;
  __device__ __forceinline__ void apply(int m, int n, const float (&v)[4], int N) {
    uint32_t word = 0u;
    if (special) word = __ldg(special + (size_t)m * pitch + (n >> 5)) >> (n & 31);  // n%4==0: 4 bits of one word
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (n + q >= N) continue;
      float z = v[q] + __ldg(bias + n + q);
      if (addend) z += addend[(long long)m * ld + n + q];
      const bool sp = (word >> q) & 1u;
      const bool y = sp && is_member(m_indptr, m_indices, m, n + q + e_lo);
      float l, g;
      bce_elem<false>(z, y, sp ? tpw : tnw, scale, l, g);
      loss_acc += l;
      if (dz) dz[(long long)m * ld + n + q] = g;
      if (dzs) dzs[(long long)m * ld + n + q] = g * sign_of(sign_out, pitch, m, n + q);
    }
  }
  __device__ __forceinline__ void finish() {
    __shared__ float red[8];
    float s = warp_sum(loss_acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += red[w];
      part[blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
  }
};

template <class Epi>
int launch_gemm(cudaStream_t st, const GemmArgs& g, const Epi& epi, int nsplit = 1) {
  dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), nsplit);
  NTF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, NTF_ERR_UNSUPPORTED, "gemm grid too large (%u,%u,%u)", grid.x, grid.y, grid.z);
  NTF_COUNT_LAUNCH; sgemm_kernel<Epi><<<grid, 256, 0, st>>>(g, epi);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

// out[i] = sum_k parts[k*stride + i], fixed order
__global__ void sum_parts_kernel(const float* __restrict__ parts, int nparts, size_t n, size_t stride, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < nparts; ++k) s += parts[(size_t)k * stride + i];
  out[i] = s;
}

// single-block fixed-order reduction of per-CTA loss partials (double accumulator): loss_out[0] = scale * sum
__global__ void loss_reduce_kernel(const float* __restrict__ part, int n, float scale, float* __restrict__ loss_out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = (float)(red[0] * (double)scale);
}

// column sums of a [R,C] matrix in two fixed-order passes
constexpr int CS_ROWS = 64;
__global__ void colsum_pass1(const float* __restrict__ X, int R, int C, float* __restrict__ part) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int r0 = blockIdx.y * CS_ROWS, r1 = min(R, r0 + CS_ROWS);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += X[(size_t)r * C + c];
  part[(size_t)blockIdx.y * C + c] = s;
}
}  // namespace

int ntf_sum_parts_impl(cudaStream_t st, const float* parts, int nparts, size_t n, size_t stride, float* out) {
  NTF_COUNT_LAUNCH; sum_parts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(parts, nparts, n, stride, out);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
int ntf_loss_reduce_impl(cudaStream_t st, const float* part, int n, float scale, float* loss_out) {
  NTF_COUNT_LAUNCH; loss_reduce_kernel<<<1, 256, 0, st>>>(part, n, scale, loss_out);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
int ntf_colsum_impl(cudaStream_t st, const float* X, int R, int C, float* out, float* part_ws) {
  const int np = cdiv(R, CS_ROWS);
  NTF_REQUIRE(np <= 65535, NTF_ERR_UNSUPPORTED, "colsum: too many rows %d", R);
  NTF_COUNT_LAUNCH; colsum_pass1<<<dim3(cdiv(C, 128), np), 128, 0, st>>>(X, R, C, part_ws);
  NTF_COUNT_LAUNCH; sum_parts_kernel<<<cdiv(C, 256), 256, 0, st>>>(part_ws, np, (size_t)C, (size_t)C, out);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

extern "C" int ntf_sum_parts(ntf_ctx* ctx, void* stream, const float* parts, int nparts, size_t n, size_t part_stride, float* out) {
  NTF_REQUIRE(ctx && parts && out && nparts > 0, NTF_ERR_BAD_ARG, "sum_parts: bad argument");
  return ntf_sum_parts_impl(as_stream(stream), parts, nparts, n, part_stride, out);
}

// =========================================================================================================
// hidden layers
// =========================================================================================================
extern "C" int ntf_dense_fwd(ntf_ctx* ctx, void* stream, const float* A, const float* W, const float* b, int B, int in,
                             int out, int act, float* Y) {
  NTF_REQUIRE(ctx && A && W && Y, NTF_ERR_BAD_ARG, "dense_fwd: null pointer");
  NTF_REQUIRE(B > 0 && in > 0 && out > 0 && act >= 0 && act <= 2, NTF_ERR_BAD_ARG, "dense_fwd: B=%d in=%d out=%d act=%d", B, in, out, act);
  GemmArgs g{B, out, in, A, in, 1, W, in, 1, in};
  EpiBiasAct e{Y, out, b, nullptr, 0, nullptr, act, 0};
  return launch_gemm(as_stream(stream), g, e);
}

// Flipout hidden layer (bayesian-torch LinearFlipout):  Y = act( A W^T + b + (A_s W_delta^T + b_delta) * s_out )
extern "C" size_t ntf_dense_flipout_fwd_workspace_bytes(int B, int out) { return align_up((size_t)B * out * sizeof(float), 256); }

extern "C" int ntf_dense_flipout_fwd(ntf_ctx* ctx, void* stream, const float* A, const float* W, const float* b, const float* A_s,
                                     const float* W_delta, const float* b_delta, const uint32_t* sign_out, int pitch_words, int B, int in,
                                     int out, int act, float* Y, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ctx && A && W && b && A_s && W_delta && b_delta && sign_out && Y && workspace, NTF_ERR_BAD_ARG, "dense_flipout_fwd: null pointer");
  NTF_REQUIRE(B > 0 && in > 0 && out > 0 && act >= 0 && act <= 2 && pitch_words * 32 >= out, NTF_ERR_BAD_ARG,
              "dense_flipout_fwd: B=%d in=%d out=%d act=%d pitch=%d", B, in, out, act, pitch_words);
  NTF_REQUIRE(workspace_bytes >= ntf_dense_flipout_fwd_workspace_bytes(B, out), NTF_ERR_WORKSPACE, "dense_flipout_fwd: workspace too small");
  float* T = (float*)workspace;
  int rc;
  GemmArgs g1{B, out, in, A_s, in, 1, W_delta, in, 1, in};
  EpiBiasAct e1{T, out, b_delta, sign_out, pitch_words, nullptr, 0, 0};
  if ((rc = launch_gemm(as_stream(stream), g1, e1))) return rc;
  GemmArgs g2{B, out, in, A, in, 1, W, in, 1, in};
  EpiBiasAct e2{Y, out, b, nullptr, 0, T, act, 0};
  return launch_gemm(as_stream(stream), g2, e2);
}

static int dense_bwd_nsplit(int B) { return B > 1024 ? cdiv(B, 512) : 1; }

extern "C" size_t ntf_dense_bwd_workspace_bytes(int B, int in, int out) {
  const int ns = dense_bwd_nsplit(B);
  return ns > 1 ? align_up((size_t)ns * in * out * sizeof(float), 256) : 0;
}

extern "C" int ntf_dense_bwd(ntf_ctx* ctx, void* stream, const float* A, const float* W, const float* dZ, int B, int in,
                             int out, float* dW, float* dA, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ctx && A && W && dZ && dW, NTF_ERR_BAD_ARG, "dense_bwd: null pointer");
  NTF_REQUIRE(B > 0 && in > 0 && out > 0, NTF_ERR_BAD_ARG, "dense_bwd: B=%d in=%d out=%d", B, in, out);
  NTF_REQUIRE(workspace_bytes >= ntf_dense_bwd_workspace_bytes(B, in, out), NTF_ERR_WORKSPACE, "dense_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int ns = dense_bwd_nsplit(B);
  // dW[o,i] = sum_r dZ[r,o] * A[r,i]
  GemmArgs g{out, in, B, dZ, 1, out, A, 1, in, ns > 1 ? 512 : B};
  int rc;
  if (ns > 1) {
    EpiStore e{(float*)workspace, in, (long long)in * out};
    if ((rc = launch_gemm(st, g, e, ns))) return rc;
    if ((rc = ntf_sum_parts_impl(st, (const float*)workspace, ns, (size_t)in * out, (size_t)in * out, dW))) return rc;
  } else {
    EpiStore e{dW, in, 0};
    if ((rc = launch_gemm(st, g, e))) return rc;
  }
  if (dA) {  // dA[r,i] = sum_o dZ[r,o] * W[o,i]
    GemmArgs g2{B, in, out, dZ, out, 1, W, 1, in, out};
    EpiStore e2{dA, in, 0};
    if ((rc = launch_gemm(st, g2, e2))) return rc;
  }
  return NTF_OK;
}

// =========================================================================================================
// output layer, NTF_FP32
// =========================================================================================================
static int out_dA_nsplit(int E) { return E > 4096 ? cdiv(E, 2048) : 1; }

struct OutWs {
  float *dz, *dzs, *T, *loss_part, *cs_part, *dA_part;
  size_t total;
};
static OutWs out_ws_layout(char* base, int B, int h, int E, bool flip, bool train) {
  OutWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes, 256); return (float*)p; };
  const size_t BE = (size_t)B * E * sizeof(float);
  w.dz = take(BE);
  w.dzs = flip ? take(BE) : nullptr;
  w.T = flip ? take(BE) : nullptr;
  w.loss_part = take((size_t)cdiv(B, BM) * cdiv(E, BN) * sizeof(float));
  w.cs_part = take((size_t)cdiv(B, CS_ROWS) * E * sizeof(float));
  w.dA_part = take((size_t)out_dA_nsplit(E) * B * h * sizeof(float));
  w.total = off;
  (void)train;
  return w;
}

size_t ntf_out_train_fp32_workspace_bytes(int B, int h, int E, int flipout) {
  return out_ws_layout(nullptr, B, h, E, flipout != 0, true).total;
}

int ntf_out_train_fp32(ntf_ctx* ctx, cudaStream_t st, const ntf_out_train_args* a, void* workspace, size_t workspace_bytes) {
  const bool flip = a->W_delta != nullptr;
  const bool train = a->dW != nullptr;
  NTF_REQUIRE(workspace_bytes >= ntf_out_train_fp32_workspace_bytes(a->B, a->h, a->E, flip), NTF_ERR_WORKSPACE,
              "out_train(fp32): workspace %zu < %zu", workspace_bytes, ntf_out_train_fp32_workspace_bytes(a->B, a->h, a->E, flip));
  OutWs w = out_ws_layout((char*)workspace, a->B, a->h, a->E, flip, train);
  const int B = a->B, h = a->h, E = a->E;
  int rc;
  if (flip) {  // T = (A_s dW^T + b_delta) * s_out
    GemmArgs g{B, E, h, a->A_s, h, 1, a->W_delta, h, 1, h};
    EpiBiasAct e{w.T, E, a->b_delta, a->sign_out, a->pitch_words, nullptr, 0, 0};
    if ((rc = launch_gemm(st, g, e))) return rc;
  }
  {
    GemmArgs g{B, E, h, a->A, h, 1, a->W, h, 1, h};
    EpiLoss e{a->b, flip ? w.T : nullptr, E, a->special, a->pitch_words, a->m_indptr, a->m_indices, a->e_lo, a->tpw, a->tnw,
              a->loss_scale, train ? w.dz : nullptr, a->sign_out, (train && flip) ? w.dzs : nullptr, w.loss_part, 0.f};
    if ((rc = launch_gemm(st, g, e))) return rc;
    if ((rc = ntf_loss_reduce_impl(st, w.loss_part, cdiv(B, BM) * cdiv(E, BN), a->loss_scale, a->loss_out))) return rc;
  }
  if (!train) return NTF_OK;
  auto grads = [&](const float* dz, const float* Ain, const float* Wmat, float* dW, float* db, float* dA) -> int {
    int r;
    {  // dW[e,c] = sum_n dz[n,e] * Ain[n,c]
      GemmArgs g{E, h, B, dz, 1, E, Ain, 1, h, B};
      EpiStore e{dW, h, 0};
      if ((r = launch_gemm(st, g, e))) return r;
    }
    if ((r = ntf_colsum_impl(st, dz, B, E, db, w.cs_part))) return r;
    if (dA) {  // dA[n,c] = sum_e dz[n,e] * Wmat[e,c]   (split over E, fixed-order combine)
      const int ns = out_dA_nsplit(E);
      GemmArgs g{B, h, E, dz, E, 1, Wmat, 1, h, ns > 1 ? 2048 : E};
      if (ns > 1) {
        EpiStore e{w.dA_part, h, (long long)B * h};
        if ((r = launch_gemm(st, g, e, ns))) return r;
        if ((r = ntf_sum_parts_impl(st, w.dA_part, ns, (size_t)B * h, (size_t)B * h, dA))) return r;
      } else {
        EpiStore e{dA, h, 0};
        if ((r = launch_gemm(st, g, e))) return r;
      }
    }
    return NTF_OK;
  };
  if ((rc = grads(w.dz, a->A, a->W, a->dW, a->db, a->dA))) return rc;
  if (flip && (rc = grads(w.dzs, a->A_s, a->W_delta, a->dW_delta, a->db_delta, a->dA_s))) return rc;
  (void)ctx;
  return NTF_OK;
}

// =========================================================================================================
// inference scores, NTF_FP32
// =========================================================================================================
int ntf_infer_scores_fp32(cudaStream_t st, const float* A, const float* W, const float* b, int B, int h, int E,
                          const float* A_s, const float* W_delta, const float* b_delta, const uint32_t* sign_out,
                          int pitch_words, int accumulate, float* P, float* T_ws) {
  int rc;
  const bool flip = W_delta != nullptr;
  if (flip) {
    GemmArgs g{B, E, h, A_s, h, 1, W_delta, h, 1, h};
    EpiBiasAct e{T_ws, E, b_delta, sign_out, pitch_words, nullptr, 0, 0};
    if ((rc = launch_gemm(st, g, e))) return rc;
  }
  GemmArgs g{B, E, h, A, h, 1, W, h, 1, h};
  // note: EpiBiasAct reads addend with the same leading dimension as Y
  EpiBiasAct e{P, E, b, nullptr, 0, flip ? T_ws : nullptr, 2, accumulate};
  return launch_gemm(st, g, e);
}
