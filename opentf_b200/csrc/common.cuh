// common.cuh -- shared helpers for libntf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ntf_b200.h"

struct ntf_ctx {
  int device;
  int sm_count;
  int cc_major, cc_minor;
  size_t smem_optin;
  void* encode_tiled;  // cuTensorMapEncodeTiled, resolved through cudaGetDriverEntryPoint
  // ntf_fnn_step forks the parts of a step that depend on the batch's CSR only (negative sampling + condition planes; the slot
  // fill of the input layer's backward pass) onto side streams and joins them where their results are needed; inside a stream
  // capture the same calls become parallel branches of the graph
  cudaStream_t side[2];
  cudaEvent_t ev_fork, ev_join[2];
  cudaEvent_t ev_fork_opt, ev_join_opt;  // second fork of a step: Adam on the output layer's segment next to the input layer's backward pass
  // data-parallel ranks (ntf_fnn_step_args.comm): the gradient all-reduces run on their own stream so that the output layer's segment
  // is exchanged while the backward pass through the hidden layers runs, and stepped while the rest is exchanged
  const ntf_dyn* dyn_override;  // ntf_set_dyn: while set, the entry points that take step / lr / adam_t as host arguments enqueue kernels that read them from this device block
  cudaEvent_t ev_hot_fork, ev_hot_join;  // input layer's backward: the hot skills' kernels next to the warp-per-skill reduction
  cudaStream_t comm_st;
  cudaEvent_t ev_ar[2], ev_bwd;
  cudaStream_t copy_st, copy_dn;  // the streaming entry point's copies, one stream per direction (ntf_host_batch_upload / ntf_host_loss_download)
  cudaEvent_t ev_up[2], ev_step[2], ev_loss[2];
  cudaEvent_t ev_finish;  // the output layer's deferred reductions (loss, fused db) done on side[0]: the optimiser of the hidden layers' segment waits for it
};

// ---- launch accounting (bench.py reports how many of OUR kernels ran in the timed region) ----------------
extern unsigned long long g_ntf_launches;
#define NTF_COUNT_LAUNCH (++g_ntf_launches)

// ---- error plumbing -------------------------------------------------------------------------------------
void ntf_set_error(const char* fmt, ...);

#define NTF_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ntf_set_error(__VA_ARGS__);     \
      return (code);                  \
    }                                 \
  } while (0)

#define NTF_CUDA(call)                                                                            \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      ntf_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return NTF_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

#define NTF_LAUNCH_CHECK() NTF_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- device helpers -------------------------------------------------------------------------------------
#define NTF_LRELU_SLOPE 0.01f

__device__ __forceinline__ float lrelu(float z) { return z > 0.f ? z : NTF_LRELU_SLOPE * z; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Weighted BCE-with-logits on x = lrelu(z), and its gradient wrt z.
//   loss = w*((1-y)*x + max(-x,0) + log1p(exp(-|x|)))      d/dz = w*(sigmoid(x)-y)*scale*(z>0 ? 1 : 0.01)
// FAST=false: libm-accurate expf/log1pf (the fp32 parity mode).  FAST=true: MUFU ex2/lg2/rcp.
template <bool FAST>
__device__ __forceinline__ void bce_elem(float z, bool y, float w, float scale, float& loss, float& dz) {
  const float x = lrelu(z);
  const float ax = fabsf(x);
  float e, l1p, r;
  if (FAST) {
    e = __expf(-ax);
    l1p = __logf(1.f + e);
    r = __frcp_rn(1.f + e);
  } else {
    e = expf(-ax);
    l1p = log1pf(e);
    r = 1.f / (1.f + e);
  }
  const float sig = x >= 0.f ? r : e * r;
  const float yf = y ? 1.f : 0.f;
  loss = w * ((1.f - yf) * x + fmaxf(-x, 0.f) + l1p);
  dz = w * (sig - yf) * scale * (z > 0.f ? 1.f : NTF_LRELU_SLOPE);
}

template <bool FAST>
__device__ __forceinline__ float sigmoid_lrelu(float z) {
  const float x = lrelu(z);
  if (FAST) {
    const float e = __expf(-fabsf(x));
    const float r = __frcp_rn(1.f + e);
    return x >= 0.f ? r : e * r;
  }
  return 1.f / (1.f + expf(-x));  // torch.sigmoid's fp32 formula
}

// torch.optim.Adam's single-tensor update (fnn.py:104,139), one element:  m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2);
// denom = sqrt(v)/sqrt(bc2) + eps; p.addcdiv_(m, denom, -lr/bc1).  Every operation rounded on its own (no fused multiply-add), written
// with explicit-rounding intrinsics so that every kernel
// that steps parameters (adam.cu, peer.cu) performs the SAME operations -- the compiler may not re-associate or fuse them differently
// from one kernel to the next, and data-parallel replicas / single-GPU runs stay bit-comparable.
struct AdamK { float one_minus_b1, b2, one_minus_b2, bc2_sqrt, eps, neg_step; };
static inline AdamK adam_consts(double lr, double beta1, double beta2, double eps, int64_t step) {
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  return AdamK{(float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)sqrt(bc2), (float)eps, (float)(-(lr / bc1))};
}
__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamK& k) {
  m = __fadd_rn(m, __fmul_rn(k.one_minus_b1, __fsub_rn(g, m)));
  v = __fadd_rn(__fmul_rn(v, k.b2), __fmul_rn(__fmul_rn(k.one_minus_b2, g), g));
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), k.bc2_sqrt), k.eps);
  p = __fadd_rn(p, __fmul_rn(k.neg_step, __fdiv_rn(m, denom)));
}

// is expert j a member of team `row`?  (member rows are short: 2-10 entries, sorted)
__device__ __forceinline__ bool is_member(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                          int row, int j) {
  const int b = indptr[row], e = indptr[row + 1];
  for (int p = b; p < e; ++p)
    if (indices[p] == j) return true;
  return false;
}

// ---- Philox4x32-10 (Salmon et al. 2011), the counter RNG restated in oracle/sampler_oracle.py -------------
struct Philox4 {
  uint32_t v[4];
};
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}
