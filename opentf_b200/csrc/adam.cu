// adam.cu -- K6: torch.optim.Adam(betas=(.9,.999), eps=1e-8, weight_decay=0, amsgrad=False) (fnn.py:104,139) as ONE
// launch over the flat parameter arena.  HBM-bound: 28 bytes per parameter per step (read p,g,m,v; write p,m,v).
// Operation order follows torch's single-tensor path: m.lerp_(g,1-b1); v.mul_(b2).addcmul_(g,g,1-b2);
// denom = sqrt(v)/sqrt(bc2) + eps; p.addcdiv_(m, denom, -lr/bc1).
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace {

// `shadow` (nullable): fp16 image of the parameters [sh_lo4, sh_hi4) (float4 units of this call's range), rewritten with the stepped values --
// the operand image the tensor-core output layer TMA-loads (out_tc2.cu), kept current without a conversion pass of its own
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, size_t n4, size_t n, AdamK k, const ntf_dyn* __restrict__ dyn,
                                                   uint2* __restrict__ shadow, size_t sh_lo4, size_t sh_hi4) {
  if (dyn) k = AdamK{dyn->one_minus_b1, dyn->b2, dyn->one_minus_b2, dyn->bc2_sqrt, dyn->eps, dyn->neg_step};  // a replayed graph: this step's constants
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
    adam1(P.x, G.x, M.x, V.x, k); adam1(P.y, G.y, M.y, V.y, k); adam1(P.z, G.z, M.z, V.z, k); adam1(P.w, G.w, M.w, V.w, k);
    reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
    if (shadow && i >= sh_lo4 && i < sh_hi4) {
      const __half2 lo = __floats2half2_rn(P.x, P.y), hi = __floats2half2_rn(P.z, P.w);
      shadow[i - sh_lo4] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) adam1(p[i], g[i], m[i], v[i], k);
}
}  // namespace

// `dyn` (device, nullable): the constants come from the block ntf_dyn_update wrote for this step instead of the arguments
int ntf_adam_step_impl(ntf_ctx* ctx, cudaStream_t st, float* p, const float* g, float* m, float* v, size_t n, double lr, double beta1,
                       double beta2, double eps, int64_t step, const ntf_dyn* dyn, void* shadow, size_t sh_off, size_t sh_n) {
  NTF_REQUIRE(ctx && p && g && m && v, NTF_ERR_BAD_ARG, "adam_step: null pointer");
  NTF_REQUIRE(dyn || step >= 1, NTF_ERR_BAD_ARG, "adam_step: step=%lld (1-based)", (long long)step);
  if (n == 0) return NTF_OK;
  const bool aligned = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16) == 0;
  const AdamK k = adam_consts(lr, beta1, beta2, eps, step >= 1 ? step : 1);
  const size_t n4 = aligned ? n / 4 : 0;
  NTF_REQUIRE(!shadow || (aligned && (sh_off % 4) == 0 && (sh_n % 4) == 0 && sh_off + sh_n <= n4 * 4 && ((uintptr_t)shadow & 7) == 0), NTF_ERR_BAD_ARG,
              "adam_step: fp16 shadow range [%zu,+%zu) must be 4-float aligned inside the stepped range", sh_off, sh_n);
  const size_t work = n4 ? n4 : n;
  const int blocks = (int)((work + 255) / 256 < (size_t)ctx->sm_count * 16 ? (work + 255) / 256 : (size_t)ctx->sm_count * 16);
  NTF_COUNT_LAUNCH; adam_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n4, n, k, dyn, (uint2*)shadow, sh_off / 4, (sh_off + sh_n) / 4);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

extern "C" int ntf_adam_step(ntf_ctx* ctx, void* stream, float* p, const float* g, float* m, float* v, size_t n, double lr,
                             double beta1, double beta2, double eps, int64_t step) {
  return ntf_adam_step_impl(ctx, as_stream(stream), p, g, m, v, n, lr, beta1, beta2, eps, step, ctx ? ctx->dyn_override : nullptr, nullptr, 0, 0);
}

namespace {
__global__ void dyn_update_kernel(ntf_dyn* dyn, ntf_dyn v) { *dyn = v; }
}  // namespace

extern "C" int ntf_dyn_update(ntf_ctx* ctx, void* stream, ntf_dyn* dyn, uint64_t step, double lr, double beta1, double beta2, double eps,
                              int64_t adam_t) {
  NTF_REQUIRE(ctx && dyn, NTF_ERR_BAD_ARG, "dyn_update: null pointer");
  const AdamK k = adam_consts(lr, beta1, beta2, eps, adam_t >= 1 ? adam_t : 1);
  ntf_dyn v;
  v.step = step;
  v.one_minus_b1 = k.one_minus_b1; v.b2 = k.b2; v.one_minus_b2 = k.one_minus_b2; v.bc2_sqrt = k.bc2_sqrt; v.eps = k.eps; v.neg_step = k.neg_step;
  NTF_COUNT_LAUNCH; dyn_update_kernel<<<1, 1, 0, as_stream(stream)>>>(dyn, v);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
