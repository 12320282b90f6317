// flipout.cu -- Bnn parameter pass (bayesian-torch 0.5.0 LinearFlipout + get_kl_loss; bnn.py:19-25, fnn.py:136,149).
// sigma = softplus(rho); delta = sigma*eps is the weight perturbation shared by the batch; KL per tensor is a MEAN.
#include "common.cuh"

namespace {
constexpr uint32_t PURPOSE_EPS = 0x65707330u, PURPOSE_SIGN = 0x73676e30u;

__device__ __forceinline__ float softplus_f(float r) { return log1pf(expf(r)); }

__device__ __forceinline__ void normal4(uint64_t seed, uint64_t step, uint32_t stream_id, uint32_t blk, float (&z)[4]) {
  const Philox4 r = philox4x32_10(blk, stream_id, (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)seed,
                                  (uint32_t)(seed >> 32) ^ PURPOSE_EPS);
  // Box-Muller on two pairs; u in (0,1]
  const float u0 = ((r.v[0] >> 8) + 1u) * (1.0f / 16777216.0f), u1 = (r.v[1] >> 8) * (1.0f / 16777216.0f);
  const float u2 = ((r.v[2] >> 8) + 1u) * (1.0f / 16777216.0f), u3 = (r.v[3] >> 8) * (1.0f / 16777216.0f);
  const float ra = sqrtf(-2.f * logf(u0)), rb = sqrtf(-2.f * logf(u2));
  float s, c;
  sincospif(2.f * u1, &s, &c); z[0] = ra * c; z[1] = ra * s;
  sincospif(2.f * u3, &s, &c); z[2] = rb * c; z[3] = rb * s;
}

__global__ void fill_normal_kernel(uint64_t seed, uint64_t step, uint32_t stream_id, size_t n, float* __restrict__ out, const ntf_dyn* __restrict__ dyn) {
  if (dyn) step = dyn->step;  // a replayed graph: this step's counter
  const size_t nb = (n + 3) / 4;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (size_t)gridDim.x * blockDim.x) {
    float z[4];
    normal4(seed, step, stream_id, (uint32_t)b, z);
#pragma unroll
    for (int q = 0; q < 4; ++q) if (b * 4 + q < n) out[b * 4 + q] = z[q];
  }
}

__global__ void fill_sign_bits_kernel(uint64_t seed, uint64_t step, uint32_t stream_id, size_t n_words, uint32_t* __restrict__ bits,
                                      const ntf_dyn* __restrict__ dyn) {
  if (dyn) step = dyn->step;
  const size_t nb = (n_words + 3) / 4;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (size_t)gridDim.x * blockDim.x) {
    const Philox4 r = philox4x32_10((uint32_t)b, stream_id, (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)seed,
                                    (uint32_t)(seed >> 32) ^ PURPOSE_SIGN);
#pragma unroll
    for (int q = 0; q < 4; ++q) if (b * 4 + q < n_words) bits[b * 4 + q] = r.v[q];
  }
}

__global__ void apply_sign_kernel(const float* __restrict__ A, const uint32_t* __restrict__ bits, int pitch, int B, int h,
                                  float* __restrict__ As) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * h) return;
  const int n = (int)(i / h), c = (int)(i % h);
  const bool neg = (bits[(size_t)n * pitch + (c >> 5)] >> (c & 31)) & 1u;
  As[i] = neg ? -A[i] : A[i];
}

// Y[n,c] (+)= X[n,c] * sign(n,c)
__global__ void add_signed_kernel(const float* __restrict__ X, const uint32_t* __restrict__ bits, int pitch, int B, int h, int accumulate,
                                  float* __restrict__ Y) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * h) return;
  const int n = (int)(i / h), c = (int)(i % h);
  const bool neg = (bits[(size_t)n * pitch + (c >> 5)] >> (c & 31)) & 1u;
  const float v = neg ? -X[i] : X[i];
  Y[i] = accumulate ? Y[i] + v : v;
}

// delta = softplus(rho)*eps ; per-block KL partial -> part[blockIdx.x]
__global__ void __launch_bounds__(256) flipout_prepare_kernel(const float* __restrict__ mu, const float* __restrict__ rho,
                                                              const float* __restrict__ eps, size_t n, float* __restrict__ delta,
                                                              float* __restrict__ part) {
  float kl = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float sg = softplus_f(rho[i]), m = mu[i];
    delta[i] = sg * eps[i];
    kl += -logf(sg) + 0.5f * (sg * sg + m * m) - 0.5f;  // prior N(0,1): log(1) - log(sg) + (sg^2 + mu^2)/2 - 1/2
  }
  __shared__ float red[8];
  kl = warp_sum(kl);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = kl;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    part[blockIdx.x] = t;
  }
}

__global__ void kl_finish_kernel(const float* __restrict__ part, int n, float kl_scale, float* __restrict__ kl_out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) kl_out[0] += (float)(red[0] * (double)kl_scale);
}

__global__ void flipout_grads_kernel(const float* __restrict__ mu, const float* __restrict__ rho, const float* __restrict__ eps,
                                     const float* __restrict__ g_delta, size_t n, float kl_gscale, float* __restrict__ g_mu,
                                     float* __restrict__ g_rho) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float r = rho[i], sg = softplus_f(r);
    const float dsig = 1.f / (1.f + expf(-r));  // d softplus / d rho
    g_mu[i] += kl_gscale * mu[i];
    g_rho[i] = (g_delta[i] * eps[i] + kl_gscale * (sg - 1.f / sg)) * dsig;
  }
}
}  // namespace

static int grid_for(const ntf_ctx* ctx, size_t n) {
  size_t b = (n + 255) / 256, cap = (size_t)ctx->sm_count * 8;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

extern "C" int ntf_fill_normal(ntf_ctx* ctx, void* stream, uint64_t seed, uint64_t step, uint32_t stream_id, size_t n, float* out) {
  NTF_REQUIRE(ctx && out, NTF_ERR_BAD_ARG, "fill_normal: null pointer");
  if (!n) return NTF_OK;
  NTF_COUNT_LAUNCH; fill_normal_kernel<<<grid_for(ctx, (n + 3) / 4), 256, 0, as_stream(stream)>>>(seed, step, stream_id, n, out, ctx->dyn_override);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

extern "C" int ntf_fill_sign_bits(ntf_ctx* ctx, void* stream, uint64_t seed, uint64_t step, uint32_t stream_id, size_t n_words,
                                  uint32_t* bits) {
  NTF_REQUIRE(ctx && bits, NTF_ERR_BAD_ARG, "fill_sign_bits: null pointer");
  if (!n_words) return NTF_OK;
  NTF_COUNT_LAUNCH; fill_sign_bits_kernel<<<grid_for(ctx, (n_words + 3) / 4), 256, 0, as_stream(stream)>>>(seed, step, stream_id, n_words, bits, ctx->dyn_override);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

extern "C" int ntf_apply_sign(ntf_ctx* ctx, void* stream, const float* A, const uint32_t* bits, int pitch_words, int B, int h, float* As) {
  NTF_REQUIRE(ctx && A && bits && As, NTF_ERR_BAD_ARG, "apply_sign: null pointer");
  NTF_REQUIRE(B > 0 && h > 0 && pitch_words * 32 >= h, NTF_ERR_BAD_ARG, "apply_sign: B=%d h=%d pitch=%d", B, h, pitch_words);
  NTF_COUNT_LAUNCH; apply_sign_kernel<<<(unsigned)(((size_t)B * h + 255) / 256), 256, 0, as_stream(stream)>>>(A, bits, pitch_words, B, h, As);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

// Y += X * sign: the input-sign half of the Flipout backward, dA = dz mu_W + ((dz*s_out) W_delta) * s_in
extern "C" int ntf_add_signed(ntf_ctx* ctx, void* stream, const float* X, const uint32_t* bits, int pitch_words, int B, int h, float* Y) {
  NTF_REQUIRE(ctx && X && bits && Y, NTF_ERR_BAD_ARG, "add_signed: null pointer");
  NTF_REQUIRE(B > 0 && h > 0 && pitch_words * 32 >= h, NTF_ERR_BAD_ARG, "add_signed: B=%d h=%d pitch=%d", B, h, pitch_words);
  NTF_COUNT_LAUNCH; add_signed_kernel<<<(unsigned)(((size_t)B * h + 255) / 256), 256, 0, as_stream(stream)>>>(X, bits, pitch_words, B, h, 1, Y);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

extern "C" size_t ntf_flipout_prepare_workspace_bytes(const ntf_ctx* ctx) { return (size_t)(ctx ? ctx->sm_count : 148) * 8 * sizeof(float); }

extern "C" int ntf_flipout_prepare(ntf_ctx* ctx, void* stream, const float* mu, const float* rho, const float* eps, size_t n,
                                   float kl_scale, float* delta, float* kl_out, void* workspace, size_t workspace_bytes) {
  NTF_REQUIRE(ctx && mu && rho && eps && delta && kl_out && workspace, NTF_ERR_BAD_ARG, "flipout_prepare: null pointer");
  NTF_REQUIRE(workspace_bytes >= ntf_flipout_prepare_workspace_bytes(ctx), NTF_ERR_WORKSPACE, "flipout_prepare: workspace too small");
  if (!n) return NTF_OK;
  const int blocks = grid_for(ctx, n);
  NTF_COUNT_LAUNCH; flipout_prepare_kernel<<<blocks, 256, 0, as_stream(stream)>>>(mu, rho, eps, n, delta, (float*)workspace);
  NTF_COUNT_LAUNCH; kl_finish_kernel<<<1, 256, 0, as_stream(stream)>>>((const float*)workspace, blocks, kl_scale, kl_out);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

extern "C" int ntf_flipout_grads(ntf_ctx* ctx, void* stream, const float* mu, const float* rho, const float* eps,
                                 const float* g_delta, size_t n, float kl_gscale, float* g_mu, float* g_rho) {
  NTF_REQUIRE(ctx && mu && rho && eps && g_delta && g_mu && g_rho, NTF_ERR_BAD_ARG, "flipout_grads: null pointer");
  if (!n) return NTF_OK;
  NTF_COUNT_LAUNCH; flipout_grads_kernel<<<grid_for(ctx, n), 256, 0, as_stream(stream)>>>(mu, rho, eps, g_delta, n, kl_gscale, g_mu, g_rho);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}
