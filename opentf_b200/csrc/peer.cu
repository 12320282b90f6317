// peer.cu -- data-parallel gradient exchange + optimiser as ONE pass over peer memory.
//
// The textbook sequence for data-parallel ranks is  all-reduce(grads) -> Adam on every rank: 2*(G-1)/G * 4P bytes over the links
// per rank, then 28P bytes of HBM traffic per rank, with every rank repeating the same update.  Here a rank OWNS one slice of the
// flat arena (1/G of it):
//
//   barrier      every rank's gradients are complete                              (flags in peer memory, release/acquire .sys)
//   reduce+step  g = sum over ranks r = 0..G-1 of grads_r[i]  -- 16-byte loads straight from the peers' arenas over NVLink,
//                Adam on (p, m, v)[i] of the slice, new p[i] stored into EVERY rank's parameter arena (peer stores)
//   barrier      every slice has landed everywhere
//
// i.e. reduce-scatter, optimiser and all-gather fused in one kernel: 4P*(G-1)/G bytes in and out per rank (both directions of the
// links busy at once), Adam's HBM traffic and its m/v state cut by G, no staging buffers, and the sum is taken in rank order by ONE
// rank per element, so the replicas stay bit-identical.  Arenas and flags are plain cudaMalloc blocks shared through CUDA IPC
// (ntf_peer_*); one process per GPU.
#include <stdlib.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data written by another GPU before the barrier: read past L1
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// flag block of a rank (uint32 words): [barrier b = 2*channel + phase][peer q] = the epoch rank q has reached at that barrier; [32+b] =
// this rank's own epoch counter (device-side, so that a replayed graph needs no changing argument); [36] = error word (a wait gave up).
// Two channels: the two exchanges of a step (output layer's segment on a side stream, the rest on the main stream) may overlap.
constexpr int FLAG_EPOCH = 4 * NTF_MAX_PEERS, FLAG_ERR = FLAG_EPOCH + 4;
static_assert((FLAG_ERR + 1) * 4 <= NTF_PEER_FLAG_BYTES, "flag block");

__global__ void peer_barrier_kernel(ntf_peers pr, int phase /* 2*channel + {0,1} */) {
  __shared__ uint32_t epoch_s;
  uint32_t* mine = pr.flags[pr.rank];
  if (threadIdx.x == 0) epoch_s = ++mine[FLAG_EPOCH + phase];
  __syncthreads();
  const uint32_t epoch = epoch_s;
  const int q = threadIdx.x;
  if (q < pr.world) {
    __threadfence_system();  // everything this GPU wrote before (gradients / parameter slices) is visible before the flag is
    st_release_sys(pr.flags[q] + phase * NTF_MAX_PEERS + pr.rank, epoch);
    const uint32_t* slot = mine + phase * NTF_MAX_PEERS + q;
    unsigned long long t0 = 0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int32_t)(ld_acquire_sys(slot) - epoch) < 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) { mine[FLAG_ERR] = 1u + (uint32_t)q; break; }  // 20 s: a peer never arrived -- give up instead of hanging the GPU
      __nanosleep(40);
    }
  }
}

// Software-pipelined: a thread issues the peer loads of its NEXT batch (UNR grid-strided elements, UNR*G independent 16-byte loads)
// before it steps and stores the current one.  A slice is only a few passes of the grid, so without the overlap the kernel would
// first read over the links and then write over them, using each direction half of the time; with it the inbound gradient loads
// of pass k+1 and the outbound parameter stores of pass k are in flight together (NVLink is full duplex), and enough loads are
// outstanding to cover the link + switch latency (Little's law).
template <int G, int UNR>
__global__ void __launch_bounds__(512) peer_reduce_adam_kernel(ntf_peers pr, float* __restrict__ m, float* __restrict__ v, size_t lo4, size_t hi4,
                                                               AdamK k, const ntf_dyn* __restrict__ dyn) {
  if (dyn) k = AdamK{dyn->one_minus_b1, dyn->b2, dyn->one_minus_b2, dyn->bc2_sqrt, dyn->eps, dyn->neg_step};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float4* P = reinterpret_cast<float4*>(pr.params[pr.rank]);
  float4* M = reinterpret_cast<float4*>(m);
  float4* V = reinterpret_cast<float4*>(v);
  float4 g[2][UNR][G];
  auto load = [&](int buf, size_t i0) {
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const size_t i = i0 + (size_t)u * stride;
      if (i < hi4) {
#pragma unroll
        for (int r = 0; r < G; ++r) g[buf][u][r] = ld_peer(reinterpret_cast<const float4*>(pr.grads[r]) + i);
      }
    }
  };
  auto step_store = [&](int buf, size_t i0) {
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const size_t i = i0 + (size_t)u * stride;
      if (i < hi4) {
        float4 p = P[i], mm = M[i], vv = V[i];
        float4 s = g[buf][u][0];
#pragma unroll
        for (int r = 1; r < G; ++r) { s.x += g[buf][u][r].x; s.y += g[buf][u][r].y; s.z += g[buf][u][r].z; s.w += g[buf][u][r].w; }  // rank order: the same sum on every rank
        adam1(p.x, s.x, mm.x, vv.x, k); adam1(p.y, s.y, mm.y, vv.y, k); adam1(p.z, s.z, mm.z, vv.z, k); adam1(p.w, s.w, mm.w, vv.w, k);
        M[i] = mm; V[i] = vv;
#pragma unroll
        for (int r = 0; r < G; ++r) reinterpret_cast<float4*>(pr.params[r])[i] = p;
      }
    }
  };
  size_t i0 = lo4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= hi4) return;
  load(0, i0);
  for (;;) {  // (two passes per trip so that the buffer index is a compile-time constant: g stays in registers)
    const size_t i1 = i0 + stride * UNR;
    if (i1 < hi4) load(1, i1);
    step_store(0, i0);
    if (i1 >= hi4) break;
    i0 = i1 + stride * UNR;
    if (i0 < hi4) load(0, i0);
    step_store(1, i1);
    if (i0 >= hi4) break;
  }
}

}  // namespace

// floats [off, off+n) of the arena, n and off multiples of 4: rank r owns the r-th of `world` nearly equal float4 ranges
static void peer_slice(size_t off, size_t n, int rank, int world, size_t* lo4, size_t* hi4) {
  const size_t n4 = n / 4, o4 = off / 4;
  *lo4 = o4 + n4 * (size_t)rank / (size_t)world;
  *hi4 = o4 + n4 * (size_t)(rank + 1) / (size_t)world;
}

int ntf_peer_exchange_adam_impl(ntf_ctx* ctx, cudaStream_t st, const ntf_peers* pr, float* adam_m, float* adam_v, size_t off, size_t n, double lr,
                                double beta1, double beta2, double eps, int64_t step, const ntf_dyn* dyn, int channel) {
  NTF_REQUIRE(ctx && pr && adam_m && adam_v && (channel == 0 || channel == 1), NTF_ERR_BAD_ARG, "peer_exchange_adam: null pointer / channel");
  NTF_REQUIRE(pr->world >= 1 && pr->world <= NTF_MAX_PEERS && pr->rank >= 0 && pr->rank < pr->world, NTF_ERR_BAD_ARG,
              "peer_exchange_adam: rank %d of %d (at most %d)", pr->rank, pr->world, NTF_MAX_PEERS);
  NTF_REQUIRE((off % 4) == 0 && (n % 4) == 0, NTF_ERR_BAD_ARG, "peer_exchange_adam: segment [%zu,+%zu) must be a multiple of 4 floats", off, n);
  NTF_REQUIRE(dyn || step >= 1, NTF_ERR_BAD_ARG, "peer_exchange_adam: step=%lld (1-based)", (long long)step);
  for (int r = 0; r < pr->world; ++r)
    NTF_REQUIRE(pr->grads[r] && pr->params[r] && pr->flags[r] && (((uintptr_t)pr->grads[r] | (uintptr_t)pr->params[r]) % 16) == 0, NTF_ERR_BAD_ARG,
                "peer_exchange_adam: arena of rank %d missing or not 16-byte aligned", r);
  NTF_REQUIRE((((uintptr_t)adam_m | (uintptr_t)adam_v) % 16) == 0, NTF_ERR_BAD_ARG, "peer_exchange_adam: Adam state not 16-byte aligned");
  if (n == 0) return NTF_OK;
  size_t lo4, hi4;
  peer_slice(off, n, pr->rank, pr->world, &lo4, &hi4);
  const AdamK k = adam_consts(lr, beta1, beta2, eps, step >= 1 ? step : 1);
  NTF_COUNT_LAUNCH; peer_barrier_kernel<<<1, 32, 0, st>>>(*pr, 2 * channel);
  NTF_LAUNCH_CHECK();
  if (hi4 > lo4) {
    const size_t work = hi4 - lo4;
    const int unr = pr->world <= 2 ? 2 : 1;
    size_t want = (work + (size_t)512 * unr * 4 - 1) / ((size_t)512 * unr * 4);  // CTAs for ~4 pipelined passes per thread
    const size_t cap = (size_t)ctx->sm_count * 2, floor_ = (size_t)ctx->sm_count;  // (at least one CTA per SM issuing)
    if (want > cap) want = cap;
    if (want < floor_) want = floor_ < (work + 511) / 512 ? floor_ : (work + 511) / 512;
    const int blocks = (int)(want ? want : 1);
    NTF_COUNT_LAUNCH;
    switch (pr->world) {
#define CASE(G) case G: peer_reduce_adam_kernel<G, (G <= 2 ? 2 : 1)><<<blocks, 512, 0, st>>>(*pr, adam_m, adam_v, lo4, hi4, k, dyn); break;
      CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    }
    NTF_LAUNCH_CHECK();
  }
  NTF_COUNT_LAUNCH; peer_barrier_kernel<<<1, 32, 0, st>>>(*pr, 2 * channel + 1);
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

extern "C" int ntf_peer_exchange_adam(ntf_ctx* ctx, void* stream, const ntf_peers* peers, float* adam_m, float* adam_v, size_t offset, size_t n,
                                      double lr, double beta1, double beta2, double eps, int64_t step, int channel) {
  return ntf_peer_exchange_adam_impl(ctx, as_stream(stream), peers, adam_m, adam_v, offset, n, lr, beta1, beta2, eps, step, ctx ? ctx->dyn_override : nullptr, channel);
}

// ---- expert-sharded output layer (BASELINE configs[3]): the one exchange of a step is dA = sum over shards of dz_shard . W_shard,
// [B, h] fp32 (512 KB at B = 1000): latency-bound, so it is ONE pass over peer memory between two flag barriers instead of a collective
// call between two halves of the step.  Every rank's partial sits in a peer-visible block (ntf_peers::grads[r]); every rank reads all of
// them (16-byte loads over NVLink) and adds them in RANK ORDER into its own private result -- the same sum, bit for bit, on every rank,
// so the replicated hidden layers stay identical without exchanging their gradients.
namespace {
template <int G>
__global__ void __launch_bounds__(256) peer_allreduce_kernel(ntf_peers pr, size_t n4, float4* __restrict__ dst) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v[G];
#pragma unroll
    for (int r = 0; r < G; ++r) v[r] = ld_peer(reinterpret_cast<const float4*>(pr.grads[r]) + i);
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < G; ++r) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
    dst[i] = s;
  }
}
}  // namespace

int ntf_peer_allreduce_impl(ntf_ctx* ctx, cudaStream_t st, const ntf_peers* pr, size_t n, float* dst, int channel) {
  NTF_REQUIRE(ctx && pr && dst && (channel == 0 || channel == 1), NTF_ERR_BAD_ARG, "peer_allreduce: null pointer / channel");
  NTF_REQUIRE(pr->world >= 1 && pr->world <= NTF_MAX_PEERS && pr->rank >= 0 && pr->rank < pr->world, NTF_ERR_BAD_ARG, "peer_allreduce: rank %d of %d", pr->rank, pr->world);
  NTF_REQUIRE((n % 4) == 0 && ((uintptr_t)dst % 16) == 0, NTF_ERR_BAD_ARG, "peer_allreduce: n must be a multiple of 4 floats, dst 16-byte aligned");
  for (int r = 0; r < pr->world; ++r)
    NTF_REQUIRE(pr->grads[r] && pr->flags[r] && ((uintptr_t)pr->grads[r] % 16) == 0, NTF_ERR_BAD_ARG, "peer_allreduce: block of rank %d missing or misaligned", r);
  NTF_REQUIRE(dst != pr->grads[pr->rank], NTF_ERR_BAD_ARG, "peer_allreduce: the result must not alias the exchange block (peers are still reading it)");
  if (n == 0) return NTF_OK;
  NTF_COUNT_LAUNCH; peer_barrier_kernel<<<1, 32, 0, st>>>(*pr, 2 * channel);  // every rank's partial is complete
  NTF_LAUNCH_CHECK();
  const size_t n4 = n / 4;
  const int blocks = (int)((n4 + 255) / 256 < (size_t)ctx->sm_count * 2 ? (n4 + 255) / 256 : (size_t)ctx->sm_count * 2);
  NTF_COUNT_LAUNCH;
  switch (pr->world) {
#define CASE(G) case G: peer_allreduce_kernel<G><<<blocks, 256, 0, st>>>(*pr, n4, reinterpret_cast<float4*>(dst)); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
  NTF_LAUNCH_CHECK();
  NTF_COUNT_LAUNCH; peer_barrier_kernel<<<1, 32, 0, st>>>(*pr, 2 * channel + 1);  // everybody has read every block: it may be rewritten
  NTF_LAUNCH_CHECK();
  return NTF_OK;
}

extern "C" int ntf_peer_allreduce(ntf_ctx* ctx, void* stream, const ntf_peers* peers, size_t n, float* dst, int channel) {
  return ntf_peer_allreduce_impl(ctx, as_stream(stream), peers, n, dst, channel);
}

// ---- peer-visible memory: plain cudaMalloc blocks + CUDA IPC handles (one process per GPU) -------------------------------------
extern "C" int ntf_peer_alloc(ntf_ctx* ctx, size_t bytes, void** out) {
  NTF_REQUIRE(ctx && out && bytes > 0, NTF_ERR_BAD_ARG, "peer_alloc: bad argument");
  int cur = 0;
  NTF_CUDA(cudaGetDevice(&cur));
  NTF_CUDA(cudaSetDevice(ctx->device));
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaSuccess) e = cudaMemset(*out, 0, bytes);
  cudaSetDevice(cur);
  NTF_REQUIRE(e == cudaSuccess, NTF_ERR_CUDA, "peer_alloc(%zu): %s", bytes, cudaGetErrorString(e));
  return NTF_OK;
}

extern "C" int ntf_peer_free(ntf_ctx* ctx, void* p) {
  NTF_REQUIRE(ctx, NTF_ERR_BAD_ARG, "peer_free: null ctx");
  if (p) NTF_CUDA(cudaFree(p));
  return NTF_OK;
}

extern "C" int ntf_peer_export(ntf_ctx* ctx, void* p, void* handle64) {
  NTF_REQUIRE(ctx && p && handle64, NTF_ERR_BAD_ARG, "peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == NTF_PEER_HANDLE_BYTES, "handle size");
  NTF_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, p));
  return NTF_OK;
}

extern "C" int ntf_peer_import(ntf_ctx* ctx, const void* handle64, void** out) {
  NTF_REQUIRE(ctx && handle64 && out, NTF_ERR_BAD_ARG, "peer_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  int cur = 0;
  NTF_CUDA(cudaGetDevice(&cur));
  NTF_CUDA(cudaSetDevice(ctx->device));
  const cudaError_t e = cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
  cudaSetDevice(cur);
  NTF_REQUIRE(e == cudaSuccess, NTF_ERR_CUDA, "peer_import: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  return NTF_OK;
}

extern "C" int ntf_peer_release(ntf_ctx* ctx, void* imported) {
  NTF_REQUIRE(ctx, NTF_ERR_BAD_ARG, "peer_release: null ctx");
  if (imported) NTF_CUDA(cudaIpcCloseMemHandle(imported));
  return NTF_OK;
}
