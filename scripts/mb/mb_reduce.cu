// micro-benchmark: how fast can one SM (and 148 at once) add a [128 x 128] fp32 tile into a small [B,128] matrix in L2?
//   mode 0: cp.reduce.async.bulk.tensor (TMA reduce-add) of 16 KB chunks from shared memory, `depth` chunks in flight
//   mode 1: red.global.add.v4.f32 from registers (128 threads, thread = row, 32 columns per chunk)
//   mode 2: red.global.add.f32 scalar, coalesced (lane = column)
//   mode 3: mode 0 with ONE 64 KB-equivalent issue burst per tile (4 chunks issued back to back from 4 buffers, then wait all)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_reduce mb_reduce.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap map, float* dA, int B, int tiles, int depth, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int r = threadIdx.x;
  const uint32_t sbase = smem_u32(sm);
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 1.0f + i;
  __syncthreads();
  const long long t0 = clock64();
  int nbuf = 0;
  for (int t = 0; t < tiles; ++t) {
    const int row0 = ((t + blockIdx.x) % (B / 128)) * 128;
    for (int c = 0; c < 4; ++c) {
      if (MODE == 0 || MODE == 3) {
        const int buf = nbuf % depth;
        if (r == 0) {
          // wait until at most depth-1 reduces are still reading shared memory
          switch (depth) {
            case 1: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
            case 2: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
            case 3: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
            default: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
          }
        }
        __syncthreads();
        uint8_t* row = sm + buf * 16384 + r * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(row + ((q ^ (r & 7)) << 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (r == 0) {
          asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                       ::"l"(reinterpret_cast<uint64_t>(&map)), "r"(c * 32), "r"(row0), "r"(sbase + buf * 16384) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++nbuf;
      } else if (MODE == 1) {
        float* dst = dA + (size_t)(row0 + r) * 128 + c * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "f"(v[4 * q]), "f"(v[4 * q + 1]), "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
      } else if (MODE == 2) {
        // lane = column: warp w handles rows w*32.., 32 rows x 32 columns per chunk
        const int w = r >> 5, l = r & 31;
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(dA + (size_t)(row0 + w * 32 + i) * 128 + c * 32 + l, v[i]);
      }
    }
  }
  if (MODE == 0 || MODE == 3) { if (r == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
  __syncthreads();
  const long long t1 = clock64();
  if (r == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int B = 1024, tiles = 64;
  float* dA; long long* cyc;
  CK(cudaMalloc(&dA, (size_t)B * 128 * 4)); CK(cudaMemset(dA, 0, (size_t)B * 128 * 4));
  CK(cudaMalloc(&cyc, 1024 * 8));
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  CUtensorMap map;
  const cuuint64_t gdim[2] = {128, (cuuint64_t)B}; const cuuint64_t gstride[1] = {128 * 4};
  const cuuint32_t box[2] = {32, 128}; const cuuint32_t estr[2] = {1, 1};
  CUresult cr = ((EncodeTiledFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { printf("encode failed %d\n", (int)cr); return 1; }
  CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  long long h[1024];
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int grid : {1, 148, 296}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int depth = 1; depth <= (mode == 0 ? 4 : 1); ++depth) {
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          if (mode == 0) k<0><<<grid, 128, 65536>>>(map, dA, B, tiles, depth, cyc);
          else if (mode == 1) k<1><<<grid, 128, 0>>>(map, dA, B, tiles, depth, cyc);
          else k<2><<<grid, 128, 0>>>(map, dA, B, tiles, depth, cyc);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
        long long mx = 0; double avg = 0; for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; avg += h[i]; }
        avg /= grid;
        const double bytes = (double)grid * tiles * 65536;
        printf("grid %3d mode %d depth %d: %.1f us, avg %.0f cyc/tile(64KB) max %.0f; aggregate %.2f TB/s of reduce payload\n", grid, mode, depth, ms * 1e3,
               avg / tiles, (double)mx / tiles, bytes / (ms * 1e-3) / 1e12);
      }
    }
  }
  return 0;
}
