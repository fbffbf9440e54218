// Microbenchmark: how fast can an SM push whole RoIAlign output tiles ([32 ch][49] fp32 = 6272 contiguous bytes) out of
// shared memory?  mode 0: the kernel's way - 49 conflict-free STS per lane, fence.proxy.async, ONE cp.async.bulk
// shared -> global per tile (the warp reuses its tile after wait_group.read);  mode 1: the same STS, then the tile
// leaves with coalesced 16-byte stores (13 LDS.128 + 13 STG.128 per tile);  mode 2: registers straight to global with
// 4-byte stores at stride 49 words (what a kernel without the tile would do).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tile_store tile_store.cu
// B200, 148 CTAs, 10 warps each (write-only): bulk store 6.2 TB/s (294 cycles per tile per SM), coalesced 16-byte stores
// 6.2 TB/s, stride-49 4-byte stores 0.5 TB/s.  So the tile path of the RoIAlign kernel is not capped by the bulk-store engine:
// the BEV workload (every RoI a whole tile) spends 470 cycles per tile per SM.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int kTileWords = 32 * 49;

template <int MODE>
__global__ void k(float* out, int items, int warps) {
  extern __shared__ __align__(128) float tiles[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* tile = tiles + (size_t)warp * kTileWords;
  float acc[49];
#pragma unroll
  for (int i = 0; i < 49; ++i) acc[i] = (float)(i + lane);
  float* base = out + ((size_t)blockIdx.x * warps + warp) * (size_t)items * kTileWords;
  for (int it = 0; it < items; ++it) {
    float* o = base + (size_t)it * kTileWords;
#pragma unroll
    for (int i = 0; i < 49; ++i) acc[i] += 1.0f;      // (something to store)
    if (MODE == 0) {
      if (it > 0) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
      }
#pragma unroll
      for (int i = 0; i < 49; ++i) tile[lane * 49 + i] = acc[i];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(o), "r"(smem_u32(tile)), "r"(kTileWords * 4) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 49; ++i) tile[lane * 49 + i] = acc[i];
      __syncwarp();
      const float4* t4 = reinterpret_cast<const float4*>(tile);
      float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        const int e = lane + 32 * j;
        if (e < kTileWords / 4) o4[e] = t4[e];
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int i = 0; i < 49; ++i) o[lane * 49 + i] = acc[i];
    }
  }
  if (MODE == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE>
void run(const char* name, int warps) {
  const int blocks = 148, items = 600;
  float* out;
  cudaMalloc(&out, sizeof(float) * (size_t)blocks * warps * items * kTileWords);
  const size_t smem = sizeof(float) * (size_t)warps * kTileWords;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, warps * 32, smem>>>(out, items, warps);
  cudaEventRecord(e0);
  k<MODE><<<blocks, warps * 32, smem>>>(out, items, warps);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)blocks * warps * items * kTileWords * 4;
  printf("%-44s %2d warps/SM: %7.1f GB/s (%.1f GB/s per SM, %.0f cycles per tile per SM at 1.965 GHz), %s\n", name, warps, bytes / ms / 1e6,
         bytes / ms / 1e6 / blocks, 1.965e9 * (ms * 1e-3) / ((double)warps * items), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main(int argc, char** argv) {
  for (int warps : {4, 10, 16}) {
    run<0>("bulk store of the tile (cp.async.bulk)", warps);
    run<1>("coalesced 16-byte stores from the tile", warps);
    run<2>("4-byte stores from registers, stride 49", warps);
  }
  return 0;
}
