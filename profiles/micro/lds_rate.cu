// Microbenchmark: sustained shared-memory load rate of one SM for conflict-free 32/64/128-bit loads
// as a function of resident warps (what ceiling does the RoIAlign tap stream run against?).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_rate lds_rate.cu ; run: ./lds_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int VEC>
__global__ void lds_kernel(int iters, float* out, long long* cycles) {
  extern __shared__ float smem[];
  for (int i = threadIdx.x; i < 24576; i += blockDim.x) smem[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // lane-contiguous addresses (conflict-free); every warp walks its own window
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)(lane * 4 * VEC) + (uint32_t)(warp * 1024);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // 32 distinct loads per iteration; the window start moves with `it` so nothing can be merged
    const uint32_t b = base + (uint32_t)((it & 7) * 128);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const uint32_t a = b + (uint32_t)(k * 128 * VEC);
      if (VEC == 1) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
        acc[k & 7] += v;
      } else if (VEC == 2) {
        float v, w;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v), "=f"(w) : "r"(a) : "memory");
        acc[k & 7] += v + w;
      } else {
        float v, w, x, y;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v), "=f"(w), "=f"(x), "=f"(y) : "r"(a) : "memory");
        acc[k & 7] += (v + w) + (x + y);
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}

template <int VEC>
void run(int warps) {
  const int iters = 2000, blocks = 148;
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * blocks * warps * 32);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  cudaFuncSetAttribute(lds_kernel<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
  lds_kernel<VEC><<<blocks, warps * 32, 98304>>>(iters, out, cyc);
  lds_kernel<VEC><<<blocks, warps * 32, 98304>>>(iters, out, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += (double)h[i];
  avg /= blocks;
  const double wavefronts = (double)iters * 32 * warps * VEC;   // 128 bytes each
  printf("LDS.%d  warps/SM %2d : %.3f wavefronts/clk/SM (%.3f instr/clk/SM)\n", 32 * VEC, warps, wavefronts / avg,
         (double)iters * 32 * warps / avg);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {2, 4, 8, 10, 12, 16, 32}) run<1>(w);
  for (int w : {2, 4, 8, 16, 32}) run<2>(w);
  for (int w : {2, 4, 8, 16, 32}) run<4>(w);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
