// Microbenchmark: are SMALL bulk copies (cp.async.bulk shared -> global, 48..112 bytes, one per lane) a cheaper
// way to write scattered per-channel runs than 4-byte stores?  Every lane owns a "channel": runs are 196 bytes
// apart, like the RoIAlign output.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_small bulk_small.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: one bulk copy of `bytes` per lane per item; mode 1: the same bytes with 4-byte stores, lane = channel
// (32 sectors per instruction); mode 2: 4-byte stores with lanes along the run (what the kernel's staging does).
template <int MODE>
__global__ void k(float* out, int items, int bytes, long long* cycles) {
  __shared__ __align__(128) float tile[8][32 * 36];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = lane; i < 32 * 36; i += 32) tile[warp][i] = (float)i;
  __syncwarp();
  const int words = bytes / 4;
  const long long t0 = clock64();
  // every warp writes its own sequence of RoI tiles: item -> 32 channels x 49 floats
  float* base = out + ((size_t)blockIdx.x * 8 + warp) * (size_t)items * 32 * 49;
  for (int it = 0; it < items; ++it) {
    float* o = base + (size_t)it * 32 * 49;
    if (MODE == 0) {
      if (it > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      // 16-byte aligned source and destination: channel c at o + c*49 floats = 196 c bytes -> only c % 4 == 0 is
      // aligned in this toy layout, so use 52-float channel stride here (208 bytes) to keep every lane legal
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)it * 32 * 52 + lane * 52),
                   "r"(smem_u32(&tile[warp][lane * 36])), "r"(bytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    } else if (MODE == 1) {
      for (int j = 0; j < words; ++j) o[lane * 49 + j] = tile[warp][lane * 36 + j];
    } else {
      // flat order over [32 ch][words]: consecutive lanes walk along the runs
      for (int e = lane; e < 32 * words; e += 32) {
        const int c = e / words, j = e - c * words;
        o[c * 49 + j] = tile[warp][c * 36 + j];
      }
    }
  }
  if (MODE == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int bytes) {
  const int blocks = 148, items = 400;
  float* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * (size_t)blocks * 8 * items * 32 * 52);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(out, items, bytes, cyc);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, items, bytes, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double total = (double)blocks * 8 * items * 32 * bytes;
  printf("%-28s %3d B/run: %8.3f ms  %7.1f GB/s  (%.1f runs/us/SM)  %s\n", name, bytes, ms, total / ms / 1e6,
         (double)8 * items * 32 / (ms * 1e3), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int b : {48, 80, 112}) {
    run<0>("bulk copy per lane", b);
    run<1>("STG lane = channel", b);
    run<2>("STG lanes along runs", b);
  }
  return 0;
}
