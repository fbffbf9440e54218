// Shared helpers for the sm_100a detection-glue kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/b2d_glue.h"

namespace b2d {

constexpr int kNumSMs = 148;          // B200: 2 dies x 74 SMs
constexpr int kMaxSortElems = 16384;  // in-CTA bitonic capacity (128 KB of u64 keys)
constexpr int kSelectBits = 14;       // first-level radix: top bits of the score key (64 KB shared histogram per CTA)
constexpr int kSelectBins = 1 << kSelectBits;
constexpr int kSelectShift = 32 - kSelectBits;

extern std::atomic<uint64_t> g_launches;
extern thread_local int tl_cuda_error;

inline int cuda_fail(cudaError_t e) {
  tl_cuda_error = static_cast<int>(e);
  return B2D_ERR_CUDA;
}

#define B2D_CUDA(expr)                                   \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return ::b2d::cuda_fail(_e);  \
  } while (0)

// Call after every kernel launch: counts it and converts launch errors.
#define B2D_LAUNCHED()                                   \
  do {                                                   \
    ::b2d::g_launches.fetch_add(1, std::memory_order_relaxed); \
    cudaError_t _e = cudaPeekAtLastError();              \
    if (_e != cudaSuccess) return ::b2d::cuda_fail(_e);  \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ----- exact (non-contracted) fp32 arithmetic: the reference runs separate ATen ops, so
// ----- every intermediate is rounded; nvcc would otherwise fuse a*b+c into one FMA.
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// torch.clamp(x, lo, hi) = min(max(x, lo), hi) with NaN propagated.
__device__ __forceinline__ float clampf(float x, float lo, float hi) {
  if (x != x) return x;
  return fminf(fmaxf(x, lo), hi);
}

// Monotone map float -> uint32 (larger float => larger key); NaN sorts first in a
// descending sort (torch.sort semantics), -0.0 == +0.0.
__device__ __forceinline__ uint32_t score_key(float s) {
  if (s != s) return 0xFFFFFFFFu;
  uint32_t u = __float_as_uint(s + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// composite = key<<32 | ~index : descending composite == (score desc, index asc)
__device__ __forceinline__ uint64_t composite_key(uint32_t key, uint32_t idx) {
  return (static_cast<uint64_t>(key) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ uint32_t composite_index(uint64_t c) {
  return 0xFFFFFFFFu - static_cast<uint32_t>(c & 0xFFFFFFFFull);
}

// Largest float <= a double threshold: for float x, ((double)x > thr) <=> (x > floor_f(thr)).
inline float float_floor_of(double thr) {
  float f = static_cast<float>(thr);
  if (static_cast<double>(f) > thr) f = nextafterf(f, -INFINITY);
  return f;
}

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may become resident while the kernel
// before it in the stream is still running; it must call pdl_wait() before it touches anything an earlier kernel
// wrote (a no-op in a normal launch).  pdl_trigger() lets the NEXT kernel's CTAs become resident early.  Used on the
// few-frame paths, where a call is a chain of ten short kernels and the gaps between them are a fifth of its time.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
#ifdef B2D_AB_NOPDL
  pdl = false;      // A/B timing build
#endif
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk TMA helpers (sm_90+/sm_100a PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}


}  // namespace b2d
