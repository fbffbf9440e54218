// MC-dropout reductions over T stacked head passes.
// Reference: utils/loss_utils.py:103-141 (compute_bbox_cov / compute_bbox_var /
// categorical_entropy / categorical_mutual_information); datasets/db.py:264-303 (sort).
#include "common.cuh"

namespace b2d {

// samples [T, m] -> var [m].  One thread per column; T is small (10-20) so each thread
// streams T coalesced rows.  mode 0 reproduces the reference's single-pass formula and its
// summation order (sequential over T, like torch.sum over dim 0 of a [T, m] tensor).
__global__ void __launch_bounds__(256) mc_variance_kernel(int T, int m, const float* __restrict__ x, int mode,
                                                          float* __restrict__ var) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  float s = 0.0f, s2 = 0.0f;
  for (int t = 0; t < T; ++t) {
    const float v = __ldg(x + (size_t)t * m + j);
    s = fadd(s, v);
    s2 = fadd(s2, fmul(v, v));
  }
  float out;
  if (mode == 0) {
    // var = (sum x^2 + (-(sum x)^2 / n)) / (n - 1)        loss_utils.py:115-119
    const float sq = fmul(s, s);
    out = fdiv(fadd(s2, fdiv(-sq, (float)T)), (float)(T - 1));
  } else {
    // diag(E[xx^T]) - mu^2                                 loss_utils.py:104-111
    const float mu = fdiv(s, (float)T);
    out = fsub(fdiv(s2, (float)T), fmul(mu, mu));
  }
  var[j] = fmaxf(out, 0.0f);
}

// logits [T, n, K] -> mutual information [n], entropy of the mean softmax [n].
__global__ void __launch_bounds__(128) mc_class_kernel(int T, int n, int K, const float* __restrict__ logits,
                                                       float* __restrict__ mi, float* __restrict__ ent) {
  extern __shared__ float s_mean[];  // [blockDim][K]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* mean = s_mean + (size_t)threadIdx.x * K;
  for (int c = 0; c < K; ++c) mean[c] = 0.0f;
  float neg_ent_sum = 0.0f;  // sum_t sum_c p log2 p
  for (int t = 0; t < T; ++t) {
    const float* z = logits + ((size_t)t * n + i) * K;
    float mx = -INFINITY;
    for (int c = 0; c < K; ++c) mx = fmaxf(mx, __ldg(z + c));
    float den = 0.0f;
    for (int c = 0; c < K; ++c) den += expf(__ldg(z + c) - mx);
    float acc = 0.0f;
    for (int c = 0; c < K; ++c) {
      const float p = expf(__ldg(z + c) - mx) / den;
      mean[c] += p;
      acc += p * log2f(p);
    }
    neg_ent_sum += acc;
  }
  float total = 0.0f;
  for (int c = 0; c < K; ++c) {
    const float p = mean[c] / (float)T;
    total += p * log2f(p);
  }
  total = -total;
  if (ent) ent[i] = total;
  if (mi) mi[i] = neg_ent_sum / (float)T + total;
}

// key[r] = mean_j var[r][j]; order = stable argsort (ascending, or descending of -key).
__global__ void __launch_bounds__(1024) var_sort_kernel(int n, int cols, const float* __restrict__ var, int descending,
                                                        float* __restrict__ key_out, int32_t* __restrict__ order) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  int n_pad = 2;
  while (n_pad < n) n_pad <<= 1;
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
    uint64_t k = 0ull;
    if (i < n) {
      float s = 0.0f;
      for (int j = 0; j < cols; ++j) s += __ldg(var + (size_t)i * cols + j);
      const float mean = cols > 0 ? s / (float)cols : 0.0f;
      if (key_out) key_out[i] = mean;
      // argsort(x) ascending == descending sort of -x; argsort(-x) == descending sort of x
      k = composite_key(score_key(descending ? mean : -mean), (uint32_t)i);
    }
    keys[i] = k;
  }
  __syncthreads();
  const int half = n_pad >> 1;
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int p = threadIdx.x; p < half; p += blockDim.x) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const int q = i | j;
        const uint64_t a = keys[i], b = keys[q];
        const bool desc = (i & k) == 0;
        if ((a < b) == desc) {
          keys[i] = b;
          keys[q] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) order[i] = (int32_t)composite_index(keys[i]);
}

}  // namespace b2d

using namespace b2d;

extern "C" int b2d_mc_variance(int T, int m, const float* samples, int mode, float* var, void* stream) {
  if (T <= 0 || m < 0 || mode < 0 || mode > 1) return B2D_ERR_INVALID_ARG;
  if (m == 0) return B2D_OK;
  if (!samples || !var) return B2D_ERR_INVALID_ARG;
  mc_variance_kernel<<<ceil_div(m, 256), 256, 0, as_stream(stream)>>>(T, m, samples, mode, var);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_mc_class_uncertainty(int T, int n, int K, const float* logits, float* mutual_info, float* entropy,
                                        void* stream) {
  if (T <= 0 || n < 0 || K <= 0) return B2D_ERR_INVALID_ARG;
  if (n == 0) return B2D_OK;
  if (!logits) return B2D_ERR_INVALID_ARG;
  if (K > 96) return B2D_ERR_UNSUPPORTED;
  mc_class_kernel<<<ceil_div(n, 128), 128, sizeof(float) * 128 * K, as_stream(stream)>>>(T, n, K, logits, mutual_info,
                                                                                       entropy);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_var_sort(int n, int cols, const float* var, int descending, float* key, int32_t* order,
                            void* stream) {
  if (n < 0 || cols < 0) return B2D_ERR_INVALID_ARG;
  if (n == 0) return B2D_OK;
  if (!var || !order) return B2D_ERR_INVALID_ARG;
  if (n > kMaxSortElems) return B2D_ERR_UNSUPPORTED;
  int n_pad = 2;
  while (n_pad < n) n_pad <<= 1;
  const size_t smem = sizeof(uint64_t) * n_pad;
  B2D_CUDA(cudaFuncSetAttribute(var_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(sizeof(uint64_t) * kMaxSortElems)));
  var_sort_kernel<<<1, 1024, smem, as_stream(stream)>>>(n, cols, var, descending, key, order);
  B2D_LAUNCHED();
  return B2D_OK;
}
