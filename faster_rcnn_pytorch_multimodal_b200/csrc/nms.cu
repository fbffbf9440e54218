// Greedy NMS on score-sorted boxes, one CTA per frame, no global bitmask.
//
// Semantics restated from torchvision's CPU kernel (the reference's third-party call,
// layer_utils/proposal_layer.py:46, utils/filter_predictions.py:67,69):
//   area = (x2-x1)*(y2-y1); box j is suppressed by an earlier KEPT box i iff
//   (double)(inter / (area_i + area_j - inter)) > thresh;  0/0 = NaN never suppresses.
//
// Only kept boxes can suppress, so instead of the O(n^2) n x n/64 bitmask of the stock
// kernel the sweep walks candidates in tiles of 64:
//   A. tile x kept-list IoU (kept boxes staged in shared memory) -> 64 alive bits
//   B. 64 x 64 intra-tile IoU block -> one u64 suppression row per candidate
//   C. ballot-free serial resolve over the alive bits (find-first-set walk)
// and stops as soon as max_keep boxes are kept (post_nms_topN), which the reference's
// keep[:post] slice makes equivalent.  Scratch traffic to HBM: none.
#include "common.cuh"

namespace b2d {

constexpr int kNmsThreads = 1024;
constexpr int kTile = 64;
constexpr int kGroup = kNmsThreads / kTile;  // threads cooperating on one candidate (16)
constexpr int kKeptSmem = 2560;              // kept boxes cached in shared memory (40 KB)

__device__ __forceinline__ bool iou_exceeds(const float4 a, const float area_a, const float4 b, const float thr_f) {
  const float w = fmaxf(0.0f, fsub(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  const float h = fmaxf(0.0f, fsub(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  const float inter = fmul(w, h);
  if (!(inter > 0.0f) && thr_f >= 0.0f) return false;  // iou is 0, -0 or NaN: never > thr
  const float area_b = fmul(fsub(b.z, b.x), fsub(b.w, b.y));
  const float iou = fdiv(inter, fsub(fadd(area_a, area_b), inter));
  return iou > thr_f;
}

__global__ void __launch_bounds__(kNmsThreads) nms_sorted_kernel(const float4* __restrict__ boxes_all, int n,
                                                                 const int32_t* __restrict__ n_valid, float thr_f,
                                                                 int max_keep, int32_t* __restrict__ keep_all,
                                                                 int32_t* __restrict__ num_keep) {
  __shared__ float4 s_kept[kKeptSmem];
  __shared__ float4 s_tile[kTile];
  __shared__ unsigned long long s_rows[kTile];
  __shared__ unsigned int s_dead[2];
  __shared__ int s_k;

  const int f = blockIdx.x;
  const float4* boxes = boxes_all + (size_t)f * n;
  int32_t* keep = keep_all + (size_t)f * max_keep;
  const int nv = n_valid ? min(n_valid[f], n) : n;
  const int tid = threadIdx.x;
  const int cand = tid / kGroup;  // 0..63
  const int sub = tid % kGroup;   // 0..15
  const unsigned group_mask = 0xFFFFu << ((tid & 16) ? 16 : 0);

  if (tid == 0) s_k = 0;
  __syncthreads();

  for (int base = 0; base < nv; base += kTile) {
    const int K = s_k;
    if (K >= max_keep) break;
    const int tile_n = min(kTile, nv - base);
    if (tid < kTile) s_tile[tid] = tid < tile_n ? boxes[base + tid] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 2) s_dead[tid] = 0u;
    __syncthreads();

    // ---- A: candidate vs. every kept box
    const float4 me = s_tile[cand];
    const float my_area = fmul(fsub(me.z, me.x), fsub(me.w, me.y));
    bool dead = false;
    if (cand < tile_n) {
      for (int kk = sub; kk < K; kk += kGroup) {
        const float4 kb = kk < kKeptSmem ? s_kept[kk] : boxes[keep[kk]];
        // torchvision: iarea + areas[j] - inter with i the kept box; fp add is commutative
        if (iou_exceeds(me, my_area, kb, thr_f)) {
          dead = true;
          break;
        }
      }
    }
    const unsigned dead_ballot = __ballot_sync(0xFFFFFFFFu, dead);
    if ((tid & 31) == 0) {
      // warp covers candidates 2w and 2w+1
      const int w = tid >> 5;
      unsigned bits = ((dead_ballot & 0xFFFFu) ? 1u : 0u) | ((dead_ballot >> 16) ? 2u : 0u);
      if (bits) atomicOr(&s_dead[w >> 4], bits << ((2 * w) & 31));
    }

    // ---- B: intra-tile block, row `cand` vs columns sub*4 .. sub*4+3 (only j > cand)
    unsigned lo = 0u, hi = 0u;
    if (cand < tile_n) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = sub * 4 + c;
        if (j > cand && j < tile_n) {
          if (iou_exceeds(me, my_area, s_tile[j], thr_f)) {
            if (j < 32) lo |= 1u << j; else hi |= 1u << (j - 32);
          }
        }
      }
    }
    lo = __reduce_or_sync(group_mask, lo);
    hi = __reduce_or_sync(group_mask, hi);
    if (sub == 0) s_rows[cand] = ((unsigned long long)hi << 32) | lo;
    __syncthreads();

    // ---- C: resolve (single thread; iterations = boxes kept from this tile)
    if (tid == 0) {
      unsigned long long alive = ~(((unsigned long long)s_dead[1] << 32) | s_dead[0]);
      if (tile_n < 64) alive &= (1ull << tile_n) - 1ull;
      int k = K;
      while (alive && k < max_keep) {
        const int i = __ffsll((long long)alive) - 1;
        alive &= ~(1ull << i);
        alive &= ~s_rows[i];
        keep[k] = base + i;
        if (k < kKeptSmem) s_kept[k] = s_tile[i];
        ++k;
      }
      s_k = k;
    }
    __syncthreads();
  }
  if (tid == 0) num_keep[f] = s_k;
}

int launch_nms_sorted(int F, int n, const float* boxes, const int32_t* n_valid, double thresh, int max_keep,
                      int32_t* keep, int32_t* num_keep, cudaStream_t st) {
  if (max_keep <= 0 || n <= 0) {
    B2D_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int32_t) * F, st));
    return B2D_OK;
  }
  nms_sorted_kernel<<<F, kNmsThreads, 0, st>>>(reinterpret_cast<const float4*>(boxes), n, n_valid,
                                               float_floor_of(thresh), max_keep, keep, num_keep);
  B2D_LAUNCHED();
  return B2D_OK;
}

}  // namespace b2d

extern "C" int b2d_nms_sorted(int F, int n, const float* boxes, const int32_t* n_valid, double thresh, int max_keep,
                              int32_t* keep, int32_t* num_keep, void* stream) {
  if (F <= 0 || n < 0 || !boxes || !keep || !num_keep) return B2D_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(boxes) & 15u) != 0) return B2D_ERR_INVALID_ARG;  // float4 loads
  return b2d::launch_nms_sorted(F, n, boxes, n_valid, thresh, max_keep, keep, num_keep, b2d::as_stream(stream));
}
