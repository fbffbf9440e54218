// Proposal stage: fg-score select (2-level radix), in-CTA bitonic sort, decode + clip,
// greedy NMS (nms.cu) and the post-NMS gather.
//
// Reference behaviour restated here: layer_utils/proposal_layer.py:18-57,
// layer_utils/proposal_top_layer.py:18-59, model/bbox_transform.py:75-105,235-257.
//
// Data flow per frame (all launches cover every frame of the batch):
//   1. score_hist_kernel      N fg scores -> 65536-bin histogram of the top 16 key bits
//   2. score_threshold_kernel histogram -> threshold bin b* (smallest set of top bins holding >= k)
//   3. score_compact_kernel   scores with bin >= b*  -> candidate composites (unordered)
//   4. sort_decode_kernel     candidates -> exact top-k, sorted (score desc, index asc);
//                             decodes + clips ONLY those k boxes (the reference decodes all N)
//   5. nms_sorted_kernel      (nms.cu)
//   6. proposal_gather_kernel kept positions -> rois / scores / anchors_3d rows
#include <cooperative_groups.h>

#include "common.cuh"

namespace b2d {

int launch_nms_sorted(int F, int n, const float* boxes, const int32_t* n_valid, double thresh, int max_keep,
                      int32_t* keep, int32_t* num_keep, void* kept_scratch, cudaStream_t st);
size_t nms_cluster_workspace_bytes(int F, int max_keep);

struct ProposalWs {
  uint32_t* hist;         // [F][kSelectBins]
  uint32_t* sel;          // [F][4]: thr_bin, n_cand (atomic), k, n_above
  uint64_t* cand;         // [F][N]
  float* sorted_boxes;    // [F][kmax][4]
  float* sorted_scores;   // [F][kmax]
  int32_t* sorted_index;  // [F][kmax]
  int32_t* n_sorted;      // [F]
  int32_t* keep;          // [F][max_out]
  int32_t* num_keep;      // [F]
  void* nms_kept;         // kept boxes of the cluster NMS (few-frame path)
  uint64_t* cand2;        // [F][N]: second buffer of the chunk-sort + merge path (pre_nms > kMaxSortElems only)
  size_t bytes;
};

// Few frames per call (the reference API issues ONE): the single-CTA-per-frame bitonic sort leaves 147 SMs idle for
// 100 us.  Up to kRankMaxFrames frames the candidate list is sorted in runs of kRunLen, one CTA each, and every
// candidate finds its rank by binary search in the other runs (run_sort_kernel, rank_scatter_kernel).
constexpr int kRankMaxFrames = 8;
constexpr int kRunLen = 512;                                       // candidates per sorted run (one CTA of 256 threads)
static_assert(kRunLen == 512, "rank_scatter_kernel: ten binary-search steps per run");

static int top_k_of(int N, int pre) { return (pre > 0 && pre < N) ? pre : N; }
static int max_out_of(int N, int pre, int post) {
  int k = top_k_of(N, pre);
  return (post > 0 && post < k) ? post : k;
}

static ProposalWs carve(void* base, int F, int N, int pre, int post) {
  ProposalWs w;
  const int k = top_k_of(N, pre);
  const int mo = max_out_of(N, pre, post);
  size_t off = 0;
  char* p = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  w.hist = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * (size_t)F * kSelectBins));
  w.sel = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * (size_t)F * 4));
  w.cand = reinterpret_cast<uint64_t*>(take(sizeof(uint64_t) * (size_t)F * N));
  w.sorted_boxes = reinterpret_cast<float*>(take(sizeof(float) * (size_t)F * k * 4));
  w.sorted_scores = reinterpret_cast<float*>(take(sizeof(float) * (size_t)F * k));
  w.sorted_index = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F * k));
  w.n_sorted = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F));
  w.keep = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F * mo));
  w.num_keep = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F));
  const size_t nk = nms_cluster_workspace_bytes(F, mo);
  w.nms_kept = nk ? take(nk) : nullptr;
  w.cand2 = k > kMaxSortElems ? reinterpret_cast<uint64_t*>(take(sizeof(uint64_t) * (size_t)F * N)) : nullptr;
  w.bytes = off;
  return w;
}

// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float fg_score(const float* __restrict__ cls_prob, int f, int n_loc, int A, int n) {
  const int l = n / A;
  const int a = n - l * A;
  return __ldg(cls_prob + ((size_t)f * n_loc + l) * (2 * A) + A + a);
}

// Flat anchor index n -> fg score, with the division by A done by multiply-high (A * magic >= 2^32).
__device__ __forceinline__ float fg_score_fast(const float* __restrict__ frame_prob, int A, uint32_t magic, int n) {
  int l = (int)__umulhi((uint32_t)n, magic);
  int a = n - l * A;
  if (a >= A) {   // magic rounds down by at most one
    a -= A;
    ++l;
  }
  return __ldg(frame_prob + (size_t)l * (2 * A) + A + a);
}

constexpr int kSelThreads = 512;
constexpr int kSelChunk = 16384;   // scores per CTA

// Histogram of the top kSelectBits key bits.  The histogram is privatised in shared memory (one
// global atomic per element is bound by the L2 request rate: 15 M of them took 250 us for 64
// Waymo frames) and only its non-zero bins are merged into the frame's global histogram.
__global__ void __launch_bounds__(kSelThreads) score_hist_kernel(const float* __restrict__ cls_prob, int n_loc, int A,
                                                                 int N, uint32_t magic, uint32_t* __restrict__ hist) {
  extern __shared__ uint32_t s_hist[];
  const int f = blockIdx.y;
  for (int i = threadIdx.x; i < kSelectBins; i += kSelThreads) s_hist[i] = 0u;
  __syncthreads();
  const float* fp = cls_prob + (size_t)f * n_loc * (2 * A);
  const int n0 = blockIdx.x * kSelChunk, n1 = min(N, n0 + kSelChunk);
  for (int n = n0 + threadIdx.x; n < n1; n += kSelThreads)
    atomicAdd(&s_hist[score_key(fg_score_fast(fp, A, magic, n)) >> kSelectShift], 1u);
  __syncthreads();
  uint32_t* h = hist + (size_t)f * kSelectBins;
  for (int i = threadIdx.x; i < kSelectBins; i += kSelThreads) {
    const uint32_t c = s_hist[i];
    if (c) atomicAdd(h + i, c);
  }
}

// One CTA per frame: find the smallest set of top bins holding at least k scores.
__global__ void __launch_bounds__(1024) score_threshold_kernel(const uint32_t* __restrict__ hist,
                                                               uint32_t* __restrict__ sel, int k) {
  __shared__ uint32_t part[1024];
  const int f = blockIdx.x;
  const int t = threadIdx.x;
  const uint32_t* h = hist + (size_t)f * kSelectBins;
  constexpr int kPer = kSelectBins / 1024;
  uint32_t local = 0;
#pragma unroll 8
  for (int b = 0; b < kPer; ++b) local += h[t * kPer + b];
  part[t] = local;
  __syncthreads();
  // suffix sum over threads (inclusive), Hillis-Steele on 1024 entries
  for (int d = 1; d < 1024; d <<= 1) {
    uint32_t v = (t + d < 1024) ? part[t + d] : 0u;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  const uint32_t incl = part[t];           // scores in bins >= t*kPer
  const uint32_t above = incl - local;     // scores in bins owned by higher threads
  if (t == 0) {
    sel[f * 4 + 1] = 0u;
    sel[f * 4 + 2] = (uint32_t)k;
  }
  if (k > 0 && above < (uint32_t)k && incl >= (uint32_t)k) {
    uint32_t cum = above;
    for (int b = kPer - 1; b >= 0; --b) {
      const uint32_t c = h[t * kPer + b];
      if (cum + c >= (uint32_t)k) {
        sel[f * 4 + 0] = (uint32_t)(t * kPer + b);
        sel[f * 4 + 3] = cum;
        break;
      }
      cum += c;
    }
  }
  if (k <= 0 && t == 0) {
    sel[f * 4 + 0] = kSelectBins;  // nothing selected
    sel[f * 4 + 3] = 0u;
  }
}

// Candidates (scores in bins >= the threshold bin) -> composites, unordered.  A CTA counts its
// candidates first and reserves its output range with ONE global atomic.
// tiling measured on 128 Waymo frames (whole proposal stage, us): 256 x 16 x 1 tile 367, 512 x 8 x 4 347, 512 x 16 x 2 398,
// 1024 x 8 x 2 346, 512 x 4 x 8 358, 512 x 8 x 2 339
#define B2D_CMP_THREADS 512
#define B2D_CMP_ITER 8
#define B2D_CMP_TILES 2
constexpr int kCmpThreads = B2D_CMP_THREADS;
constexpr int kCmpIter = B2D_CMP_ITER;                         // scores per thread and tile
constexpr int kCmpTile = kCmpThreads * kCmpIter;
constexpr int kCmpChunk = kCmpTile * B2D_CMP_TILES;            // scores per CTA of the compact pass

__global__ void __launch_bounds__(kCmpThreads) score_compact_kernel(const float* __restrict__ cls_prob, int n_loc,
                                                                    int A, int N, uint32_t magic,
                                                                    uint32_t* __restrict__ sel,
                                                                    uint64_t* __restrict__ cand) {
  __shared__ uint32_t s_warp[kCmpThreads / 32];
  __shared__ uint32_t s_base;
  const int f = blockIdx.y;
  const uint32_t thr_bin = sel[f * 4 + 0];
  uint64_t* out = cand + (size_t)f * N;
  const float* fp = cls_prob + (size_t)f * n_loc * (2 * A);
  const int c0 = blockIdx.x * kCmpChunk, n1 = min(N, c0 + kCmpChunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int n0 = c0; n0 < n1; n0 += kCmpTile) {
    uint32_t keys[kCmpIter];
    uint32_t mine = 0;     // bit it: element it of this thread is a candidate
#pragma unroll
    for (int it = 0; it < kCmpIter; ++it) {
      const int n = n0 + it * kCmpThreads + threadIdx.x;
      keys[it] = 0;
      if (n < n1) {
        keys[it] = score_key(fg_score_fast(fp, A, magic, n));
        if ((keys[it] >> kSelectShift) >= thr_bin) mine |= 1u << it;
      }
    }
    // exclusive prefix of the per-thread counts over the CTA
    const uint32_t cnt = (uint32_t)__popc(mine);
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += v;
    }
    if (n0 > c0) __syncthreads();          // the previous tile's s_warp / s_base are dead
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = lane < kCmpThreads / 32 ? s_warp[lane] : 0u;
      uint32_t wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, d);
        if (lane >= d) wi += v;
      }
      if (lane < kCmpThreads / 32) s_warp[lane] = wi - w;
      if (lane == 31) s_base = wi ? atomicAdd(sel + f * 4 + 1, wi) : 0u;
    }
    __syncthreads();
    uint32_t pos = s_base + s_warp[warp] + incl - cnt;
#pragma unroll
    for (int it = 0; it < kCmpIter; ++it)
      if (mine & (1u << it)) out[pos++] = composite_key(keys[it], (uint32_t)(n0 + it * kCmpThreads + threadIdx.x));
  }
}

// ---------------------------------------------------------------------------------------
// Few frames per call (the reference API issues ONE): histogram, threshold and compaction in ONE cooperative launch
// with two grid-wide barriers, instead of memset + three kernels + memset whose launch gaps cost more than their work
// (46 us of a 190 us frame).  The grid is sized to be co-resident (cudaLaunchCooperativeKernel checks it); CTAs
// [f * ctas_per_frame, (f + 1) * ctas_per_frame) own frame f, each a contiguous strip of its scores.
//   phase 0  zero the frame's histogram and its counters
//   phase 1  histogram of the top kSelectBits key bits, privatised in shared memory (64 KB); only the non-zero bins
//            are merged with global atomics (scores pile up in a few bins: straight global atomics serialise on them -
//            42 us for one Waymo frame)
//   phase 2  every CTA finds the threshold bin of its frame on its own (512 threads x 32 bins, suffix sums), then
//            compacts its strip: CTA-wide prefix, ONE reservation atomic per 4096 scores
// The second pass over the scores hits L2.
constexpr int kFusedThreads = 512;
constexpr int kFusedIter = 8;                                  // scores per thread and compaction tile
constexpr int kFusedTile = kFusedThreads * kFusedIter;
static_assert(kSelectBins == kFusedThreads * 32, "threshold phase: 32 bins per thread");

__global__ void __launch_bounds__(kFusedThreads) select_fused_kernel(const float* __restrict__ cls_prob, int n_loc, int A,
                                                                     int N, uint32_t magic, int k, int ctas_per_frame,
                                                                     uint32_t* __restrict__ hist, uint32_t* __restrict__ sel,
                                                                     uint64_t* __restrict__ cand) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ uint32_t s_hist[];       // [kSelectBins]
  __shared__ uint32_t s_warp[kFusedThreads / 32];
  __shared__ uint32_t s_thr[2];
  __shared__ uint32_t s_base;
  const int f = blockIdx.x / ctas_per_frame, g = blockIdx.x - f * ctas_per_frame;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* h = hist + (size_t)f * kSelectBins;
  // ---- phase 0
  for (int i = g * kFusedThreads + tid; i < kSelectBins; i += ctas_per_frame * kFusedThreads) h[i] = 0u;
  if (g == 0 && tid == 0) {
    sel[f * 4 + 1] = 0u;
    sel[f * 4 + 2] = (uint32_t)k;
  }
  for (int i = tid; i < kSelectBins; i += kFusedThreads) s_hist[i] = 0u;
  // ---- phase 1 (the private part runs before the barrier: it does not touch the global histogram)
  const float* fp = cls_prob + (size_t)f * n_loc * (2 * A);
  const int chunk = (N + ctas_per_frame - 1) / ctas_per_frame;
  const int n0 = min(N, g * chunk), n1 = min(N, n0 + chunk);
  __syncthreads();
#pragma unroll 8
  for (int n = n0 + tid; n < n1; n += kFusedThreads)
    atomicAdd(&s_hist[score_key(fg_score_fast(fp, A, magic, n)) >> kSelectShift], 1u);
  grid.sync();
  for (int i = tid; i < kSelectBins; i += kFusedThreads) {
    const uint32_t c = s_hist[i];
    if (c) atomicAdd(h + i, c);
  }
  grid.sync();
  // ---- phase 2: threshold bin = the smallest set of top bins holding at least k scores (as score_threshold_kernel)
  {
    uint32_t c[32];
    const uint4* hv = reinterpret_cast<const uint4*>(h + tid * 32);
    uint32_t local = 0u;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 v = __ldcg(hv + q);
      c[4 * q] = v.x, c[4 * q + 1] = v.y, c[4 * q + 2] = v.z, c[4 * q + 3] = v.w;
      local += v.x + v.y + v.z + v.w;
    }
    uint32_t incl = local;                     // suffix sum over the lanes of the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_down_sync(0xFFFFFFFFu, incl, d);
      if (lane + d < 32) incl += v;
    }
    if (lane == 0) s_warp[warp] = incl;
    if (tid == 0) {
      s_thr[0] = kSelectBins;                  // k <= 0: nothing selected
      s_thr[1] = 0u;
    }
    __syncthreads();
    for (int w = warp + 1; w < kFusedThreads / 32; ++w) incl += s_warp[w];
    const uint32_t above = incl - local;       // scores in bins owned by higher threads
    if (k > 0 && above < (uint32_t)k && incl >= (uint32_t)k) {
      uint32_t cum = above;
#pragma unroll
      for (int b = 31; b >= 0; --b) {
        if (cum + c[b] >= (uint32_t)k) {
          s_thr[0] = (uint32_t)(tid * 32 + b);
          s_thr[1] = cum;
          break;
        }
        cum += c[b];
      }
    }
    __syncthreads();
  }
  const uint32_t thr_bin = s_thr[0];
  if (g == 0 && tid == 0) {
    sel[f * 4 + 0] = thr_bin;
    sel[f * 4 + 3] = s_thr[1];
  }
  // ---- compaction of the strip
  uint64_t* out = cand + (size_t)f * N;
  for (int t0 = n0; t0 < n1; t0 += kFusedTile) {
    uint32_t keys[kFusedIter];
    uint32_t mine = 0u;
#pragma unroll
    for (int it = 0; it < kFusedIter; ++it) {
      const int n = t0 + it * kFusedThreads + tid;
      keys[it] = 0u;
      if (n < n1) {
        keys[it] = score_key(fg_score_fast(fp, A, magic, n));
        if ((keys[it] >> kSelectShift) >= thr_bin) mine |= 1u << it;
      }
    }
    const uint32_t cnt = (uint32_t)__popc(mine);
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += v;
    }
    __syncthreads();                           // s_warp / s_base of the previous tile (and of the threshold phase) are dead
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const uint32_t w = lane < kFusedThreads / 32 ? s_warp[lane] : 0u;
      uint32_t wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, d);
        if (lane >= d) wi += v;
      }
      if (lane < kFusedThreads / 32) s_warp[lane] = wi - w;
      if (lane == 31) s_base = wi ? atomicAdd(sel + f * 4 + 1, wi) : 0u;
    }
    __syncthreads();
    uint32_t pos = s_base + s_warp[warp] + incl - cnt;
#pragma unroll
    for (int it = 0; it < kFusedIter; ++it)
      if (mine & (1u << it)) out[pos++] = composite_key(keys[it], (uint32_t)(t0 + it * kFusedThreads + tid));
  }
}

// co-resident CTAs of select_fused_kernel on the current device (0: cooperative launches unavailable)
static int fused_select_capacity() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev == cached_dev) return cached;
  int coop = 0, sms = 0, per_sm = 0;
  if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaFuncSetAttribute(select_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(sizeof(uint32_t) * kSelectBins)) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, select_fused_kernel, kFusedThreads,
                                                    sizeof(uint32_t) * kSelectBins) != cudaSuccess) {
    cudaGetLastError();
    sms = per_sm = 0;
  }
  cached_dev = dev;
  cached = sms * per_sm;
  return cached;
}

// ---------------------------------------------------------------------------------------
// In-CTA exact top-k + sort.  Dynamic smem: n_pad u64 keys (<= kMaxSortElems).
__device__ void bitonic_sort_desc(uint64_t* s, int n_pad) {
  const int half = n_pad >> 1;
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int p = threadIdx.x; p < half; p += blockDim.x) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const int q = i | j;
        const uint64_t a = s[i], b = s[q];
        const bool desc = (i & k) == 0;  // descending runs first => final order descending
        if ((a < b) == desc) {
          s[i] = b;
          s[q] = a;
        }
      }
      __syncthreads();
    }
  }
}

// The same network for n_pad = THREADS * E keys, with most of it off shared memory: thread t owns keys
// [t * E, (t + 1) * E), so the steps with stride j < E are register compare-exchanges, the steps with E <= j < 32 E
// exchange with a lane of the same warp by shuffle, and only the strides j >= 32 E go through shared memory.  For 8192
// keys that is 15 shared-memory steps with a CTA barrier instead of 91 (the barrier-per-step form was bound by
// shared-memory latency: 10 of 17 stall cycles per issued instruction in sort_decode_kernel).
template <int E, int THREADS>
__device__ void bitonic_sort_desc_regs(uint64_t* s) {
  constexpr int n_pad = THREADS * E;
  const int t = threadIdx.x, lane = t & 31, base = t * E;
  uint64_t v[E];
  auto load = [&]() {
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = s[base + e];
  };
  auto store = [&]() {
#pragma unroll
    for (int e = 0; e < E; ++e) s[base + e] = v[e];
  };
  // strides below E: both keys in this thread (J compile-time, k decides the direction of the pair's run)
  auto local_steps = [&](int k, int jmax) {
#pragma unroll
    for (int j = E / 2; j >= 1; j >>= 1) {
      if (j > jmax) continue;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if ((e & j) == 0) {
          const bool desc = ((base + e) & k) == 0;
          const uint64_t a = v[e], b = v[e | j];
          if ((a < b) == desc) {
            v[e] = b;
            v[e | j] = a;
          }
        }
      }
    }
  };
  // strides E * m, m < 32: the partner key sits in lane ^ m at the same register index
  auto shuffle_steps = [&](int k, int mmax) {
    for (int m = mmax; m >= 1; m >>= 1) {
      const bool lower = (lane & m) == 0;
      const bool desc = (base & k) == 0;               // (k >= 2 E here: the whole thread lies in one run)
      const bool want_max = lower == desc;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const uint64_t o = __shfl_xor_sync(0xFFFFFFFFu, v[e], m);
        v[e] = want_max ? (o > v[e] ? o : v[e]) : (o < v[e] ? o : v[e]);
      }
    }
  };
  load();
  // runs up to 32 E: registers and shuffles only
  for (int k = 2; k <= 32 * E; k <<= 1) {
    if (k >= 2 * E) shuffle_steps(k, k / (2 * E));
    local_steps(k, k / 2);
  }
  store();
  __syncthreads();
  for (int k = 64 * E; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j >= 32 * E; j >>= 1) {
      for (int p = t; p < n_pad / 2; p += THREADS) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const int q = i | j;
        const uint64_t a = s[i], b = s[q];
        const bool desc = (i & k) == 0;
        if ((a < b) == desc) {
          s[i] = b;
          s[q] = a;
        }
      }
      __syncthreads();
    }
    load();
    shuffle_steps(k, 16);
    local_steps(k, E / 2);
    store();
    __syncthreads();
  }
}

// n_pad keys, 1024 threads: the register / shuffle form when the list fills the CTA, the plain form otherwise
__device__ __forceinline__ void bitonic_sort_desc_cta(uint64_t* s, int n_pad) {
  if (blockDim.x == 1024 && n_pad == 16384) bitonic_sort_desc_regs<16, 1024>(s);
  else if (blockDim.x == 1024 && n_pad == 8192) bitonic_sort_desc_regs<8, 1024>(s);
  else if (blockDim.x == 1024 && n_pad == 4096) bitonic_sort_desc_regs<4, 1024>(s);
  else bitonic_sort_desc(s, n_pad);
}

// Slow path (more than kMaxSortElems candidates, e.g. massive score ties): exact MSB radix
// select of the k-th largest composite over the global candidate list; returns it.
__device__ uint64_t radix_select_kth(const uint64_t* __restrict__ cand, int m, int k, uint32_t* hist256,
                                     uint64_t* bcast) {
  uint64_t prefix = 0, mask = 0;
  int k_rem = k;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist256[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const uint64_t c = cand[i];
      if ((c & mask) == prefix) atomicAdd(&hist256[(c >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int cum = 0, d = 255;
      for (; d > 0; --d) {
        if (cum + (int)hist256[d] >= k_rem) break;
        cum += (int)hist256[d];
      }
      bcast[0] = (uint64_t)d;
      bcast[1] = (uint64_t)cum;
    }
    __syncthreads();
    prefix |= bcast[0] << shift;
    mask |= 0xFFull << shift;
    k_rem -= (int)bcast[1];
    __syncthreads();
  }
  return prefix;
}

struct DecodeArgs {
  const float* cls_prob;
  const float* bbox_pred;
  const float* info;     // [F][7]
  const float* anchors;  // [N][4]
  int n_loc, A, N;
};

__device__ __forceinline__ void decode_clip(const float* __restrict__ anc, const float* __restrict__ d,
                                            const float* __restrict__ info, float out[4]) {
  // bbox_transform_inv (bbox_transform.py:82-103) then clip_boxes (:252-255); every op rounded.
  const float x1 = anc[0], y1 = anc[1], x2 = anc[2], y2 = anc[3];
  const float w = fadd(fsub(x2, x1), 1.0f);
  const float h = fadd(fsub(y2, y1), 1.0f);
  const float diag = __fsqrt_rn(fadd(fmul(w, w), fmul(h, h)));
  const float cx = fadd(x1, fmul(0.5f, w));
  const float cy = fadd(y1, fmul(0.5f, h));
  const float pcx = fadd(fmul(d[0], diag), cx);
  const float pcy = fadd(fmul(d[1], diag), cy);
  const float pw = fmul(expf(d[2]), w);
  const float ph = fmul(expf(d[3]), h);
  const float hx = fmul(0.5f, pw), hy = fmul(0.5f, ph);
  const float xlo = info[0], xhi = fsub(info[1], 1.0f), ylo = info[2], yhi = fsub(info[3], 1.0f);
  out[0] = clampf(fsub(pcx, hx), xlo, xhi);
  out[1] = clampf(fsub(pcy, hy), ylo, yhi);
  out[2] = clampf(fadd(pcx, hx), xlo, xhi);
  out[3] = clampf(fadd(pcy, hy), ylo, yhi);
}

__global__ void __launch_bounds__(1024) sort_decode_kernel(DecodeArgs a, const uint32_t* __restrict__ sel,
                                                           const uint64_t* __restrict__ cand_all,
                                                           float* __restrict__ sorted_boxes,
                                                           float* __restrict__ sorted_scores,
                                                           int32_t* __restrict__ sorted_index,
                                                           int32_t* __restrict__ n_sorted, int k_cap,
                                                           int decode, int only_overflow) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  __shared__ uint32_t hist256[256];
  __shared__ uint64_t bcast[2];
  __shared__ int s_count;

  pdl_trigger();
  pdl_wait();
  const int f = blockIdx.x;
  const int m = (int)sel[f * 4 + 1];
  if (only_overflow && m <= kMaxSortElems) return;     // run_sort_kernel + rank_scatter_kernel sorted this frame
  const int k = min((int)sel[f * 4 + 2], m);
  const uint64_t* cand = cand_all + (size_t)f * a.N;

  int cnt;
  if (m <= kMaxSortElems) {
    for (int i = threadIdx.x; i < m; i += blockDim.x) keys[i] = cand[i];
    cnt = m;
  } else {
    const uint64_t kth = radix_select_kth(cand, m, k, hist256, bcast);
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const uint64_t c = cand[i];
      if (c >= kth) keys[atomicAdd(&s_count, 1)] = c;  // composites are unique: exactly k survive
    }
    __syncthreads();
    cnt = s_count;
  }
  int n_pad = 2;
  while (n_pad < cnt) n_pad <<= 1;
  for (int i = cnt + threadIdx.x; i < n_pad; i += blockDim.x) keys[i] = 0ull;  // below every real key
  __syncthreads();
  bitonic_sort_desc_cta(keys, n_pad);

  if (threadIdx.x == 0) n_sorted[f] = k;
  float* ob = sorted_boxes + (size_t)f * k_cap * 4;
  float* os = sorted_scores + (size_t)f * k_cap;
  int32_t* oi = sorted_index + (size_t)f * k_cap;
  const float* info = a.info + f * 7;
  for (int i = threadIdx.x; i < k_cap; i += blockDim.x) {
    float box[4] = {0.f, 0.f, 0.f, 0.f};
    float sc = 0.f;
    int idx = 0;
    if (i < k) {
      idx = (int)composite_index(keys[i]);
      sc = fg_score(a.cls_prob, f, a.n_loc, a.A, idx);
      if (decode) {
        const float* anc = a.anchors + (size_t)idx * 4;
        const float* d = a.bbox_pred + ((size_t)f * a.N + idx) * 4;
        const float an[4] = {__ldg(anc), __ldg(anc + 1), __ldg(anc + 2), __ldg(anc + 3)};
        const float dd[4] = {__ldg(d), __ldg(d + 1), __ldg(d + 2), __ldg(d + 3)};
        decode_clip(an, dd, info, box);
      }
    }
    ob[i * 4 + 0] = box[0];
    ob[i * 4 + 1] = box[1];
    ob[i * 4 + 2] = box[2];
    ob[i * 4 + 3] = box[3];
    os[i] = sc;
    oi[i] = idx;
  }
}

// Few-frame sort.  grid (runs, F): the candidate list is cut into runs of kRunLen composites, each bitonic-sorted
// (descending) in shared memory by its own CTA and written back in place.  CTAs beyond the frame's candidate count
// leave at once.
__global__ void __launch_bounds__(kRunLen / 2) run_sort_kernel(const uint32_t* __restrict__ sel, uint64_t* __restrict__ cand_all,
                                                               int N) {
  __shared__ uint64_t s_key[kRunLen];
  pdl_trigger();
  pdl_wait();
  const int f = blockIdx.y;
  const int m = (int)sel[f * 4 + 1];
  const int i0 = blockIdx.x * kRunLen;
  if (m > kMaxSortElems || i0 >= m) return;
  uint64_t* cand = cand_all + (size_t)f * N;
  for (int j = threadIdx.x; j < kRunLen; j += kRunLen / 2) s_key[j] = i0 + j < m ? cand[i0 + j] : 0ull;   // 0 < every key
  __syncthreads();
  bitonic_sort_desc_regs<2, kRunLen / 2>(s_key);
  for (int j = threadIdx.x; j < kRunLen; j += kRunLen / 2)
    if (i0 + j < m) cand[i0 + j] = s_key[j];
}

// rank -> position.  Candidate i sits at position p of sorted run r; its rank in the whole list = p + the number of
// larger composites in every other run, found by binary search (composites are unique): ~10 steps per run instead of
// a comparison with every candidate (rank by counting was 64 M comparisons and 20 us for one Waymo frame).  The
// searches of four runs advance together (independent load chains; the runs stay in L1).  Candidate i then goes to row
// rank of the sorted list (only the first k rows exist); decode + clip as in sort_decode_kernel.  Rows k .. k_cap-1
// (N < pre_nms) are zero-filled.
constexpr int kScatterThreads = 1024;                  // four lanes per candidate
constexpr int kScatterCands = kScatterThreads / 4;

__global__ void __launch_bounds__(kScatterThreads) rank_scatter_kernel(DecodeArgs a, const uint32_t* __restrict__ sel,
                                                                       const uint64_t* __restrict__ cand_all,
                                                                       float* __restrict__ sorted_boxes,
                                                                       float* __restrict__ sorted_scores,
                                                                       int32_t* __restrict__ sorted_index,
                                                                       int32_t* __restrict__ n_sorted, int k_cap, int decode) {
  pdl_trigger();
  pdl_wait();
  const int f = blockIdx.y;
  const int m = (int)sel[f * 4 + 1];
  if (m > kMaxSortElems) return;
  const int k = min((int)sel[f * 4 + 2], m);
  const int sub = threadIdx.x & 3;
  const int i = blockIdx.x * kScatterCands + (threadIdx.x >> 2);
  if (i == 0 && sub == 0) n_sorted[f] = k;
  float* ob = sorted_boxes + (size_t)f * k_cap * 4;
  float* os = sorted_scores + (size_t)f * k_cap;
  int32_t* oi = sorted_index + (size_t)f * k_cap;
  if (i >= k && i < k_cap) {          // tail rows of a frame with fewer than k_cap candidates
    if (sub == 0) {
      os[i] = 0.f;
      oi[i] = 0;
    }
    ob[i * 4 + sub] = 0.f;
  }
  if ((int)(blockIdx.x * kScatterCands) >= m) return;
  // every CTA stages the whole (sorted-in-runs) list in shared memory with ONE bulk copy: a single L2 round trip
  // instead of ~50 dependent ones per thread (the size is rounded up to 16 bytes: the extra element lies inside the
  // caller's workspace and is never used)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // a frame's list starts on an odd 8-byte word when N is odd and f is: the copy then starts one element earlier (the
  // last element of the previous frame's list) so that source and destination stay 16-byte aligned
  const uint64_t* src = cand_all + (size_t)f * a.N;
  const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) >> 3) & 1u;
  uint64_t* cand = reinterpret_cast<uint64_t*>(smem_raw) + mis;
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = (((uint32_t)m + mis) * 8u + 15u) & ~15u;
    mbar_expect_tx(&bar, bytes);
    bulk_g2s(smem_raw, src - mis, bytes, &bar);
  }
  mbar_wait(&bar, 0);
  const bool live = i < m;
  const uint64_t v = live ? cand[i] : 0ull;
  const int my_run = i / kRunLen;
  const int nruns = (m + kRunLen - 1) / kRunLen;
  // lane `sub` of the candidate's quad searches runs sub, sub + 4, ...; four searches advance together (independent
  // load chains), so a list of 16 runs costs each lane ONE round of ten steps
  int r = sub == 0 ? i - my_run * kRunLen : 0;
  for (int r0 = sub; r0 < nruns; r0 += 16) {
    int lo[4], hi[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int run = r0 + 4 * q;
      lo[q] = min(run * kRunLen, m);
      hi[q] = (live && run < nruns && run != my_run) ? min(lo[q] + kRunLen, m) : lo[q];      // empty range: contributes nothing
    }
    const int base[4] = {lo[0], lo[1], lo[2], lo[3]};
    // descending runs: first position whose element is not greater than v
#pragma unroll 1
    for (int step = 0; step < 10; ++step) {          // kRunLen = 512: at most 10 halvings
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (lo[q] < hi[q]) {
          const int mid = (lo[q] + hi[q]) >> 1;
          if (cand[mid] > v) lo[q] = mid + 1; else hi[q] = mid;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) r += lo[q] - base[q];
  }
  r += __shfl_xor_sync(0xFFFFFFFFu, r, 1);
  r += __shfl_xor_sync(0xFFFFFFFFu, r, 2);
  if (!live || r >= k || sub != 0) return;
  const int idx = (int)composite_index(v);
  float box[4] = {0.f, 0.f, 0.f, 0.f};
  if (decode) {
    const float* anc = a.anchors + (size_t)idx * 4;
    const float* d = a.bbox_pred + ((size_t)f * a.N + idx) * 4;
    const float an[4] = {__ldg(anc), __ldg(anc + 1), __ldg(anc + 2), __ldg(anc + 3)};
    const float dd[4] = {__ldg(d), __ldg(d + 1), __ldg(d + 2), __ldg(d + 3)};
    decode_clip(an, dd, a.info + f * 7, box);
  }
  *reinterpret_cast<float4*>(ob + (size_t)r * 4) = make_float4(box[0], box[1], box[2], box[3]);
  os[r] = fg_score(a.cls_prob, f, a.n_loc, a.A, idx);
  oi[r] = idx;
}

// ---------------------------------------------------------------------------------------
// Lists longer than the in-CTA capacity (pre_nms_topN > 16384 or <= 0, argsort / nms of more than 16384 boxes):
// bitonic-sort chunks of kMaxSortElems in shared memory, then merge runs pairwise.  A merge pass is one thread per
// element: its position in the merged run = its position in its own run + the number of elements of the partner
// run that precede it (binary search; composites are unique, so there are no ties to break).
__global__ void __launch_bounds__(1024) chunk_sort_kernel(uint64_t* __restrict__ keys_all, const uint32_t* __restrict__ sel,
                                                          int stride, int n_fixed) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  const int f = blockIdx.y;
  const int m = sel ? (int)sel[f * 4 + 1] : n_fixed;
  const int c0 = blockIdx.x * kMaxSortElems;
  if (c0 >= m) return;
  uint64_t* g = keys_all + (size_t)f * stride + c0;
  const int cnt = min(kMaxSortElems, m - c0);
  int n_pad = 2;
  while (n_pad < cnt) n_pad <<= 1;
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x) keys[i] = i < cnt ? g[i] : 0ull;
  __syncthreads();
  bitonic_sort_desc_cta(keys, n_pad);
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) g[i] = keys[i];
}

__global__ void __launch_bounds__(256) merge_pass_kernel(const uint64_t* __restrict__ src_all, uint64_t* __restrict__ dst_all,
                                                         const uint32_t* __restrict__ sel, int stride, int n_fixed,
                                                         int run) {
  const int f = blockIdx.y;
  const int m = sel ? (int)sel[f * 4 + 1] : n_fixed;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint64_t* src = src_all + (size_t)f * stride;
  uint64_t* dst = dst_all + (size_t)f * stride;
  const int r = i / run, own = i - r * run;
  const int p0 = (r ^ 1) * run, p1 = min(p0 + run, m);        // partner run
  const uint64_t v = src[i];
  int cnt = 0;
  if (p0 < m) {
    // descending runs: number of partner elements greater than v
    int lo = p0, hi = p1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (src[mid] > v) lo = mid + 1; else hi = mid;
    }
    cnt = lo - p0;
  }
  dst[(size_t)(r & ~1) * run + own + cnt] = v;
}

// sorted composites -> the first k rows of the sorted list, decoded + clipped (rows k .. k_cap-1 zero-filled)
__global__ void __launch_bounds__(256) sorted_decode_kernel(DecodeArgs a, const uint32_t* __restrict__ sel,
                                                            const uint64_t* __restrict__ sorted_all,
                                                            float* __restrict__ sorted_boxes,
                                                            float* __restrict__ sorted_scores,
                                                            int32_t* __restrict__ sorted_index,
                                                            int32_t* __restrict__ n_sorted, int k_cap, int decode) {
  const int f = blockIdx.y;
  const int m = (int)sel[f * 4 + 1];
  const int k = min((int)sel[f * 4 + 2], m);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) n_sorted[f] = k;
  if (i >= k_cap) return;
  float box[4] = {0.f, 0.f, 0.f, 0.f};
  float sc = 0.f;
  int idx = 0;
  if (i < k) {
    idx = (int)composite_index(sorted_all[(size_t)f * a.N + i]);
    sc = fg_score(a.cls_prob, f, a.n_loc, a.A, idx);
    if (decode) {
      const float* anc = a.anchors + (size_t)idx * 4;
      const float* d = a.bbox_pred + ((size_t)f * a.N + idx) * 4;
      const float an[4] = {__ldg(anc), __ldg(anc + 1), __ldg(anc + 2), __ldg(anc + 3)};
      const float dd[4] = {__ldg(d), __ldg(d + 1), __ldg(d + 2), __ldg(d + 3)};
      decode_clip(an, dd, a.info + f * 7, box);
    }
  }
  float* ob = sorted_boxes + ((size_t)f * k_cap + i) * 4;
  ob[0] = box[0], ob[1] = box[1], ob[2] = box[2], ob[3] = box[3];
  sorted_scores[(size_t)f * k_cap + i] = sc;
  sorted_index[(size_t)f * k_cap + i] = idx;
}

// Sorts the first m (device count, or n_fixed) composites of every frame's list, descending; returns the buffer
// that holds the result (a or b).
static int sort_large(int F, int n_max, int stride, uint64_t* a, uint64_t* b, const uint32_t* sel, int n_fixed,
                      cudaStream_t st, uint64_t** result) {
  const size_t smem = sizeof(uint64_t) * kMaxSortElems;
  B2D_CUDA(cudaFuncSetAttribute(chunk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  chunk_sort_kernel<<<dim3(ceil_div(n_max, kMaxSortElems), F), 1024, smem, st>>>(a, sel, stride, n_fixed);
  B2D_LAUNCHED();
  uint64_t *src = a, *dst = b;
  for (long long run = kMaxSortElems; run < n_max; run *= 2) {
    merge_pass_kernel<<<dim3(ceil_div(n_max, 256), F), 256, 0, st>>>(src, dst, sel, stride, n_fixed, (int)run);
    B2D_LAUNCHED();
    uint64_t* t = src;
    src = dst;
    dst = t;
  }
  *result = src;
  return B2D_OK;
}

// keep positions -> output rows (proposal_layer.py:48-55); pads the tail with zeros.
__global__ void __launch_bounds__(256) proposal_gather_kernel(
    const float* __restrict__ sorted_boxes, const float* __restrict__ sorted_scores,
    const int32_t* __restrict__ sorted_index, const int32_t* __restrict__ keep,
    const int32_t* __restrict__ num_keep, const float* __restrict__ anchors_3d, int k_cap, int max_out,
    int batch_index_stride, float* __restrict__ rois, float* __restrict__ roi_scores,
    float* __restrict__ roi_a3d, int32_t* __restrict__ roi_anchor, int32_t* __restrict__ num_out) {
  pdl_trigger();
  pdl_wait();
  const int f = blockIdx.y;
  const int nk = num_keep[f];
  if (blockIdx.x == 0 && threadIdx.x == 0) num_out[f] = nk;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max_out; i += gridDim.x * blockDim.x) {
    float r[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    float sc = 0.f;
    int idx = -1;
    if (i < nk) {
      const int p = keep[(size_t)f * max_out + i];
      const float* b = sorted_boxes + ((size_t)f * k_cap + p) * 4;
      r[0] = (float)(f * batch_index_stride);
      r[1] = b[0];
      r[2] = b[1];
      r[3] = b[2];
      r[4] = b[3];
      sc = sorted_scores[(size_t)f * k_cap + p];
      idx = sorted_index[(size_t)f * k_cap + p];
    }
    float* o = rois + ((size_t)f * max_out + i) * 5;
#pragma unroll
    for (int c = 0; c < 5; ++c) o[c] = r[c];
    roi_scores[(size_t)f * max_out + i] = sc;
    if (roi_anchor) roi_anchor[(size_t)f * max_out + i] = idx;
    if (roi_a3d) {
      float* o3 = roi_a3d + ((size_t)f * max_out + i) * 7;
#pragma unroll
      for (int c = 0; c < 7; ++c) o3[c] = (idx >= 0 && anchors_3d) ? __ldg(anchors_3d + (size_t)idx * 7 + c) : 0.f;
    }
  }
}

// proposal_top_layer output rows: rois + scores + the selected anchors.
__global__ void __launch_bounds__(256) proposal_top_gather_kernel(
    const float* __restrict__ sorted_boxes, const float* __restrict__ sorted_scores,
    const int32_t* __restrict__ sorted_index, const float* __restrict__ anchors, int top_n,
    int batch_index_stride, float* __restrict__ rois, float* __restrict__ roi_scores,
    float* __restrict__ roi_anchors) {
  const int f = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < top_n; i += gridDim.x * blockDim.x) {
    const float* b = sorted_boxes + ((size_t)f * top_n + i) * 4;
    float* o = rois + ((size_t)f * top_n + i) * 5;
    o[0] = (float)(f * batch_index_stride);
    o[1] = b[0];
    o[2] = b[1];
    o[3] = b[2];
    o[4] = b[3];
    roi_scores[(size_t)f * top_n + i] = sorted_scores[(size_t)f * top_n + i];
    const int idx = sorted_index[(size_t)f * top_n + i];
#pragma unroll
    for (int c = 0; c < 4; ++c) roi_anchors[((size_t)f * top_n + i) * 4 + c] = __ldg(anchors + (size_t)idx * 4 + c);
  }
}

__global__ void __launch_bounds__(1024) argsort_desc_kernel(const float* __restrict__ scores, int n,
                                                            int32_t* __restrict__ order) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  const int f = blockIdx.x;
  int n_pad = 2;
  while (n_pad < n) n_pad <<= 1;
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x)
    keys[i] = i < n ? composite_key(score_key(scores[(size_t)f * n + i]), (uint32_t)i) : 0ull;
  __syncthreads();
  bitonic_sort_desc_cta(keys, n_pad);
  for (int i = threadIdx.x; i < n; i += blockDim.x) order[(size_t)f * n + i] = (int32_t)composite_index(keys[i]);
}

// ---------------------------------------------------------------------------------------
static int select_sort(int F, int n_loc, int A, const float* cls_prob, const float* bbox_pred, const float* info,
                       const float* anchors, int k, int decode, const ProposalWs& w, cudaStream_t st) {
  const int N = n_loc * A;
  const uint32_t magic = A > 1 ? (uint32_t)(0x100000000ull / (uint32_t)A) : 0xFFFFFFFFu;   // floor(2^32 / A): quotient low by at most 1
  const bool few = F <= kRankMaxFrames;
  bool fused = false;
#ifdef B2D_AB_NOFUSED
  if (false) {      // A/B timing build
#else
  if (few) {        // (for many frames the three-kernel path is as fast: 369 vs 412 us at 128 frames, 226 vs 215 at 32)
#endif
    // one cooperative launch: about 4 scores per thread, at most what the device holds at once
    const int cap = fused_select_capacity() / F;
    int cpf = ceil_div(N, kFusedThreads * 4);
    if (cpf > cap) cpf = cap;
    if (cpf >= 1) {
      const float* a0 = cls_prob;
      int a1 = n_loc, a2 = A, a3 = N, a5 = k, a6 = cpf;
      uint32_t a4 = magic;
      uint32_t *a7 = w.hist, *a8 = w.sel;
      uint64_t* a9 = w.cand;
      void* args[] = {&a0, &a1, &a2, &a3, &a4, &a5, &a6, &a7, &a8, &a9};
      if (cudaLaunchCooperativeKernel(reinterpret_cast<void*>(select_fused_kernel), dim3(cpf * F), dim3(kFusedThreads), args,
                                      sizeof(uint32_t) * kSelectBins, st) == cudaSuccess) {
        fused = true;
        B2D_LAUNCHED();
      } else {
        cudaGetLastError();                            // (too large for the device after all: the three-kernel path below)
      }
    }
  }
  if (!fused) {
    B2D_CUDA(cudaMemsetAsync(w.hist, 0, sizeof(uint32_t) * (size_t)F * kSelectBins, st));
    dim3 grid(ceil_div(N, kSelChunk), F);
    const size_t hsmem = sizeof(uint32_t) * kSelectBins;
    B2D_CUDA(cudaFuncSetAttribute(score_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
    score_hist_kernel<<<grid, kSelThreads, hsmem, st>>>(cls_prob, n_loc, A, N, magic, w.hist);
    B2D_LAUNCHED();
    score_threshold_kernel<<<F, 1024, 0, st>>>(w.hist, w.sel, k);
    B2D_LAUNCHED();
    dim3 cgrid(ceil_div(N, kCmpChunk), F);
    score_compact_kernel<<<cgrid, kCmpThreads, 0, st>>>(cls_prob, n_loc, A, N, magic, w.sel, w.cand);
    B2D_LAUNCHED();
  }
  const size_t smem = sizeof(uint64_t) * kMaxSortElems;
  B2D_CUDA(cudaFuncSetAttribute(sort_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DecodeArgs a{cls_prob, bbox_pred, info, anchors, n_loc, A, N};
  if (k > kMaxSortElems) {
    // more boxes wanted than one CTA can sort: chunk sort + merge passes over the whole candidate list
    uint64_t* sorted = nullptr;
    const int rc = sort_large(F, N, N, w.cand, w.cand2, w.sel, 0, st, &sorted);
    if (rc != B2D_OK) return rc;
    sorted_decode_kernel<<<dim3(ceil_div(k, 256), F), 256, 0, st>>>(a, w.sel, sorted, w.sorted_boxes, w.sorted_scores,
                                                                   w.sorted_index, w.n_sorted, k, decode);
    B2D_LAUNCHED();
    return B2D_OK;
  }
  if (few) {
    // the candidate count is only known on the device: size the grid for the largest list the path takes
    // (CTAs beyond the count leave at once); a frame with more candidates (massive score ties) is left to the
    // radix-select slow path of sort_decode_kernel below.  The three kernels are chained by programmatic dependent
    // launches (each waits for its predecessor on the device before it reads anything).
    const int cap = min(N, kMaxSortElems);
    B2D_CUDA(launch_pdl(run_sort_kernel, dim3(ceil_div(cap, kRunLen), F), dim3(kRunLen / 2), 0, st, true, w.sel, w.cand, N));
    B2D_LAUNCHED();
    dim3 sgrid(ceil_div(max(cap, k), kScatterCands), F);
    const size_t rsmem = sizeof(uint64_t) * (size_t)cap + 16;
    B2D_CUDA(cudaFuncSetAttribute(rank_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
    B2D_CUDA(launch_pdl(rank_scatter_kernel, sgrid, dim3(kScatterThreads), rsmem, st, true, a, w.sel, w.cand, w.sorted_boxes,
                        w.sorted_scores, w.sorted_index, w.n_sorted, k, decode));
    B2D_LAUNCHED();
    if (N <= kMaxSortElems) return B2D_OK;             // the candidate list can never overflow
  }
  B2D_CUDA(launch_pdl(sort_decode_kernel, dim3(F), dim3(1024), smem, st, few, a, w.sel, w.cand, w.sorted_boxes,
                      w.sorted_scores, w.sorted_index, w.n_sorted, k, decode, few ? 1 : 0));
  B2D_LAUNCHED();
  return B2D_OK;
}

}  // namespace b2d

using namespace b2d;

extern "C" size_t b2d_proposal_workspace_bytes(int num_frames, int n_loc, int num_anchors, int pre_nms,
                                               int post_nms) {
  if (num_frames <= 0 || n_loc <= 0 || num_anchors <= 0) return 0;
  return carve(nullptr, num_frames, n_loc * num_anchors, pre_nms, post_nms).bytes;
}

extern "C" int b2d_max_pre_nms(void) { return kMaxSortElems; }

extern "C" int b2d_proposal(int F, int n_loc, int A, const float* cls_prob, const float* bbox_pred,
                            const float* info, const float* anchors, const float* anchors_3d, int pre_nms,
                            int post_nms, double nms_thresh, int batch_index_stride, float* rois,
                            float* roi_scores, float* roi_a3d, int32_t* roi_anchor, int32_t* num_out,
                            void* workspace, size_t workspace_bytes, void* stream) {
  if (F <= 0 || n_loc <= 0 || A <= 0 || !cls_prob || !bbox_pred || !info || !anchors || !rois || !roi_scores ||
      !num_out)
    return B2D_ERR_INVALID_ARG;
  if ((long long)n_loc * A > 0x7FFFFFFFll / 8) return B2D_ERR_UNSUPPORTED;
  const int N = n_loc * A;
  const int k = top_k_of(N, pre_nms);
  const int mo = max_out_of(N, pre_nms, post_nms);
  ProposalWs w = carve(workspace, F, N, pre_nms, post_nms);
  if (!workspace || workspace_bytes < w.bytes) return B2D_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  int rc = select_sort(F, n_loc, A, cls_prob, bbox_pred, info, anchors, k, 1, w, st);
  if (rc != B2D_OK) return rc;
  rc = launch_nms_sorted(F, k, w.sorted_boxes, w.n_sorted, nms_thresh, mo, w.keep, w.num_keep, w.nms_kept, st);
  if (rc != B2D_OK) return rc;
  dim3 grid(ceil_div(mo, 256), F);
  B2D_CUDA(launch_pdl(proposal_gather_kernel, grid, dim3(256), 0, st, F <= kRankMaxFrames, w.sorted_boxes, w.sorted_scores,
                      w.sorted_index, w.keep, w.num_keep, anchors_3d, k, mo, batch_index_stride, rois, roi_scores, roi_a3d,
                      roi_anchor, num_out));
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_proposal_top(int F, int n_loc, int A, const float* cls_prob, const float* bbox_pred,
                                const float* info, const float* anchors, int top_n, int batch_index_stride,
                                float* rois, float* roi_scores, float* roi_anchors, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (F <= 0 || n_loc <= 0 || A <= 0 || !cls_prob || !bbox_pred || !info || !anchors || !rois || !roi_scores ||
      !roi_anchors || top_n <= 0)
    return B2D_ERR_INVALID_ARG;
  const int N = n_loc * A;
  if (top_n > N) return B2D_ERR_UNSUPPORTED;  // the reference's random-fill branch stays on the host
  ProposalWs w = carve(workspace, F, N, top_n, top_n);
  if (!workspace || workspace_bytes < w.bytes) return B2D_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  int rc = select_sort(F, n_loc, A, cls_prob, bbox_pred, info, anchors, top_n, 1, w, st);
  if (rc != B2D_OK) return rc;
  dim3 grid(ceil_div(top_n, 256), F);
  proposal_top_gather_kernel<<<grid, 256, 0, st>>>(w.sorted_boxes, w.sorted_scores, w.sorted_index, anchors, top_n,
                                                   batch_index_stride, rois, roi_scores, roi_anchors);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_proposal_debug_sorted(int F, int n_loc, int A, int pre_nms, int post_nms, const void* workspace,
                                         float* sorted_boxes, float* sorted_scores, int32_t* sorted_index,
                                         void* stream) {
  if (!workspace || F <= 0) return B2D_ERR_INVALID_ARG;
  const int N = n_loc * A;
  const int k = top_k_of(N, pre_nms);
  ProposalWs w = carve(const_cast<void*>(workspace), F, N, pre_nms, post_nms);
  cudaStream_t st = as_stream(stream);
  if (sorted_boxes)
    B2D_CUDA(cudaMemcpyAsync(sorted_boxes, w.sorted_boxes, sizeof(float) * (size_t)F * k * 4,
                             cudaMemcpyDeviceToDevice, st));
  if (sorted_scores)
    B2D_CUDA(cudaMemcpyAsync(sorted_scores, w.sorted_scores, sizeof(float) * (size_t)F * k,
                             cudaMemcpyDeviceToDevice, st));
  if (sorted_index)
    B2D_CUDA(cudaMemcpyAsync(sorted_index, w.sorted_index, sizeof(int32_t) * (size_t)F * k,
                             cudaMemcpyDeviceToDevice, st));
  return B2D_OK;
}

namespace b2d {
__global__ void __launch_bounds__(256) argsort_keys_kernel(const float* __restrict__ scores, int n, uint64_t* __restrict__ keys) {
  const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[(size_t)f * n + i] = composite_key(score_key(scores[(size_t)f * n + i]), (uint32_t)i);
}
__global__ void __launch_bounds__(256) argsort_index_kernel(const uint64_t* __restrict__ keys, int n, int32_t* __restrict__ order) {
  const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) order[(size_t)f * n + i] = (int32_t)composite_index(keys[(size_t)f * n + i]);
}
}  // namespace b2d

extern "C" size_t b2d_argsort_workspace_bytes(int F, int n) {
  if (F <= 0 || n <= kMaxSortElems) return 0;
  return 2 * align_up(sizeof(uint64_t) * (size_t)F * n, 256);
}

extern "C" int b2d_argsort_desc(int F, int n, const float* scores, int32_t* order, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (F <= 0 || n < 0 || !scores || !order) return B2D_ERR_INVALID_ARG;
  if (n == 0) return B2D_OK;
  if (n > kMaxSortElems) {
    const size_t half = align_up(sizeof(uint64_t) * (size_t)F * n, 256);
    if (!workspace || workspace_bytes < 2 * half) return B2D_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    uint64_t* a = static_cast<uint64_t*>(workspace);
    uint64_t* b = reinterpret_cast<uint64_t*>(static_cast<char*>(workspace) + half);
    argsort_keys_kernel<<<dim3(ceil_div(n, 256), F), 256, 0, st>>>(scores, n, a);
    B2D_LAUNCHED();
    uint64_t* sorted = nullptr;
    const int rc = sort_large(F, n, n, a, b, nullptr, n, st, &sorted);
    if (rc != B2D_OK) return rc;
    argsort_index_kernel<<<dim3(ceil_div(n, 256), F), 256, 0, st>>>(sorted, n, order);
    B2D_LAUNCHED();
    return B2D_OK;
  }
  const size_t smem = sizeof(uint64_t) * kMaxSortElems;
  B2D_CUDA(cudaFuncSetAttribute(argsort_desc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int n_pad = 2;
  while (n_pad < n) n_pad <<= 1;
  argsort_desc_kernel<<<F, 1024, sizeof(uint64_t) * n_pad, as_stream(stream)>>>(scores, n, order);
  B2D_LAUNCHED();
  return B2D_OK;
}
