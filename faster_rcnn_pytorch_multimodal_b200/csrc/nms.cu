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
constexpr int kNmsClusterMaxFrames = 8;      // up to this many frames per call take the cluster kernel (one GPC each)

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

    // ---- C: resolve (first warp).  kept(i) = alive(i) and no kept j < i suppresses i defines the keep set uniquely;
    // iterating  K <- alive & ~(OR of the rows of K)  from K = alive fixes at least one more leading candidate per
    // round and stops at the fixed point (typically 3-5 rounds of two warp-wide OR reductions; a serial walk by one
    // thread costs ~100 cycles per kept box).  The keep list and the kept boxes are then written in parallel.
    if (tid < 32) {
      unsigned long long alive = ~(((unsigned long long)s_dead[1] << 32) | s_dead[0]);
      if (tile_n < 64) alive &= (1ull << tile_n) - 1ull;
      const unsigned long long r0 = s_rows[tid], r1 = s_rows[tid + 32];     // (rows of dead / absent candidates are never selected)
      unsigned long long kept_mask = alive;
      for (int it = 0; it < 64; ++it) {
        const unsigned long long c = (((kept_mask >> tid) & 1ull) ? r0 : 0ull) | (((kept_mask >> (tid + 32)) & 1ull) ? r1 : 0ull);
        const unsigned lo = __reduce_or_sync(0xFFFFFFFFu, (unsigned)c), hi = __reduce_or_sync(0xFFFFFFFFu, (unsigned)(c >> 32));
        const unsigned long long nxt = alive & ~(((unsigned long long)hi << 32) | lo);
        if (nxt == kept_mask) break;
        kept_mask = nxt;
      }
      int nt = __popcll(kept_mask);
      while (nt > max_keep - K) {                                          // post_nms reached inside the tile
        kept_mask &= ~(1ull << (63 - __clzll((long long)kept_mask)));
        --nt;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = tid + 32 * h;
        if ((kept_mask >> i) & 1ull) {
          const int pos = K + __popcll(kept_mask & ((1ull << i) - 1ull));
          keep[pos] = base + i;
          if (pos < kKeptSmem) s_kept[pos] = s_tile[i];
        }
      }
      if (tid == 0) s_k = K + nt;
    }
    __syncthreads();
  }
  if (tid == 0) num_keep[f] = s_k;
}


// ------------------------------------------------------------------------------------------
// Few-frame path: one thread-block CLUSTER (16 CTAs x 1024 threads) per frame instead of one CTA.
// The single-CTA kernel above is bound by one SM's issue rate (12 M IoUs for a 12000 -> 2000 train frame) and by
// its serial per-tile resolve; with one frame per call - what the reference API issues - 147 SMs idle meanwhile.
// Candidates are taken in chunks of 512.  Per chunk:
//   A (all 16 CTAs, 32 candidates each, ONE WARP per candidate): the candidate against the kept list so far
//     (global memory, written by the sweeper; lanes stride over it) -> dead bit; and its row of the chunk's
//     512 x 512 suppression bitmask (8 words of 64 bits, columns above the diagonal only): lane l evaluates
//     columns l, l + 32, ... and two ballots assemble each word.  Rows and dead flags are written straight into
//     CTA 0's shared memory through DSMEM (st.shared::cluster).
//   cluster barrier
//   B (one warp of CTA 0): greedy sweep.  Lane w < 8 holds the "removed" word of column tile w in a register; a tile
//     of 64 candidates is resolved by RUNS (all candidates up to the next one that a live candidate suppresses are
//     kept at once: two warp-wide OR reductions per run), then the rows of the kept boxes are ORed into the removed
//     words and the keep list is written in parallel - no __syncthreads, no global bitmask.
//   CTA 0 appends the kept boxes to the global kept list and publishes (count, done); cluster barrier.
// Same decisions as the single-CTA kernel: both evaluate iou_exceeds() on the same operands.
constexpr int kClusterCtas = 16;                     // non-portable cluster size (B200 allows 16 with the opt-in)
constexpr int kChunk = 512;                          // candidates per chunk
constexpr int kChunkWords = kChunk / 64;             // 8
constexpr int kRowsPerCta = kChunk / kClusterCtas;   // 32 = warps per CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_u64(uint32_t addr, unsigned long long v) {
  asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

struct NmsClusterSmem {
  unsigned long long mask[kChunk][kChunkWords];   // CTA 0: suppression rows of the chunk (32 KB)
  float4 kept[kKeptSmem];                          // every CTA: its copy of the kept list (the rest stays in L2)
  float4 box[kChunk];                              // the chunk's boxes (every CTA loads them)
  int dead[kChunk];                                // CTA 0: candidate is suppressed by an earlier chunk's kept box
  int newk[kChunk];                                // CTA 0: chunk rows kept by this sweep
  int state[2];                                    // {kept count, done}: written by CTA 0 into every CTA
};

__global__ void __launch_bounds__(kNmsThreads, 1)
nms_cluster_kernel(const float4* __restrict__ boxes_all, int n, const int32_t* __restrict__ n_valid, float thr_f,
                   int max_keep, int32_t* __restrict__ keep_all, int32_t* __restrict__ num_keep,
                   float4* __restrict__ kept_boxes_all) {
  extern __shared__ __align__(16) unsigned char nms_smem_raw[];
  NmsClusterSmem& S = *reinterpret_cast<NmsClusterSmem*>(nms_smem_raw);
  pdl_trigger();
  pdl_wait();
  const int f = blockIdx.y;
  const uint32_t cta = cluster_ctarank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* boxes = boxes_all + (size_t)f * n;
  int32_t* keep = keep_all + (size_t)f * max_keep;
  float4* kept_boxes = kept_boxes_all + (size_t)f * max_keep;
  const int nv = n_valid ? min(n_valid[f], n) : n;
  if (tid < 2) S.state[tid] = 0;
  cluster_sync_all();
  const uint32_t mask0 = map_to_cta(smem_u32(&S.mask[0][0]), 0);
  const uint32_t dead0 = map_to_cta(smem_u32(&S.dead[0]), 0);

  int K_cached = 0;                                   // kept boxes already copied into S.kept
  for (int base = 0; base < nv; base += kChunk) {
    const int K = S.state[0];
    const int cn = min(kChunk, nv - base);            // candidates in this chunk
    if (tid < kChunk) S.box[tid] = tid < cn ? boxes[base + tid] : make_float4(0.f, 0.f, 0.f, 0.f);
    // the boxes the last sweep kept: one coalesced L2 read instead of a dependent load per 32 comparisons
    for (int q = K_cached + tid; q < min(K, kKeptSmem); q += kNmsThreads) S.kept[q] = __ldcg(kept_boxes + q);
    K_cached = K;
    __syncthreads();
    // ---- A: this CTA's 32 candidates, one warp each
    const int row = (int)cta * kRowsPerCta + warp;   // candidate within the chunk
    if (row < cn) {
      const float4 me = S.box[row];
      const float my_area = fmul(fsub(me.z, me.x), fsub(me.w, me.y));
      bool dead = false;
      for (int k0 = 0; k0 < K; k0 += 32) {
        const int kk = k0 + lane;
        const bool hit = kk < K && iou_exceeds(me, my_area, kk < kKeptSmem ? S.kept[kk] : __ldcg(kept_boxes + kk), thr_f);
        if (__any_sync(0xFFFFFFFFu, hit)) {
          dead = true;
          break;
        }
      }
      if (lane == 0) asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(dead0 + (uint32_t)row * 4u), "r"(dead ? 1 : 0) : "memory");
      if (!dead) {
        // (a dead candidate's row is never read by the sweep)
        unsigned long long mine = 0ull;              // lane w < 8 ends up with word w
#pragma unroll
        for (int w = 0; w < kChunkWords; ++w) {
          unsigned lo = 0u, hi = 0u;
          if (64 * w + 63 > row) {                   // warp-uniform: the word has columns above the diagonal
            const int j0 = 64 * w + lane, j1 = j0 + 32;
            lo = __ballot_sync(0xFFFFFFFFu, j0 > row && j0 < cn && iou_exceeds(me, my_area, S.box[j0], thr_f));
            hi = __ballot_sync(0xFFFFFFFFu, j1 > row && j1 < cn && iou_exceeds(me, my_area, S.box[j1], thr_f));
          }
          if (lane == w) mine = ((unsigned long long)hi << 32) | lo;
        }
        if (lane < kChunkWords) st_cluster_u64(mask0 + (uint32_t)(row * kChunkWords + lane) * 8u, mine);
      }
    }
    cluster_sync_all();
    // ---- B: sweep (first warp of CTA 0)
    if (cta == 0 && warp == 0) {
      // lane w < 8: "removed" word of column tile w, seeded with the candidates the kept list already suppresses
      unsigned long long rem = 0ull;
#pragma unroll
      for (int w = 0; w < kChunkWords; ++w) {
        const int j0 = 64 * w + lane, j1 = j0 + 32;
        const unsigned lo = __ballot_sync(0xFFFFFFFFu, j0 < cn && S.dead[j0] != 0);
        const unsigned hi = __ballot_sync(0xFFFFFFFFu, j1 < cn && S.dead[j1] != 0);
        if (lane == w) rem = ((unsigned long long)hi << 32) | lo;
      }
      int k = K;
      for (int t = 0; t < kChunkWords && k < max_keep; ++t) {
        const int tile_n = min(64, cn - t * 64);
        if (tile_n <= 0) break;
        unsigned long long a = ~__shfl_sync(0xFFFFFFFFu, rem, t);       // undecided candidates of the tile
        if (tile_n < 64) a &= (1ull << tile_n) - 1ull;
        // this lane's two rows of the tile's 64 x 64 diagonal block (rows of dead candidates hold stale data and are
        // only ever selected through an alive bit)
        const unsigned long long d0 = S.mask[t * 64 + lane][t], d1 = S.mask[t * 64 + 32 + lane][t];
        auto or_rows = [&](unsigned long long sel) {                      // OR of the block rows named by `sel`
          const unsigned long long c = (((sel >> lane) & 1ull) ? d0 : 0ull) | (((sel >> (lane + 32)) & 1ull) ? d1 : 0ull);
          const unsigned lo = __reduce_or_sync(0xFFFFFFFFu, (unsigned)c), hi = __reduce_or_sync(0xFFFFFFFFu, (unsigned)(c >> 32));
          return ((unsigned long long)hi << 32) | lo;
        };
        // Greedy resolve of the tile as a fixed point instead of a walk.  kept(i) = a(i) and no kept j < i suppresses i
        // defines the keep set uniquely; iterating  K <- a & ~(OR of the block rows of K)  from K = a fixes at least one
        // more leading candidate per round (the first n candidates are final after n rounds) and stops at the fixed
        // point - typically 3-5 rounds of two warp-wide OR reductions, against 60 kept boxes x 85 cycles for the walk
        // (a lone warp retires a dependent instruction every ~6 cycles: the walk and the box-by-box update of the
        // removed words were 32 of the kernel's 40 us on one Waymo frame).
        unsigned long long kept_mask = a;
        for (int it = 0; it < 64; ++it) {
          const unsigned long long nxt = a & ~or_rows(kept_mask);
          if (nxt == kept_mask) break;
          kept_mask = nxt;
        }
        {
          const int room = max_keep - k;
          int nt = __popcll(kept_mask);
          while (nt > room) {                                            // post_nms reached inside the tile
            kept_mask &= ~(1ull << (63 - __clzll((long long)kept_mask)));
            --nt;
          }
        }
        // rows of the kept boxes -> removed words of the later tiles: every lane contributes its two rows, one warp-wide
        // OR per word
        {
          const bool k0 = (kept_mask >> lane) & 1ull, k1 = (kept_mask >> (lane + 32)) & 1ull;
#pragma unroll
          for (int w = 1; w < kChunkWords; ++w) {
            if (w > t) {                                                 // (uniform)
              const unsigned long long v = (k0 ? S.mask[t * 64 + lane][w] : 0ull) | (k1 ? S.mask[t * 64 + 32 + lane][w] : 0ull);
              const unsigned lo = __reduce_or_sync(0xFFFFFFFFu, (unsigned)v), hi = __reduce_or_sync(0xFFFFFFFFu, (unsigned)(v >> 32));
              if (lane == w) rem |= ((unsigned long long)hi << 32) | lo;
            }
          }
        }
        // keep list: bit i of kept_mask -> position k + (kept bits below i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = lane + 32 * h;
          if ((kept_mask >> i) & 1ull) {
            const int pos = k + __popcll(kept_mask & ((1ull << i) - 1ull));
            keep[pos] = base + t * 64 + i;
            S.newk[pos - K] = t * 64 + i;
          }
        }
        k += __popcll(kept_mask);
      }
      if (lane == 0) S.state[0] = k;
      }
    if (cta == 0) {
      __syncthreads();
      const int k1 = S.state[0];
      for (int q = tid; q < k1 - K; q += kNmsThreads) kept_boxes[K + q] = S.box[S.newk[q]];
      __threadfence();
      __syncthreads();
      // publish (count, done) to every CTA of the cluster
      if (tid < kClusterCtas) {
        const int done = (k1 >= max_keep || base + kChunk >= nv) ? 1 : 0;
        const uint32_t a = map_to_cta(smem_u32(&S.state[0]), (uint32_t)tid);
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a), "r"(k1) : "memory");
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a + 4u), "r"(done) : "memory");
      }
    }
    cluster_sync_all();
    if (S.state[1]) break;
  }
  if (cta == 0 && tid == 0) num_keep[f] = S.state[0];
}

size_t nms_cluster_workspace_bytes(int F, int max_keep) {
  return F <= kNmsClusterMaxFrames ? align_up(sizeof(float4) * (size_t)F * (max_keep > 0 ? max_keep : 0), 256) : 0;
}

// 16-CTA clusters need the non-portable opt-in and a GPC with 16 free SMs: asked once per process.
static bool cluster_nms_available() {
  // function attributes are per device: asked once per (thread, device)
  static thread_local int cached_dev = -1;
  static thread_local bool cached = false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (dev == cached_dev) return cached;
  const bool ok = [] {
    if (cudaFuncSetAttribute(nms_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(nms_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(NmsClusterSmem)) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kClusterCtas, 1);
    cfg.blockDim = dim3(kNmsThreads);
    cfg.dynamicSmemBytes = sizeof(NmsClusterSmem);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kClusterCtas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, nms_cluster_kernel, &cfg) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return n > 0;
  }();
  cached_dev = dev;
  cached = ok;
  return ok;
}

// kept_scratch: nms_cluster_workspace_bytes(F, max_keep) bytes or NULL (then every frame count takes the
// single-CTA kernel).
int launch_nms_sorted(int F, int n, const float* boxes, const int32_t* n_valid, double thresh, int max_keep,
                      int32_t* keep, int32_t* num_keep, void* kept_scratch, cudaStream_t st) {
  if (max_keep <= 0 || n <= 0) {
    B2D_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int32_t) * F, st));
    return B2D_OK;
  }
  if (kept_scratch && F <= kNmsClusterMaxFrames && n > 2 * kTile && cluster_nms_available()) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kClusterCtas, F);
    cfg.blockDim = dim3(kNmsThreads);
    cfg.dynamicSmemBytes = sizeof(NmsClusterSmem);
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kClusterCtas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // the kernel waits for its predecessor itself
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    B2D_CUDA(cudaLaunchKernelEx(&cfg, nms_cluster_kernel, reinterpret_cast<const float4*>(boxes), n, n_valid,
                                float_floor_of(thresh), max_keep, keep, num_keep, static_cast<float4*>(kept_scratch)));
    B2D_LAUNCHED();
    return B2D_OK;
  }
  nms_sorted_kernel<<<F, kNmsThreads, 0, st>>>(reinterpret_cast<const float4*>(boxes), n, n_valid,
                                               float_floor_of(thresh), max_keep, keep, num_keep);
  B2D_LAUNCHED();
  return B2D_OK;
}

}  // namespace b2d

extern "C" size_t b2d_nms_workspace_bytes(int F, int max_keep) {
  if (F <= 0 || max_keep <= 0) return 0;
  return b2d::nms_cluster_workspace_bytes(F, max_keep);
}

extern "C" int b2d_nms_sorted(int F, int n, const float* boxes, const int32_t* n_valid, double thresh, int max_keep,
                              int32_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes, void* stream) {
  if (F <= 0 || n < 0 || !boxes || !keep || !num_keep) return B2D_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(boxes) & 15u) != 0) return B2D_ERR_INVALID_ARG;  // float4 loads
  // the workspace is optional: without it (or with more frames than the cluster path takes) the single-CTA kernel runs
  const size_t need = b2d::nms_cluster_workspace_bytes(F, max_keep);
  void* ws = (workspace && need && workspace_bytes >= need) ? workspace : nullptr;
  return b2d::launch_nms_sorted(F, n, boxes, n_valid, thresh, max_keep, keep, num_keep, ws, b2d::as_stream(stream));
}
