// RoIAlign backward for 7x7 pooling with sampling_ratio 1 or 2: deterministic, atomic-free.
//
// Transpose of the forward "rows" kernel.  A CTA owns (32 channels) x (a band of feature rows) of
// grad_feat as shared-memory accumulators laid out [row][channel][pitch] (pitch = 1 mod 32 words:
// the 32 lanes = 32 channels of one pixel never conflict).  One warp owns ONE row of the band and is
// the only writer of that row; it walks the RoIs that touch the band in index order and, for every
// sample row of a RoI that lands on its row, scatters the 7 x S column samples of that bin-row:
//     acc[row][x_lo] += g * wy * hx,   acc[row][x_lo + 1] += g * wy * lx,   g = grad_out / count.
// Every accumulator has a single writer and a fixed order of additions, so results are bit-stable
// from run to run and each grad_feat element is written to HBM exactly once (the stock torchvision
// kernel issues one REDG.ADD.F32 per tap).  All geometry comes from per-RoI tables made by a prep
// kernel (14 sample rows + 14 sample columns), so the main kernel does no RoI arithmetic.
//
// grad_out is the only HBM stream with latency on the critical path: a warp would otherwise pay one
// DRAM round trip per sample row.  The [32 ch][49] slice of a RoI is 6272 contiguous bytes, so it is
// pulled into shared memory with ONE 1-D bulk TMA per RoI (its 464-byte table rides along), two batches
// of kTile RoIs in flight; lanes read the slice at stride 49 words (conflict-free).  Per (RoI, row) lane k
// tests sample row k against the warp's row, one ballot finds the samples that land on it, the warp
// contracts them, T[pw] = sum_k wy_k * g[ph_k][pw], and scatters T once.
#include "roi_common.cuh"

namespace b2d {

namespace bwd_rows {

constexpr int kCh = 32;
constexpr int kP = 7;
constexpr int kTabVec = 2 * kP * 2 + 1;   // float4 per RoI: 14 sample rows, 14 sample columns, 1 header
constexpr int kMaxWarps = 16;
#ifndef B2D_BWD_TILE
#define B2D_BWD_TILE 4
#endif
constexpr int kTile = B2D_BWD_TILE;                  // (smem sizing: 2 * kTile slots)
constexpr int kSlots = 2 * kTile;                    // RoI slices in flight (ring of single-RoI slots)
constexpr int kTileWords = kCh * kP * kP; // 1568 floats = 6272 bytes

struct Plan {
  int pitch, band, nbands, warps;
  size_t smem;
  bool ok;
};

static Plan make_plan(int H, int W) {
  Plan p{};
  p.pitch = ((W + 1 + 30) / 32) * 32 + 1;
  const size_t row_bytes = (size_t)kCh * p.pitch * 4;
  const size_t tiles = (size_t)2 * kTile * (kTileWords * 4 + kTabVec * 16);
  int band = (int)((227 * 1024 - 4096 - tiles) / row_bytes);
  if (band > kMaxWarps - 1) band = kMaxWarps - 1;
  if (band > H) band = H;
  if (band < 2) { p.ok = false; return p; }
  p.nbands = ceil_div(H, band);
  p.band = ceil_div(H, p.nbands);          // balance the bands
  p.warps = p.band + 1;                    // one warp per row + the warp that feeds the slice ring
  p.smem = (size_t)p.band * row_bytes + tiles + 128;
  p.ok = true;
  return p;
}

// tables: [0] {row_min, row_max (int bits), -, -}; [1..14] sample rows {ylo, yhi (int bits), wlo/count, whi/count}
// (invalid: ylo = yhi = -1); [15..28] sample columns {xlo*4 bytes (int bits), hx, lx, -} (invalid: weights 0, x 0)
template <int S>
__global__ void __launch_bounds__(128) prep_kernel(RoiList L, int H, int W, float scale, int aligned, float4* __restrict__ tab) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= L.n) return;
  const int r = L.ids ? L.ids[e] : e;
  const float* roi = L.rois + (size_t)r * 5;
  const float rr[5] = {roi[0], roi[1], roi[2], roi[3], roi[4]};
  const RoiGeom g = roi_geometry(rr, scale, kP, kP, S, aligned != 0);
  const float inv_cnt = 1.0f / g.count;
  float4* t = tab + (size_t)e * kTabVec;
  int rmin = H, rmax = -1;
  for (int k = 0; k < 2 * kP; ++k) {
    float4 v = make_float4(__int_as_float(-1), __int_as_float(-1), 0.f, 0.f);
    if (k < kP * S) {
      const AxisTap a = axis_tap(g.start_h, g.bin_h, S == 2 ? k / 2 : k, S == 2 ? k % 2 : 0, S, H);
      if (a.ok) {
        v = make_float4(__int_as_float(a.lo), __int_as_float(a.hi), a.wlo * inv_cnt, a.whi * inv_cnt);
        rmin = min(rmin, a.lo);
        rmax = max(rmax, a.hi);
      }
    }
    t[1 + k] = v;
  }
  for (int k = 0; k < 2 * kP; ++k) {
    float4 v = make_float4(__int_as_float(0), 0.f, 0.f, 0.f);
    if (k < kP * S) {
      const AxisTap a = axis_tap(g.start_w, g.bin_w, S == 2 ? k / 2 : k, S == 2 ? k % 2 : 0, S, W);
      if (a.ok) v = make_float4(__int_as_float(a.lo * 4), a.wlo, a.hi > a.lo ? a.whi : 0.0f, 0.f);
    }
    t[1 + 2 * kP + k] = v;
  }
  t[0] = make_float4(__int_as_float(rmin), __int_as_float(rmax), 0.f, 0.f);
}

template <int S>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
bwd_kernel(const float* __restrict__ grad_out, RoiList L, int C, int H, int W, int pitch, int band, int accumulate,
           int use_tma, const float4* __restrict__ tab, float* __restrict__ grad_feat) {
  extern __shared__ __align__(16) float acc[];
  __shared__ __align__(8) uint64_t s_full[kSlots];    // slice + table of the slot have landed (TMA bytes)
  __shared__ __align__(8) uint64_t s_empty[kSlots];   // every row warp is done with the slot
  __shared__ int s_list[kMaxWarps * 32];   // list entries of the current chunk that touch the band, in index order
  __shared__ int s_warp_cnt[kMaxWarps];
  const int f = blockIdx.z;
  const int c0 = blockIdx.y * kCh;
  const int y0 = blockIdx.x * band;
  const int rows = min(band, H - y0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarp = blockDim.x >> 5;
  const int nrow_warps = nwarp - 1;                     // warps 0 .. nrow_warps-1 own a row, the last warp feeds the ring
  const int nch = min(kCh, C - c0);
  const bool ch_ok = lane < nch;
  constexpr int bins = kP * kP;
  const int row_words = kCh * pitch;
  for (int i = tid; i < rows * row_words; i += blockDim.x) acc[i] = 0.0f;
  float* tiles = acc + (size_t)band * row_words;       // [kSlots][32 ch][49] grad_out slices ...
  float4* tabs = reinterpret_cast<float4*>(tiles + (size_t)kSlots * kTileWords);   // ... and [kSlots][29] RoI tables
  if (tid == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], nrow_warps);
    }
  }

  int first = 0, n_ent = L.n;
  if (L.seg_count) {
    first = f * L.seg_stride;
    n_ent = L.seg_count[f];
  }
  const int my_row = y0 + warp;                         // the row this warp owns (may be >= H: idle)
  const bool row_ok = warp < rows;
  float* my_acc = acc + (size_t)min(warp, band - 1) * row_words + (size_t)lane * pitch;
  const uint32_t my_acc_s = smem_u32(my_acc);
  // Slices travel through a ring of kSlots single-RoI slots handed over with mbarriers only: the feeding warp refills a
  // slot as soon as every row warp has left it, and a row warp that a RoI does not touch moves straight on to the next
  // one.  (Batches of four RoIs between CTA-wide barriers left the warps waiting for the busiest row of each batch:
  // 2.7 of 8.3 stall cycles per issued instruction.)  `g` counts the RoIs of all chunks: slot = g % kSlots,
  // use = g / kSlots gives the barrier parities on both sides.
  int g = 0;

  for (int base = 0; base < n_ent; base += (int)blockDim.x) {
    // ---- ordered list of the entries of this chunk (one per thread) that touch the band
    __syncthreads();
    const int i = base + tid;
    bool t = false;
    if (i < n_ent) {
      const int e = first + i;
      const int r = L.ids ? L.ids[e] : e;
      t = L.seg_count || (int)__ldg(L.rois + (size_t)r * 5) == f;
      if (t) {
        const float4 h = __ldg(tab + (size_t)e * kTabVec);
        t = __float_as_int(h.y) >= y0 && __float_as_int(h.x) < y0 + rows;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, t);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < nwarp; ++w) {
      if (w < warp) before += s_warp_cnt[w];
      total += s_warp_cnt[w];
    }
    if (t) s_list[before + __popc(bal & ((1u << lane) - 1u))] = first + i;
    __syncthreads();
    const int n_list = total;
    if (warp == nrow_warps) {
      // ---- feeder: one bulk TMA per slice (+ its table), as far ahead as the ring allows
      if (lane == 0) {
        for (int li = 0; li < n_list; ++li) {
          const int gi = g + li, slot = gi % kSlots, use = gi / kSlots;
          mbar_wait(&s_empty[slot], (uint32_t)((use & 1) ^ 1));      // (a fresh barrier passes the first round)
          const int e = s_list[li];
          const int r = L.ids ? L.ids[e] : e;
          uint64_t* bar = &s_full[slot];
          mbar_expect_tx(bar, (uint32_t)((use_tma ? kTileWords * 4 : 0) + kTabVec * 16));
          if (use_tma)
            bulk_g2s(tiles + (size_t)slot * kTileWords, grad_out + ((size_t)r * C + c0) * bins, (uint32_t)(kTileWords * 4), bar);
          bulk_g2s(tabs + (size_t)slot * kTabVec, tab + (size_t)e * kTabVec, (uint32_t)(kTabVec * 16), bar);
        }
      }
    } else {
      // ---- row warps: walk the list in order, each accumulating into its own row
      for (int li = 0; li < n_list; ++li) {
        const int gi = g + li, slot = gi % kSlots, use = gi / kSlots;
        mbar_wait(&s_full[slot], (uint32_t)(use & 1));
        if (row_ok) {
          const float4* t = tabs + (size_t)slot * kTabVec;
          // lane k tests sample row k of this RoI against my row
          float wy = 0.0f;
          if (lane < kP * S) {
            const float4 yr = t[1 + lane];
            if (__float_as_int(yr.x) == my_row) wy += yr.z;          // my row is the sample's lo row
            if (__float_as_int(yr.y) == my_row && __float_as_int(yr.y) != __float_as_int(yr.x)) wy += yr.w;
          }
          unsigned hits = __ballot_sync(0xffffffffu, wy != 0.0f);
          if (hits) {
            // samples that land on my row -> T[pw]
            float T[kP];
#pragma unroll
            for (int pw = 0; pw < kP; ++pw) T[pw] = 0.0f;
            const float* gt = tiles + (size_t)slot * kTileWords + lane * bins;
            const int e = s_list[li];
            const int r = L.ids ? L.ids[e] : e;
            const float* go = grad_out + ((size_t)r * C + c0 + (ch_ok ? lane : 0)) * bins;
            while (hits) {
              const int k = __ffs(hits) - 1;
              hits &= hits - 1;
              const float w = __shfl_sync(0xffffffffu, wy, k);
              const int ph = S == 2 ? k >> 1 : k;
#pragma unroll
              for (int pw = 0; pw < kP; ++pw) {
                const float gv = use_tma ? gt[ph * kP + pw] : (ch_ok ? __ldg(go + ph * kP + pw) : 0.0f);
                T[pw] = fmaf(gv, w, T[pw]);
              }
            }
#pragma unroll
            for (int k = 0; k < kP * S; ++k) {
              const float4 x4 = t[1 + 2 * kP + k];
              const float gv = T[S == 2 ? k >> 1 : k];
              const uint32_t a = my_acc_s + (uint32_t)__float_as_int(x4.x);
              // (shared-memory reductions instead of this read-modify-write - red.shared.add.f32, no load -> FMA -> store
              // chain - measured 1.71 ms against 1.07 for 8 x 256 RoIs: rejected)
              float v0, v1;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(a));
              asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(v1) : "r"(a));
              v0 = fmaf(gv, x4.y, v0);
              v1 = fmaf(gv, x4.z, v1);
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v0) : "memory");
              asm volatile("st.shared.f32 [%0+4], %1;" ::"r"(a), "f"(v1) : "memory");
            }
          }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_empty[slot])) : "memory");
      }
    }
    g += n_list;
  }
  __syncthreads();
  // ---- write-out: lanes sweep x, (channel, row) pairs over the warps
  for (int pr = warp; pr < nch * rows; pr += nwarp) {
    const int c = pr % nch, y = pr / nch;
    const float* src = acc + (size_t)y * row_words + (size_t)c * pitch;
    float* dst = grad_feat + (((size_t)f * C + c0 + c) * H + y0 + y) * W;
    for (int x = lane; x < W; x += 32) dst[x] = accumulate ? dst[x] + src[x] : src[x];
  }
}

}  // namespace bwd_rows

size_t bwd_rows_workspace_bytes(int n_list) { return align_up((size_t)n_list * bwd_rows::kTabVec * 16, 256); }

// Returns B2D_ERR_UNSUPPORTED when this path does not apply (caller falls back to the band kernel).
int roi_align_backward_rows(int F, int C, int H, int W, const float* grad_out, const RoiList& L, int PH, int PW,
                            float scale, int S, int aligned, int accumulate, float* grad_feat, void* workspace,
                            size_t workspace_bytes, cudaStream_t st) {
  using namespace bwd_rows;
  if (PH != kP || PW != kP || S < 1 || S > 2) return B2D_ERR_UNSUPPORTED;
  const Plan p = make_plan(H, W);
  if (!p.ok || !workspace || workspace_bytes < bwd_rows_workspace_bytes(L.n)) return B2D_ERR_UNSUPPORTED;
  float4* tab = static_cast<float4*>(workspace);
  // bulk TMA needs 16-byte aligned 6272-byte slices that stay inside the tensor
  const int use_tma = (C % kCh == 0) && (reinterpret_cast<uintptr_t>(grad_out) & 15u) == 0;
  dim3 grid(p.nbands, ceil_div(C, kCh), F);
#define B2D_BWD(SS)                                                                                            \
  do {                                                                                                         \
    if (L.n > 0) {                                                                                             \
      prep_kernel<SS><<<ceil_div(L.n, 128), 128, 0, st>>>(L, H, W, scale, aligned, tab);                       \
      B2D_LAUNCHED();                                                                                          \
    }                                                                                                          \
    B2D_CUDA(cudaFuncSetAttribute(bwd_kernel<SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));  \
    bwd_kernel<SS><<<grid, p.warps * 32, p.smem, st>>>(grad_out, L, C, H, W, p.pitch, p.band, accumulate,      \
                                                       use_tma, tab, grad_feat);                               \
    B2D_LAUNCHED();                                                                                            \
  } while (0)
  if (S == 2) B2D_BWD(2);
  else B2D_BWD(1);
#undef B2D_BWD
  return B2D_OK;
}

}  // namespace b2d
