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
constexpr int kTile = B2D_BWD_TILE;                  // RoI tiles per batch (two batches in flight)
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
  if (band > kMaxWarps) band = kMaxWarps;
  if (band > H) band = H;
  if (band < 2) { p.ok = false; return p; }
  p.nbands = ceil_div(H, band);
  p.band = ceil_div(H, p.nbands);          // balance the bands
  p.warps = p.band;                        // one warp per row
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
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ int s_list[kMaxWarps * 32];   // list entries of the current chunk that touch the band, in index order
  __shared__ int s_warp_cnt[kMaxWarps];
  const int f = blockIdx.z;
  const int c0 = blockIdx.y * kCh;
  const int y0 = blockIdx.x * band;
  const int rows = min(band, H - y0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarp = blockDim.x >> 5;
  const int nch = min(kCh, C - c0);
  const bool ch_ok = lane < nch;
  constexpr int bins = kP * kP;
  const int row_words = kCh * pitch;
  for (int i = tid; i < rows * row_words; i += blockDim.x) acc[i] = 0.0f;
  float* tiles = acc + (size_t)band * row_words;       // [2][kTile][32 ch][49] grad_out slices ...
  float4* tabs = reinterpret_cast<float4*>(tiles + (size_t)2 * kTile * kTileWords);   // ... and [2][kTile][29] RoI tables
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
  }
  int phase0 = 0, phase1 = 0;

  int first = 0, n_ent = L.n;
  if (L.seg_count) {
    first = f * L.seg_stride;
    n_ent = L.seg_count[f];
  }
  const int my_row = y0 + warp;                         // the row this warp owns (may be >= H: idle)
  const bool row_ok = warp < rows;
  float* my_acc = acc + (size_t)warp * row_words + (size_t)lane * pitch;
  const uint32_t my_acc_s = smem_u32(my_acc);

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
    // ---- batches of kTile RoIs: grad_out slices by bulk TMA (two batches in flight), then every warp
    // walks the batch in list order and accumulates into its own row
    const int n_batch = (n_list + kTile - 1) / kTile;
    auto issue = [&](int b) {            // one thread: slices + tables of batch b -> stage b & 1
      const int l0 = b * kTile, l1 = min(n_list, l0 + kTile);
      uint64_t* bar = &s_bar[b & 1];
      mbar_expect_tx(bar, (uint32_t)(l1 - l0) * (uint32_t)((use_tma ? kTileWords * 4 : 0) + kTabVec * 16));
      for (int li = l0; li < l1; ++li) {
        const int e = s_list[li];
        const int r = L.ids ? L.ids[e] : e;
        const int slot = (b & 1) * kTile + (li - l0);
        if (use_tma)
          bulk_g2s(tiles + (size_t)slot * kTileWords, grad_out + ((size_t)r * C + c0) * bins, (uint32_t)(kTileWords * 4), bar);
        bulk_g2s(tabs + (size_t)slot * kTabVec, tab + (size_t)e * kTabVec, (uint32_t)(kTabVec * 16), bar);
      }
    };
    if (tid == 0) {
      if (n_batch > 0) issue(0);
      if (n_batch > 1) issue(1);
    }
    for (int b = 0; b < n_batch; ++b) {
      mbar_wait(&s_bar[b & 1], (uint32_t)((((b & 1) ? phase1 : phase0) + (b >> 1)) & 1));
      if (row_ok) {
        const int l1 = min(n_list, (b + 1) * kTile);
        for (int li = b * kTile; li < l1; ++li) {
          const int slot = (b & 1) * kTile + (li - b * kTile);
          const float4* t = tabs + (size_t)slot * kTabVec;
          // lane k tests sample row k of this RoI against my row
          float wy = 0.0f;
          if (lane < kP * S) {
            const float4 yr = t[1 + lane];
            if (__float_as_int(yr.x) == my_row) wy += yr.z;          // my row is the sample's lo row
            if (__float_as_int(yr.y) == my_row && __float_as_int(yr.y) != __float_as_int(yr.x)) wy += yr.w;
          }
          unsigned hits = __ballot_sync(0xffffffffu, wy != 0.0f);
          if (!hits) continue;
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
              const float g = use_tma ? gt[ph * kP + pw] : (ch_ok ? __ldg(go + ph * kP + pw) : 0.0f);
              T[pw] = fmaf(g, w, T[pw]);
            }
          }
#pragma unroll
          for (int k = 0; k < kP * S; ++k) {
            const float4 x4 = t[1 + 2 * kP + k];
            const float g = T[S == 2 ? k >> 1 : k];
            const uint32_t a = my_acc_s + (uint32_t)__float_as_int(x4.x);
            float v0, v1;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(a));
            asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(v1) : "r"(a));
            v0 = fmaf(g, x4.y, v0);
            v1 = fmaf(g, x4.z, v1);
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v0) : "memory");
            asm volatile("st.shared.f32 [%0+4], %1;" ::"r"(a), "f"(v1) : "memory");
          }
        }
      }
      __syncthreads();                                     // everyone is done with stage b & 1
      if (tid == 0 && b + 2 < n_batch) issue(b + 2);
    }
    phase0 += (n_batch + 1) >> 1;                          // completed phases of s_bar[0] ...
    phase1 += n_batch >> 1;                                // ... and of s_bar[1]
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
