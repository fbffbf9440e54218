// RoIAlign forward, "sweep" kernel: the production path for feature maps whose rows fit a
// shared-memory ring (res101 C4 maps of KITTI / Waymo / BEV, FPN p4/p5).
//
// Why this shape.  RoIAlign with sampling_ratio 2 reads 784 taps per (roi, channel); at Waymo
// sizes that is 240 M four-byte gathers per frame, against 100 MB of HBM traffic.  The kernel is
// bound by the shared-memory pipe (128 B/clk/SM), not by HBM, so the two things that matter are
// (a) zero bank conflicts and (b) few instructions per tap.  Both follow from putting CHANNELS on
// the lanes: the ring is laid out [row][x][33] (32 channels + 1 pad word), a warp's 32 lanes read
// 32 consecutive words for every tap (always conflict-free, whatever the RoI geometry), and all
// tap indices/weights are warp-uniform, computed once per RoI by a prep kernel.
//
// Data flow per CTA = (32-channel group, frame[, item split]):
//   * the CTA sweeps the feature rows top to bottom through a ring of Rr rows; each feature byte
//     is read from HBM/L2 once per channel group and transposed NCHW -> [x][c] on the way in
//     (coalesced 128 B global reads along x, stride-33 conflict-free shared stores);
//   * rows for the next step are prefetched into registers while the current step computes;
//   * work items are (roi, bin-row ph); the prep kernel buckets them by the first feature row
//     they touch, so bucket k only needs rows [St*k, St*k + St + span_max - 1), all resident;
//   * a warp owns an item: lane = channel, loops pw and the s x s samples, 4 taps each;
//   * the 32 x PW results are staged through shared memory so global stores run along (c, pw).
// Items whose bin-row spans more rows than the ring holds (RoIs far taller than the frame) take
// a slow in-kernel path that reads global memory directly.
#include "roi_common.cuh"

namespace b2d {

constexpr int kSweepThreads = 512;
constexpr int kSweepWarps = kSweepThreads / 32;
constexpr int kCh = 32;        // channels per CTA (lanes)
constexpr int kPad = 33;       // words per pixel in the ring
constexpr int kMaxPF = 24;     // prefetch registers per thread
constexpr int kMaxPool = 16;   // PH, PW limit of this path
constexpr int kMaxGrid = 4;    // sampling_ratio limit of this path

struct SweepPlan {
  int Rr;        // ring rows
  int St;        // rows advanced per step
  int span_max;  // rows an item may span
  int nsteps;
  int XI;        // ceil(W / 32)
  size_t ring_bytes, stage_bytes;
  bool ok;
};

static SweepPlan plan_sweep(int H, int W, int PW) {
  SweepPlan p{};
  p.XI = ceil_div(W, 32);
  const size_t row_bytes = (size_t)W * kPad * sizeof(float);
  const int pwp = PW | 1;
  p.stage_bytes = (size_t)kSweepWarps * kCh * pwp * sizeof(float) + kPad * sizeof(float);
  const size_t budget = 227 * 1024 - 1024 - p.stage_bytes - (size_t)kSweepWarps * 2 * 256;   // 1 KB static shared, table slots
  int Rr = (int)(budget / row_bytes);
  const int xi_t = p.XI <= 4 ? p.XI : (p.XI <= 6 ? 6 : 12);
  if (p.XI > 12 || Rr < 5) { p.ok = false; return p; }
  if (Rr >= H) {
    p.Rr = H; p.St = H; p.span_max = H; p.nsteps = 1;
  } else {
    int st_cap = kMaxPF / (2 * xi_t);                        // prefetch registers: 2*St*XI <= kMaxPF
    int St = Rr / 4;
    if (St > st_cap) St = st_cap;
    if (St < 1) St = 1;
    p.Rr = Rr; p.St = St; p.span_max = Rr - 2 * St + 1; p.nsteps = ceil_div(H, St);
  }
  p.ring_bytes = (size_t)p.Rr * row_bytes;
  p.ok = p.span_max >= 3;
  return p;
}

struct SweepWs {
  float4* xtab;        // [n_list][PW*S]  {xlo*33, xhi*33 (int bits), w_lo, w_hi}
  float4* ytab;        // [n_list][PH*S]  {slot(ylo)*W*33, slot(yhi)*W*33 (int bits), w_lo, w_hi}
  float* count;        // [n_list] samples per bin
  int32_t* bucket_of;  // [n_list][PH]
  uint32_t* items;     // [F][items_stride]  (entry << 4) | ph, grouped by bucket
  int32_t* bucket_start;  // [F][nb + 1]
  float scale;            // spatial_scale / aligned, for the in-kernel slow path
  int aligned;
  size_t bytes;
};

static SweepWs carve_sweep(void* base, int F, int n_list, int per_frame, int H) {
  SweepWs w;
  size_t off = 0;
  char* p = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  w.xtab = reinterpret_cast<float4*>(take(sizeof(float4) * (size_t)n_list * kMaxPool * kMaxGrid));
  w.ytab = reinterpret_cast<float4*>(take(sizeof(float4) * (size_t)n_list * kMaxPool * kMaxGrid));
  w.count = reinterpret_cast<float*>(take(sizeof(float) * (size_t)n_list));
  w.bucket_of = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)n_list * kMaxPool));
  w.items = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * (size_t)F * per_frame * kMaxPool));
  w.bucket_start = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F * (H + 3)));
  w.bytes = off;
  return w;
}

// ------------------------------------------------------------------------------------------
// prep: per-RoI tap tables + items bucketed by first row.  One CTA per frame.
__global__ void __launch_bounds__(256)
roi_sweep_prep_kernel(RoiList L, int H, int W, int PH, int PW, float scale, int S, int aligned, int Rr, int St,
                      int span_max, int nsteps, int items_stride, SweepWs ws) {
  extern __shared__ int s_buckets[];   // [nb] counts, [nb] offsets, [nb] fill
  const int nb = nsteps + 1;
  int* cnt = s_buckets;
  int* offs = s_buckets + nb;
  int* fill = s_buckets + 2 * nb;
  const int f = blockIdx.x;
  for (int i = threadIdx.x; i < 3 * nb; i += blockDim.x) s_buckets[i] = 0;
  __syncthreads();
  int first = 0, n_ent = L.n;
  if (L.seg_count) {
    first = f * L.seg_stride;
    n_ent = L.seg_count[f];
  }
  for (int i = threadIdx.x; i < n_ent; i += blockDim.x) {
    const int e = first + i;
    const int r = L.ids ? L.ids[e] : e;
    const float* roi = L.rois + (size_t)r * 5;
    if (!L.seg_count && (int)roi[0] != f) continue;
    const float rr[5] = {roi[0], roi[1], roi[2], roi[3], roi[4]};
    RoiGeom g = roi_geometry(rr, scale, PH, PW, S, aligned != 0);
    ws.count[e] = g.count;
    float4* xt = ws.xtab + (size_t)e * PW * S;
    for (int k = 0; k < PW * S; ++k) {
      const AxisTap t = axis_tap(g.start_w, g.bin_w, k / S, k % S, S, W);
      xt[k] = make_float4(__int_as_float(t.lo * kPad), __int_as_float(t.hi * kPad), t.wlo, t.whi);
    }
    float4* yt = ws.ytab + (size_t)e * PH * S;
    for (int ph = 0; ph < PH; ++ph) {
      int y_first = H, y_last = -1;
      for (int iy = 0; iy < S; ++iy) {
        const AxisTap t = axis_tap(g.start_h, g.bin_h, ph, iy, S, H);
        if (t.ok) {
          y_first = min(y_first, t.lo);
          y_last = max(y_last, t.hi);
        }
      }
      int bucket = 0;
      if (y_last < 0) {
        y_first = 0;                                   // no valid sample row: zero weights, any resident row
      } else if (y_last - y_first + 1 > span_max) {
        bucket = nsteps;                               // slow path
      } else {
        bucket = y_first / St;
      }
      for (int iy = 0; iy < S; ++iy) {
        AxisTap t = axis_tap(g.start_h, g.bin_h, ph, iy, S, H);
        if (!t.ok) t.lo = t.hi = y_first;
        yt[ph * S + iy] = make_float4(__int_as_float((t.lo % Rr) * W * kPad), __int_as_float((t.hi % Rr) * W * kPad),
                                      t.wlo, t.whi);
      }
      ws.bucket_of[(size_t)e * PH + ph] = bucket;
      atomicAdd(&cnt[bucket], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    int32_t* bs = ws.bucket_start + (size_t)f * (nb + 1);
    for (int b = 0; b < nb; ++b) {
      offs[b] = run;
      bs[b] = run;
      run += cnt[b];
    }
    bs[nb] = run;
  }
  __syncthreads();
  uint32_t* items = ws.items + (size_t)f * items_stride;
  for (int i = threadIdx.x; i < n_ent; i += blockDim.x) {
    const int e = first + i;
    const int r = L.ids ? L.ids[e] : e;
    if (!L.seg_count && (int)L.rois[(size_t)r * 5] != f) continue;
    for (int ph = 0; ph < PH; ++ph) {
      const int b = ws.bucket_of[(size_t)e * PH + ph];
      const int slot = atomicAdd(&fill[b], 1);
      items[offs[b] + slot] = ((uint32_t)e << 4) | (uint32_t)ph;
    }
  }
}

// zero rows of padded list entries (seg mode), one block per (entry, frame)
__global__ void __launch_bounds__(256) roi_zero_pad_kernel(RoiList L, int per_roi, float* __restrict__ out) {
  const int f = blockIdx.y, ri = blockIdx.x;
  if (ri < L.seg_count[f]) return;
  const int e = f * L.seg_stride + ri;
  const int r = L.ids ? L.ids[e] : e;
  float* o = out + (size_t)r * per_roi;
  for (int i = threadIdx.x; i < per_roi; i += blockDim.x) o[i] = 0.0f;
}

// ------------------------------------------------------------------------------------------
template <int SMAX, int XI>
__global__ void __launch_bounds__(kSweepThreads, 1)
roi_align_fwd_sweep_kernel(const float* __restrict__ feat, RoiList L, int C, int H, int W, int PH, int PW_rt, float scale,
                           int S, int aligned, int Rr, int St, int span_max, int nsteps, int items_stride, SweepWs ws,
                           float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PWT = 0;                        // generic shape path; PW == 7 && S <= 2 takes sweep7 below
  const int PW = PW_rt;
  constexpr int kOutIters = kMaxPool;           // kCh * PW / 32 store rounds
  const int pwp = PW | 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f = blockIdx.y;
  const int c0 = blockIdx.x * kCh;
  const int nch = min(kCh, C - c0);
  const int split = gridDim.z, part = blockIdx.z;
  const int bins = PH * PW;
  const int nb = nsteps + 1;
  const int row_words = W * kPad;
  const int32_t* bstart = ws.bucket_start + (size_t)f * (nb + 1);
  const uint32_t* items = ws.items + (size_t)f * items_stride;
  const float* fbase = feat + ((size_t)f * C + c0) * H * W;
  // byte-addressed views: ring words then one zero pixel then the per-warp staging tiles
  char* ring_b = reinterpret_cast<char*>(smem);
  float* stage = smem + (size_t)Rr * row_words + kPad + (size_t)warp * kCh * pwp;
  if (tid < kPad) smem[(size_t)Rr * row_words + tid] = 0.0f;   // the pixel "right of" the last ring pixel

  // output scatter pattern of this lane: flat idx = lane + 32*j -> channel c_j = idx / PW (constant per lane)
  unsigned long long cpack = 0ull;   // 5 bits per round, up to 12 rounds; rounds beyond use the slow divide
#pragma unroll
  for (int j = 0; j < 12; ++j) cpack |= (unsigned long long)((lane + 32 * j) / PW) << (5 * j);

  // ---- row loader.  (channel, row) pair index pr = warp + 16*j: channel = pr & 31, row = pr >> 5.
  const int resident0 = min(H, St + span_max - 1);
  for (int pr = warp; pr < kCh * resident0; pr += kSweepWarps) {
    const int c = pr & 31, y = pr >> 5;
    if (c < nch) {
      const float* src = fbase + ((size_t)c * H + y) * W;
      float* dst = smem + (size_t)y * row_words + c;    // rows < Rr here, slot == row
      for (int x = lane; x < W; x += 32) dst[x * kPad] = __ldg(src + x);
    }
  }
  __syncthreads();

  constexpr int kPairs = kMaxPF / XI;   // (channel,row) pairs this warp prefetches per step
  float pf[kMaxPF];
  int p0 = resident0;                   // first row to prefetch
  int slot0 = resident0 % Rr;           // its ring slot
  for (int k = 0; k < nsteps; ++k) {
    const int prow = min(H, p0 + St) - p0;
    // ---- (a) prefetch rows [p0, p0 + prow) for step k+1 into registers
    if (prow > 0) {
#pragma unroll
      for (int j = 0; j < kPairs; ++j) {
        const int pr = warp + j * kSweepWarps;
        const int c = pr & 31, dy = pr >> 5;
        const bool on = dy < prow && c < nch;
        const float* src = fbase + ((size_t)c * H + (p0 + dy)) * W;
#pragma unroll
        for (int xi = 0; xi < XI; ++xi) {
          const int x = lane + xi * 32;
          pf[j * XI + xi] = (on && x < W) ? __ldg(src + x) : 0.0f;
        }
      }
    }
    // ---- (b) items of bucket k
    {
      const int i0 = bstart[k], i1 = bstart[k + 1];
      for (int it = i0 + part * kSweepWarps + warp; it < i1; it += kSweepWarps * split) {
        const uint32_t code = items[it];
        const int e = (int)(code >> 4), ph = (int)(code & 15u);
        const int r = L.ids ? L.ids[e] : e;
        const float4* yt = ws.ytab + (size_t)e * PH * S + ph * S;
        const float4* xt = ws.xtab + (size_t)e * PW * S;
        const float inv_cnt = 1.0f / ws.count[e];
        const char* rlo[SMAX];
        const char* rhi[SMAX];
        float hy[SMAX], ly[SMAX];
#pragma unroll
        for (int iy = 0; iy < SMAX; ++iy) {
          if (iy < S) {
            const float4 t = __ldg(yt + iy);
            rlo[iy] = ring_b + (size_t)(__float_as_int(t.x) + lane) * 4;
            rhi[iy] = ring_b + (size_t)(__float_as_int(t.y) + lane) * 4;
            hy[iy] = t.z;
            ly[iy] = t.w;
          } else {
            rlo[iy] = rhi[iy] = ring_b;
            hy[iy] = ly[iy] = 0.0f;
          }
        }
#pragma unroll
        for (int pw = 0; pw < (PWT > 0 ? PWT : 1); ++pw) {
          for (int pwr = pw; pwr < PW; pwr += (PWT > 0 ? PW : 1)) {   // runtime loop only when PWT == 0
            float acc = 0.0f;
#pragma unroll
            for (int ix = 0; ix < SMAX; ++ix) {
              if (ix < S) {
                const float4 t = __ldg(xt + pwr * S + ix);
                const int xl = __float_as_int(t.x) * 4;             // byte offset of pixel xl
                const int dx = (__float_as_int(t.y) - __float_as_int(t.x)) * 4;   // 0 at the right border, else 132
                const float hx = t.z, lx = t.w;
#pragma unroll
                for (int iy = 0; iy < SMAX; ++iy) {
                  if (iy < S) {
                    const char* pa = rlo[iy] + xl;
                    const char* pb = rhi[iy] + xl;
                    const float v00 = *reinterpret_cast<const float*>(pa);
                    const float v01 = *reinterpret_cast<const float*>(pa + kPad * 4);
                    const float v10 = *reinterpret_cast<const float*>(pb);
                    const float v11 = *reinterpret_cast<const float*>(pb + kPad * 4);
                    (void)dx;
                    const float top = fmaf(lx, v01, hx * v00);
                    const float bot = fmaf(lx, v11, hx * v10);
                    acc = fmaf(hy[iy], top, acc);
                    acc = fmaf(ly[iy], bot, acc);
                  }
                }
              }
            }
            stage[lane * pwp + pwr] = acc * inv_cnt;
          }
        }
        __syncwarp();
        float* o = out + ((size_t)r * C + c0) * bins + ph * PW;
        const int skip_o = bins - PW, skip_s = pwp - PW;
#pragma unroll
        for (int j = 0; j < kOutIters; ++j) {
          const int idx = lane + 32 * j;
          if (idx < kCh * PW) {
            const int c = j < 12 ? (int)((cpack >> (5 * j)) & 31ull) : idx / PW;
            if (c < nch) o[idx + c * skip_o] = stage[idx + c * skip_s];
          }
        }
        __syncwarp();
      }
    }
    // ---- (c) commit the prefetched rows into their ring slots (they held rows < St*k, dead now)
    if (prow > 0) {
#pragma unroll
      for (int j = 0; j < kPairs; ++j) {
        const int pr = warp + j * kSweepWarps;
        const int c = pr & 31, dy = pr >> 5;
        if (dy < prow && c < nch) {
          int slot = slot0 + dy;
          if (slot >= Rr) slot -= Rr;
          float* dst = smem + (size_t)slot * row_words + c;
#pragma unroll
          for (int xi = 0; xi < XI; ++xi) {
            const int x = lane + xi * 32;
            if (x < W) dst[x * kPad] = pf[j * XI + xi];
          }
        }
      }
    }
    p0 += prow;
    slot0 += prow;
    if (slot0 >= Rr) slot0 -= Rr;
    __syncthreads();
  }

  // ---- slow path: items spanning more rows than the ring holds; taps straight from global
  {
    const int i0 = bstart[nsteps], i1 = bstart[nsteps + 1];
    const bool ch_ok = lane < nch;
    for (int it = i0 + part * kSweepWarps + warp; it < i1; it += kSweepWarps * split) {
      const uint32_t code = items[it];
      const int e = (int)(code >> 4), ph = (int)(code & 15u);
      const int r = L.ids ? L.ids[e] : e;
      const float* roi = L.rois + (size_t)r * 5;
      const float rr[5] = {__ldg(roi), __ldg(roi + 1), __ldg(roi + 2), __ldg(roi + 3), __ldg(roi + 4)};
      const RoiGeom g = roi_geometry(rr, scale, PH, PW, S, aligned != 0);
      const float* plane = fbase + (size_t)(ch_ok ? lane : 0) * H * W;
      for (int pw = 0; pw < PW; ++pw) {
        float acc = 0.0f;
        for (int iy = 0; iy < S; ++iy) {
          const AxisTap ty = axis_tap(g.start_h, g.bin_h, ph, iy, S, H);
          for (int ix = 0; ix < S; ++ix) {
            const AxisTap tx = axis_tap(g.start_w, g.bin_w, pw, ix, S, W);
            if (!(ty.ok && tx.ok)) continue;
            const float top = tx.wlo * __ldg(plane + ty.lo * W + tx.lo) + tx.whi * __ldg(plane + ty.lo * W + tx.hi);
            const float bot = tx.wlo * __ldg(plane + ty.hi * W + tx.lo) + tx.whi * __ldg(plane + ty.hi * W + tx.hi);
            acc += ty.wlo * top + ty.whi * bot;
          }
        }
        if (ch_ok) out[((size_t)r * C + c0 + lane) * bins + ph * PW + pw] = acc / g.count;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Fast path: PW == 7, sampling_ratio 1 or 2 (cfg.POOLING_SIZE = 7, model/config.py:367).
//  * per-warp tap tables live in shared memory, double buffered and filled by 1-D bulk TMA
//    (cp.async.bulk + mbarrier) one item ahead, across step boundaries: no table load ever
//    sits on the critical path;
//  * separable evaluation with column reuse: for a pixel column x the row-combined value
//        G(x) = sum_iy  hy[iy]*F[ylo[iy]][x] + ly[iy]*F[yhi[iy]][x]
//    is shared by every sample column that touches x.  Sample columns are walked left to right
//    keeping G(cur) and G(cur+1); when the next sample's x_low is the same pixel or the next one
//    (always, for RoIs up to ~14 feature pixels wide) only one new G is evaluated.  Small RoIs drop
//    from 784 taps to ~100-300; large RoIs degrade gracefully to the direct 784.
constexpr int kTblBytes = 256;   // 14 x-entries + 2 y-entries of 16 B

template <int OFF>
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));   // not volatile: loads of an item may reorder
  return v;
}

template <int XI, int S>
__global__ void __launch_bounds__(kSweepThreads, 1)
roi_align_fwd_sweep7_kernel(const float* __restrict__ feat, RoiList L, int C, int H, int W, int PH, int /*S_rt*/, int Rr,
                            int St, int span_max, int nsteps, int items_stride, SweepWs ws,
                            float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t bars[kSweepWarps][2];
  constexpr int PW = 7;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f = blockIdx.y;
  const int c0 = blockIdx.x * kCh;
  const int nch = min(kCh, C - c0);
  const int split = gridDim.z, part = blockIdx.z;
  const int stride = kSweepWarps * split;
  const int bins = PH * PW;
  const int nb = nsteps + 1;
  const int row_words = W * kPad;
  const int32_t* bstart = ws.bucket_start + (size_t)f * (nb + 1);
  const uint32_t* items = ws.items + (size_t)f * items_stride;
  const float* fbase = feat + ((size_t)f * C + c0) * H * W;
  // dynamic shared: [table slots][ring][zero pixel][staging tiles]
  char* tbl = reinterpret_cast<char*>(smem) + (size_t)warp * 2 * kTblBytes;
  float* ring = smem + (size_t)kSweepWarps * 2 * kTblBytes / sizeof(float);
  const uint32_t ring_s = smem_u32(ring);
  float* zero_px = ring + (size_t)Rr * row_words;
  float* stage = zero_px + kPad + (size_t)warp * kCh * PW;
  if (tid < kPad) zero_px[tid] = 0.0f;
  if (lane == 0) {
    mbar_init(&bars[warp][0], 1);
    mbar_init(&bars[warp][1], 1);
  }
  unsigned cpack0 = 0u, cpack1 = 0u;   // channel of flat index lane + 32*j, 5 bits each (j = 0..6)
#pragma unroll
  for (int j = 0; j < 6; ++j) cpack0 |= (unsigned)((lane + 32 * j) / PW) << (5 * j);
  cpack1 = (unsigned)((lane + 32 * 6) / PW);

  const int resident0 = min(H, St + span_max - 1);
  for (int pr = warp; pr < kCh * resident0; pr += kSweepWarps) {
    const int c = pr & 31, y = pr >> 5;
    if (c < nch) {
      const float* src = fbase + ((size_t)c * H + y) * W;
      float* dst = ring + (size_t)y * row_words + c;
      for (int x = lane; x < W; x += 32) dst[x * kPad] = __ldg(src + x);
    }
  }

  // ---- this warp's item sequence (crosses buckets): position = (bucket k, index it)
  auto first_in = [&](int k) { return bstart[k] + part * kSweepWarps + warp; };
  auto normalize = [&](int& k, int& it) {
    while (k < nb && it >= bstart[k + 1]) {
      ++k;
      if (k < nb) it = first_in(k);
    }
  };
  const uint32_t xbytes = (uint32_t)(PW * S * 16), ybytes = (uint32_t)(S * 16);
  auto issue_tables = [&](int slot, uint32_t code) {   // one lane
    const int e = (int)(code >> 4), ph = (int)(code & 15u);
    char* dst = tbl + slot * kTblBytes;
    mbar_expect_tx(&bars[warp][slot], xbytes + ybytes);
    bulk_g2s(dst, ws.xtab + (size_t)e * PW * S, xbytes, &bars[warp][slot]);
    bulk_g2s(dst + 224, ws.ytab + ((size_t)e * PH + ph) * S, ybytes, &bars[warp][slot]);
  };

  int kc = 0, itc = first_in(0);
  normalize(kc, itc);
  int kn = kc, itn = itc + stride;
  normalize(kn, itn);
  uint32_t codec = kc < nb ? __ldg(items + itc) : 0u;
  uint32_t coden = kn < nb ? __ldg(items + itn) : 0u;
  __syncthreads();                     // ring prologue + mbarrier init visible
  unsigned parity = 0u;                // bit s = phase of slot s
  int slot = 0;
  if (lane == 0 && kc < nsteps) issue_tables(0, codec);

  constexpr int kPairs = kMaxPF / XI;
  float pf[kMaxPF];
  int p0 = resident0, slot0 = resident0 % Rr;
  const float inv_cnt = 1.0f / (float)(S * S);

  for (int k = 0; k <= nsteps; ++k) {
    const bool ring_step = k < nsteps;
    const int prow = ring_step ? min(H, p0 + St) - p0 : 0;
    if (prow > 0) {
#pragma unroll
      for (int j = 0; j < kPairs; ++j) {
        const int pr = warp + j * kSweepWarps;
        const int c = pr & 31, dy = pr >> 5;
        const bool on = dy < prow && c < nch;
        const float* src = fbase + ((size_t)c * H + (p0 + dy)) * W;
#pragma unroll
        for (int xi = 0; xi < XI; ++xi) {
          const int x = lane + xi * 32;
          pf[j * XI + xi] = (on && x < W) ? __ldg(src + x) : 0.0f;
        }
      }
    }
    while (kc == k) {
      // look two items ahead for the code, one ahead for the tables
      int k2 = kn, it2 = itn + stride;
      normalize(k2, it2);
      const uint32_t code2 = k2 < nb ? __ldg(items + it2) : 0u;
      if (lane == 0 && kn < nsteps) issue_tables(slot ^ 1, coden);
      const int e = (int)(codec >> 4), ph = (int)(codec & 15u);
      const int r = L.ids ? L.ids[e] : e;
      if (ring_step) {
        mbar_wait(&bars[warp][slot], (parity >> slot) & 1u);
        parity ^= 1u << slot;
        // shared-window byte addresses; every tap below is one LDS [reg + imm]
        const float4* tb = reinterpret_cast<const float4*>(tbl + slot * kTblBytes);
        uint32_t rb0, rb1, rb2, rb3;
        float rw0, rw1, rw2, rw3;
        {
          const float4 t0 = tb[14];
          rb0 = ring_s + (uint32_t)(__float_as_int(t0.x) + lane) * 4u;
          rb1 = ring_s + (uint32_t)(__float_as_int(t0.y) + lane) * 4u;
          rw0 = t0.z;
          rw1 = t0.w;
          rb2 = rb3 = rb0;
          rw2 = rw3 = 0.0f;
          if (S > 1) {
            const float4 t1 = tb[15];
            rb2 = ring_s + (uint32_t)(__float_as_int(t1.x) + lane) * 4u;
            rb3 = ring_s + (uint32_t)(__float_as_int(t1.y) + lane) * 4u;
            rw2 = t1.z;
            rw3 = t1.w;
          }
        }
        // Column walk, branch free.  The pixel held in g_lo after sample k is always xl_k, so whether
        // sample k can reuse / slide / must reload depends only on xl_k - xl_{k-1}: every predicate and
        // every tap address comes straight from the table, and all loads of an item are independent.
        uint32_t prev = 0x7fffffffu;
        float g_lo = 0.0f, g_hi = 0.0f;
#pragma unroll
        for (int pw = 0; pw < PW; ++pw) {
          float acc = 0.0f;
#pragma unroll
          for (int ix = 0; ix < 2; ++ix) {
            if (ix < S) {
              const float4 t = tb[pw * S + ix];
              const uint32_t xl = (uint32_t)__float_as_int(t.x) * 4u;
              const uint32_t d = xl - prev;
              prev = xl;
              const bool p_hi = d != 0u;
              const bool p_lo = p_hi && d != (uint32_t)(kPad * 4);
              const uint32_t a0 = rb0 + xl, a1 = rb1 + xl, a2 = rb2 + xl, a3 = rb3 + xl;
              float nlo = 0.0f, nhi = 0.0f;
              if (p_lo) {
                nlo = rw0 * lds_f32<0>(a0);
                nlo = fmaf(rw1, lds_f32<0>(a1), nlo);
                if (S > 1) {
                  nlo = fmaf(rw2, lds_f32<0>(a2), nlo);
                  nlo = fmaf(rw3, lds_f32<0>(a3), nlo);
                }
              }
              if (p_hi) {
                nhi = rw0 * lds_f32<kPad * 4>(a0);
                nhi = fmaf(rw1, lds_f32<kPad * 4>(a1), nhi);
                if (S > 1) {
                  nhi = fmaf(rw2, lds_f32<kPad * 4>(a2), nhi);
                  nhi = fmaf(rw3, lds_f32<kPad * 4>(a3), nhi);
                }
              }
              g_lo = p_lo ? nlo : (p_hi ? g_hi : g_lo);
              g_hi = p_hi ? nhi : g_hi;
              acc = fmaf(t.z, g_lo, acc);
              acc = fmaf(t.w, g_hi, acc);
            }
          }
          stage[lane * PW + pw] = acc * inv_cnt;
        }
        __syncwarp();
        float* o = out + ((size_t)r * C + c0) * bins + ph * PW;
        const int skip_o = bins - PW;
#pragma unroll
        for (int j = 0; j < PW; ++j) {
          const int idx = lane + 32 * j;
          const int c = j < 6 ? (int)((cpack0 >> (5 * j)) & 31u) : (int)cpack1;
          if (c < nch) o[idx + c * skip_o] = stage[idx];
        }
        __syncwarp();
        slot ^= 1;
      } else {
        // slow path (bucket nsteps): bin-row spans more rows than the ring; taps straight from global
        const bool ch_ok = lane < nch;
        const float* roi = L.rois + (size_t)r * 5;
        const float rr[5] = {__ldg(roi), __ldg(roi + 1), __ldg(roi + 2), __ldg(roi + 3), __ldg(roi + 4)};
        // scale / aligned are folded into the tables for the ring path; recover them from the prep's inputs
        const RoiGeom g = roi_geometry(rr, ws.scale, PH, PW, S, ws.aligned != 0);
        const float* plane = fbase + (size_t)(ch_ok ? lane : 0) * H * W;
        for (int pw = 0; pw < PW; ++pw) {
          float acc = 0.0f;
          for (int iy = 0; iy < S; ++iy) {
            const AxisTap ty = axis_tap(g.start_h, g.bin_h, ph, iy, S, H);
            for (int ix = 0; ix < S; ++ix) {
              const AxisTap tx = axis_tap(g.start_w, g.bin_w, pw, ix, S, W);
              if (!(ty.ok && tx.ok)) continue;
              const float top = tx.wlo * __ldg(plane + ty.lo * W + tx.lo) + tx.whi * __ldg(plane + ty.lo * W + tx.hi);
              const float bot = tx.wlo * __ldg(plane + ty.hi * W + tx.lo) + tx.whi * __ldg(plane + ty.hi * W + tx.hi);
              acc += ty.wlo * top + ty.whi * bot;
            }
          }
          if (ch_ok) out[((size_t)r * C + c0 + lane) * bins + ph * PW + pw] = acc / g.count;
        }
      }
      kc = kn; itc = itn; codec = coden;
      kn = k2; itn = it2; coden = code2;
    }
    if (prow > 0) {
#pragma unroll
      for (int j = 0; j < kPairs; ++j) {
        const int pr = warp + j * kSweepWarps;
        const int c = pr & 31, dy = pr >> 5;
        if (dy < prow && c < nch) {
          int sl = slot0 + dy;
          if (sl >= Rr) sl -= Rr;
          float* dst = ring + (size_t)sl * row_words + c;
#pragma unroll
          for (int xi = 0; xi < XI; ++xi) {
            const int x = lane + xi * 32;
            if (x < W) dst[x * kPad] = pf[j * XI + xi];
          }
        }
      }
    }
    p0 += prow;
    slot0 += prow;
    if (slot0 >= Rr) slot0 -= Rr;
    if (ring_step) __syncthreads();
  }
}

template <int SMAX, int XI>
static int launch_sweep(const SweepPlan& p, int F, int C, int H, int W, const float* feat, const RoiList& L, int PH,
                        int PW, float scale, int S, int aligned, int items_stride, const SweepWs& ws, int split,
                        float* out, cudaStream_t st) {
  dim3 grid(ceil_div(C, kCh), F, split);
  if (SMAX == 2 && PW == 7) {
    const size_t smem = p.ring_bytes + p.stage_bytes + (size_t)kSweepWarps * 2 * kTblBytes;
    if (S == 2) {
      B2D_CUDA(cudaFuncSetAttribute(roi_align_fwd_sweep7_kernel<XI, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      roi_align_fwd_sweep7_kernel<XI, 2><<<grid, kSweepThreads, smem, st>>>(feat, L, C, H, W, PH, S, p.Rr, p.St,
                                                                           p.span_max, p.nsteps, items_stride, ws, out);
    } else {
      B2D_CUDA(cudaFuncSetAttribute(roi_align_fwd_sweep7_kernel<XI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      roi_align_fwd_sweep7_kernel<XI, 1><<<grid, kSweepThreads, smem, st>>>(feat, L, C, H, W, PH, S, p.Rr, p.St,
                                                                           p.span_max, p.nsteps, items_stride, ws, out);
    }
    B2D_LAUNCHED();
    return B2D_OK;
  }
  const size_t smem = p.ring_bytes + p.stage_bytes;
  B2D_CUDA(cudaFuncSetAttribute(roi_align_fwd_sweep_kernel<SMAX, XI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  roi_align_fwd_sweep_kernel<SMAX, XI><<<grid, kSweepThreads, smem, st>>>(
      feat, L, C, H, W, PH, PW, scale, S, aligned, p.Rr, p.St, p.span_max, p.nsteps, items_stride, ws, out);
  B2D_LAUNCHED();
  return B2D_OK;
}

size_t sweep_workspace_bytes(int F, int H, int n_list) { return carve_sweep(nullptr, F, n_list, n_list, H).bytes; }

// Returns B2D_ERR_UNSUPPORTED when this path does not apply (caller falls back).
int roi_align_forward_sweep(int F, int C, int H, int W, const float* feat, const RoiList& L, int PH, int PW,
                            float scale, int S, int aligned, float* out, void* workspace, size_t workspace_bytes,
                            cudaStream_t st) {
  if (S < 1 || S > kMaxGrid || PH > kMaxPool || PW > kMaxPool || L.n >= (1 << 27)) return B2D_ERR_UNSUPPORTED;
  const SweepPlan p = plan_sweep(H, W, PW);
  if (!p.ok) return B2D_ERR_UNSUPPORTED;
  const int per_frame = L.seg_count ? L.seg_stride : L.n;
  SweepWs ws = carve_sweep(workspace, F, L.n, per_frame, H);
  if (!workspace || workspace_bytes < ws.bytes) return B2D_ERR_UNSUPPORTED;
  ws.scale = scale;
  ws.aligned = aligned;
  const int items_stride = per_frame * kMaxPool;
  const int nb = p.nsteps + 1;
  roi_sweep_prep_kernel<<<F, 256, sizeof(int) * 3 * nb, st>>>(L, H, W, PH, PW, scale, S, aligned, p.Rr, p.St, p.span_max,
                                                             p.nsteps, items_stride, ws);
  B2D_LAUNCHED();
  if (L.seg_count) {
    dim3 zg(L.seg_stride, F);
    roi_zero_pad_kernel<<<zg, 256, 0, st>>>(L, C * PH * PW, out);
    B2D_LAUNCHED();
  }
  // enough CTAs to fill the chip: split a frame's items across up to 4 CTAs per channel group
  const int groups = ceil_div(C, kCh) * F;
  int split = 1;
  while (split < 4 && groups * split < 2 * kNumSMs) split *= 2;
  const int xi = p.XI;
#define B2D_SWEEP(SM, X) return launch_sweep<SM, X>(p, F, C, H, W, feat, L, PH, PW, scale, S, aligned, items_stride, ws, split, out, st)
  if (S <= 2) {
    if (xi <= 1) B2D_SWEEP(2, 1);
    if (xi == 2) B2D_SWEEP(2, 2);
    if (xi == 3) B2D_SWEEP(2, 3);
    if (xi == 4) B2D_SWEEP(2, 4);
    if (xi <= 6) B2D_SWEEP(2, 6);
    B2D_SWEEP(2, 12);
  } else {
    if (xi <= 1) B2D_SWEEP(4, 1);
    if (xi == 2) B2D_SWEEP(4, 2);
    if (xi == 3) B2D_SWEEP(4, 3);
    if (xi == 4) B2D_SWEEP(4, 4);
    if (xi <= 6) B2D_SWEEP(4, 6);
    B2D_SWEEP(4, 12);
  }
#undef B2D_SWEEP
}

}  // namespace b2d
