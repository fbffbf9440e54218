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
  const size_t budget = 227 * 1024 - 1024 - p.stage_bytes - (size_t)kSweepWarps * 272;   // 1 KB static shared, record slots
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
  uint32_t* items;     // [F][items_stride]  (entry << 4) | ph, grouped by bucket (generic kernel)
  float4* records;     // [F][items_stride][17] self-contained item records (fast path)
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
  w.records = reinterpret_cast<float4*>(take(sizeof(float4) * (size_t)F * per_frame * 8 * 17));   // fast path: PH <= 8
  w.bucket_start = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F * (H + 4)));
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
//
// The prep kernel emits one self-contained 272-byte RECORD per work item (roi, bin-row), already
// in bucket order, so the main kernel does no per-item bookkeeping at all: a warp prefetches its
// next record into registers (17 lanes x 16 B, coalesced) while it computes the current one, drops
// it into a per-warp shared slot and runs class-specialised STRAIGHT-LINE code on it:
//   record[0]     = {roi row r, ph, class, x0 byte offset}
//   record[1]     = ring word offsets of the four pixel rows {ylo0, yhi0, ylo1, yhi1}
//   record[2]     = their weights, pre-divided by the sample count
//   record[3..16] = class 0 "dense8": Wx[7][8]; the bin-row touches <= 8 consecutive pixel columns
//                   x0..x0+7; evaluate G(x0+j) = sum_rows w*F[row][x0+j] once per column (32 taps,
//                   immediate offsets) and contract with the dense 7x8 weights: 152 instructions
//                   and 56 shared wavefronts instead of 784 taps;
//                   class 1 "direct": 14 sample columns {xl byte offset, -, hx, lx}: G at xl and
//                   xl+1 for every sample (the stock 16 taps per bin), no reuse;
//                   class 2: bin-row taller than the ring, evaluated from global memory (rare).
constexpr int kRecVec = 17;                 // float4 per record
constexpr int kRecBytes = kRecVec * 16;     // 272

template <int OFF>
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));   // not volatile: loads of an item may reorder
  return v;
}

// prep for the fast path: one thread per (entry, ph) item
template <int S>
__global__ void __launch_bounds__(512)
roi_sweep7_prep_kernel(RoiList L, int H, int W, int PH, float scale, int aligned, int Rr, int St, int span_max,
                       int nsteps, int items_stride, SweepWs ws) {
  extern __shared__ int s_buckets[];   // [nb] counts, [nb] offsets, [nb] fill
  constexpr int PW = 7;
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
  const int n_items = n_ent * PH;
  // pass A: bucket of every item
  for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
    const int e = first + i / PH, ph = i - (i / PH) * PH;
    const int r = L.ids ? L.ids[e] : e;
    const float* roi = L.rois + (size_t)r * 5;
    int bucket = -1;
    if (L.seg_count || (int)roi[0] == f) {
      const float rr[5] = {roi[0], roi[1], roi[2], roi[3], roi[4]};
      const RoiGeom g = roi_geometry(rr, scale, PH, PW, S, aligned != 0);
      int y_first = H, y_last = -1;
      for (int iy = 0; iy < S; ++iy) {
        const AxisTap t = axis_tap(g.start_h, g.bin_h, ph, iy, S, H);
        if (t.ok) {
          y_first = min(y_first, t.lo);
          y_last = max(y_last, t.hi);
        }
      }
      if (y_last < 0) bucket = 0;
      else if (y_last - y_first + 1 > span_max) bucket = nsteps;
      else bucket = y_first / St;
      atomicAdd(&cnt[bucket], 1);
    }
    ws.bucket_of[(size_t)e * PH + ph] = bucket;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    int32_t* bs = ws.bucket_start + (size_t)f * (nb + 2);
    for (int b = 0; b < nb; ++b) {
      offs[b] = run;
      bs[b] = run;
      run += cnt[b];
    }
    bs[nb] = run;
    bs[nb + 1] = run;
  }
  __syncthreads();
  // pass B: records
  float4* recs = ws.records + (size_t)f * items_stride * kRecVec;
  for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
    const int e = first + i / PH, ph = i - (i / PH) * PH;
    const int bucket = ws.bucket_of[(size_t)e * PH + ph];
    if (bucket < 0) continue;
    const int r = L.ids ? L.ids[e] : e;
    const float* roi = L.rois + (size_t)r * 5;
    const float rr[5] = {roi[0], roi[1], roi[2], roi[3], roi[4]};
    const RoiGeom g = roi_geometry(rr, scale, PH, PW, S, aligned != 0);
    float4* rec = recs + (size_t)(offs[bucket] + atomicAdd(&fill[bucket], 1)) * kRecVec;
    // rows
    int y_first = 0;
    {
      int yf = H;
      for (int iy = 0; iy < S; ++iy) {
        const AxisTap t = axis_tap(g.start_h, g.bin_h, ph, iy, S, H);
        if (t.ok) yf = min(yf, t.lo);
      }
      if (yf < H) y_first = yf;
    }
    int yo[4];
    float yw[4];
    const float inv_cnt = 1.0f / g.count;
    for (int iy = 0; iy < 2; ++iy) {
      AxisTap t;
      t.lo = t.hi = y_first;
      t.wlo = t.whi = 0.0f;
      t.ok = false;
      if (iy < S) {
        t = axis_tap(g.start_h, g.bin_h, ph, iy, S, H);
        if (!t.ok) t.lo = t.hi = y_first;
      }
      yo[2 * iy] = (t.lo % Rr) * W * kPad;
      yo[2 * iy + 1] = (t.hi % Rr) * W * kPad;
      yw[2 * iy] = t.wlo * inv_cnt;
      yw[2 * iy + 1] = t.whi * inv_cnt;
    }
    rec[1] = make_float4(__int_as_float(yo[0]), __int_as_float(yo[1]), __int_as_float(yo[2]), __int_as_float(yo[3]));
    rec[2] = make_float4(yw[0], yw[1], yw[2], yw[3]);
    // columns
    int x0 = W, x1 = -1;
    for (int k = 0; k < PW * S; ++k) {
      const AxisTap t = axis_tap(g.start_w, g.bin_w, k / S, k % S, S, W);
      if (t.ok) {
        x0 = min(x0, t.lo);
        x1 = max(x1, t.hi);
      }
    }
    int cls;
    if (bucket == nsteps) {
      cls = 2;
      x0 = 0;
    } else if (x1 < 0) {                      // no valid sample column: all-zero dense weights
      cls = 0;
      x0 = 0;
      for (int v = 3; v < kRecVec; ++v) rec[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (x1 - x0 + 1 <= 8) {
      cls = 0;
      for (int pw = 0; pw < PW; ++pw) {
        float wx[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int ix = 0; ix < S; ++ix) {
          const AxisTap t = axis_tap(g.start_w, g.bin_w, pw, ix, S, W);
          if (!t.ok) continue;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (t.lo - x0 == j) wx[j] += t.wlo;
            if (t.hi - x0 == j) wx[j] += t.whi;
          }
        }
        rec[3 + 2 * pw] = make_float4(wx[0], wx[1], wx[2], wx[3]);
        rec[4 + 2 * pw] = make_float4(wx[4], wx[5], wx[6], wx[7]);
      }
    } else {
      cls = 1;
      for (int k = 0; k < 14; ++k) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < PW * S) {
          const AxisTap t = axis_tap(g.start_w, g.bin_w, k / S, k % S, S, W);
          v = make_float4(__int_as_float(t.lo * kPad * 4), 0.f, t.wlo, t.whi);   // hi tap = lo + 1 pixel (weight 0 at the border)
        }
        rec[3 + k] = v;
      }
    }
    rec[0] = make_float4(__int_as_float(r), __int_as_float(ph), __int_as_float(cls), __int_as_float(x0 * kPad * 4));
  }
}

template <int XI, int S>
__global__ void __launch_bounds__(kSweepThreads, 1)
roi_align_fwd_sweep7_kernel(const float* __restrict__ feat, RoiList L, int C, int H, int W, int PH, int Rr, int St,
                            int span_max, int nsteps, int items_stride, SweepWs ws, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PW = 7;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f = blockIdx.y;
  const int c0 = blockIdx.x * kCh;
  const int nch = min(kCh, C - c0);
  const int split = gridDim.z, part = blockIdx.z;
  const int stride = kSweepWarps * split;
  const int my_off = part * kSweepWarps + warp;
  const int bins = PH * PW;
  const int nb = nsteps + 1;
  const int row_words = W * kPad;
  const int32_t* bstart = ws.bucket_start + (size_t)f * (nb + 2);
  const float4* recs = ws.records + (size_t)f * items_stride * kRecVec;
  const float* fbase = feat + ((size_t)f * C + c0) * H * W;
  // dynamic shared: [record slots][ring][zero pixel][staging tiles]
  float4* slot = reinterpret_cast<float4*>(smem) + (size_t)warp * kRecVec;
  float* ring = smem + (size_t)kSweepWarps * kRecBytes / sizeof(float);
  const uint32_t ring_s = smem_u32(ring);
  float* zero_px = ring + (size_t)Rr * row_words;
  float* stage = zero_px + kPad + (size_t)warp * kCh * PW;
  if (tid < kPad) zero_px[tid] = 0.0f;

  // output scatter pattern of this lane: flat index lane + 32*j of the [32][7] tile -> channel c_j
  int ooff[PW];
  unsigned omask = 0u;
#pragma unroll
  for (int j = 0; j < PW; ++j) {
    const int idx = lane + 32 * j;
    const int c = idx / PW;
    ooff[j] = idx + c * (bins - PW);
    if (c < nch) omask |= 1u << j;
  }

  const int resident0 = min(H, St + span_max - 1);
  for (int pr = warp; pr < kCh * resident0; pr += kSweepWarps) {
    const int c = pr & 31, y = pr >> 5;
    if (c < nch) {
      const float* src = fbase + ((size_t)c * H + y) * W;
      float* dst = ring + (size_t)y * row_words + c;
      for (int x = lane; x < W; x += 32) dst[x * kPad] = __ldg(src + x);
    }
  }
  __syncthreads();

  constexpr int kPairs = kMaxPF / XI;
  float pf[kMaxPF];
  int p0 = resident0, slot0 = resident0 % Rr;
  float4 rec_next = make_float4(0.f, 0.f, 0.f, 0.f);
  int have = -1;                          // item index whose record sits in rec_next

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
    const int i1 = bstart[k + 1], i2 = bstart[k + 2];
    for (int it = bstart[k] + my_off; it < i1; it += stride) {
      if (have != it && lane < kRecVec) rec_next = __ldg(recs + (size_t)it * kRecVec + lane);
      __syncwarp();
      if (lane < kRecVec) slot[lane] = rec_next;
      __syncwarp();
      int nxt = it + stride;
      if (nxt >= i1) {
        nxt = i1 + my_off;               // my first item of the next bucket, if any
        if (nxt >= i2) nxt = -1;
      }
      have = nxt;
      if (nxt >= 0 && lane < kRecVec) rec_next = __ldg(recs + (size_t)nxt * kRecVec + lane);

      const float4 hdr = slot[0];
      const int r = __float_as_int(hdr.x), ph = __float_as_int(hdr.y), cls = __float_as_int(hdr.z);
      if (cls != 2) {
        const float4 ro = slot[1], rw = slot[2];
        const uint32_t xb = (uint32_t)__float_as_int(hdr.w) + (uint32_t)lane * 4u + ring_s;
        const uint32_t a0 = xb + (uint32_t)__float_as_int(ro.x) * 4u, a1 = xb + (uint32_t)__float_as_int(ro.y) * 4u;
        const uint32_t a2 = xb + (uint32_t)__float_as_int(ro.z) * 4u, a3 = xb + (uint32_t)__float_as_int(ro.w) * 4u;
        if (cls == 0) {
          float G[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float g;
            switch (j) {   // immediate offsets j * 132
              case 0: g = rw.x * lds_f32<0>(a0); g = fmaf(rw.y, lds_f32<0>(a1), g); if (S > 1) { g = fmaf(rw.z, lds_f32<0>(a2), g); g = fmaf(rw.w, lds_f32<0>(a3), g); } break;
              case 1: g = rw.x * lds_f32<132>(a0); g = fmaf(rw.y, lds_f32<132>(a1), g); if (S > 1) { g = fmaf(rw.z, lds_f32<132>(a2), g); g = fmaf(rw.w, lds_f32<132>(a3), g); } break;
              case 2: g = rw.x * lds_f32<264>(a0); g = fmaf(rw.y, lds_f32<264>(a1), g); if (S > 1) { g = fmaf(rw.z, lds_f32<264>(a2), g); g = fmaf(rw.w, lds_f32<264>(a3), g); } break;
              case 3: g = rw.x * lds_f32<396>(a0); g = fmaf(rw.y, lds_f32<396>(a1), g); if (S > 1) { g = fmaf(rw.z, lds_f32<396>(a2), g); g = fmaf(rw.w, lds_f32<396>(a3), g); } break;
              case 4: g = rw.x * lds_f32<528>(a0); g = fmaf(rw.y, lds_f32<528>(a1), g); if (S > 1) { g = fmaf(rw.z, lds_f32<528>(a2), g); g = fmaf(rw.w, lds_f32<528>(a3), g); } break;
              case 5: g = rw.x * lds_f32<660>(a0); g = fmaf(rw.y, lds_f32<660>(a1), g); if (S > 1) { g = fmaf(rw.z, lds_f32<660>(a2), g); g = fmaf(rw.w, lds_f32<660>(a3), g); } break;
              case 6: g = rw.x * lds_f32<792>(a0); g = fmaf(rw.y, lds_f32<792>(a1), g); if (S > 1) { g = fmaf(rw.z, lds_f32<792>(a2), g); g = fmaf(rw.w, lds_f32<792>(a3), g); } break;
              default: g = rw.x * lds_f32<924>(a0); g = fmaf(rw.y, lds_f32<924>(a1), g); if (S > 1) { g = fmaf(rw.z, lds_f32<924>(a2), g); g = fmaf(rw.w, lds_f32<924>(a3), g); } break;
            }
            G[j] = g;
          }
#pragma unroll
          for (int pw = 0; pw < PW; ++pw) {
            const float4 wa = slot[3 + 2 * pw], wb = slot[4 + 2 * pw];
            float acc = wa.x * G[0];
            acc = fmaf(wa.y, G[1], acc);
            acc = fmaf(wa.z, G[2], acc);
            acc = fmaf(wa.w, G[3], acc);
            acc = fmaf(wb.x, G[4], acc);
            acc = fmaf(wb.y, G[5], acc);
            acc = fmaf(wb.z, G[6], acc);
            acc = fmaf(wb.w, G[7], acc);
            stage[lane * PW + pw] = acc;
          }
        } else {
#pragma unroll
          for (int pw = 0; pw < PW; ++pw) {
            float acc = 0.0f;
#pragma unroll
            for (int ix = 0; ix < S; ++ix) {
              const float4 t = slot[3 + pw * S + ix];
              // t.x is an absolute column byte offset; the record's x0 is already inside a0..a3
              const uint32_t xo = (uint32_t)__float_as_int(t.x) - (uint32_t)__float_as_int(hdr.w);
              float glo = rw.x * lds_f32<0>(a0 + xo), ghi = rw.x * lds_f32<132>(a0 + xo);
              glo = fmaf(rw.y, lds_f32<0>(a1 + xo), glo);
              ghi = fmaf(rw.y, lds_f32<132>(a1 + xo), ghi);
              if (S > 1) {
                glo = fmaf(rw.z, lds_f32<0>(a2 + xo), glo);
                ghi = fmaf(rw.z, lds_f32<132>(a2 + xo), ghi);
                glo = fmaf(rw.w, lds_f32<0>(a3 + xo), glo);
                ghi = fmaf(rw.w, lds_f32<132>(a3 + xo), ghi);
              }
              acc = fmaf(t.z, glo, acc);
              acc = fmaf(t.w, ghi, acc);
            }
            stage[lane * PW + pw] = acc;
          }
        }
        __syncwarp();
        float* o = out + ((size_t)r * C + c0) * bins + ph * PW;
#pragma unroll
        for (int j = 0; j < PW; ++j)
          if (omask & (1u << j)) o[ooff[j]] = stage[lane + 32 * j];
      } else {
        // class 2: bin-row spans more rows than the ring; taps straight from global memory
        const bool ch_ok = lane < nch;
        const float* roi = L.rois + (size_t)r * 5;
        const float rr[5] = {__ldg(roi), __ldg(roi + 1), __ldg(roi + 2), __ldg(roi + 3), __ldg(roi + 4)};
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
  const int nb = p.nsteps + 1;
  if (SMAX == 2 && PW == 7 && PH <= 8) {
    const size_t smem = p.ring_bytes + p.stage_bytes + (size_t)kSweepWarps * kRecBytes;
    const int rec_stride = items_stride / kMaxPool * 8;     // records region holds 8 bin-rows per entry
    if (S == 2) {
      roi_sweep7_prep_kernel<2><<<F, 512, sizeof(int) * 3 * nb, st>>>(L, H, W, PH, scale, aligned, p.Rr, p.St, p.span_max,
                                                                     p.nsteps, rec_stride, ws);
      B2D_LAUNCHED();
      B2D_CUDA(cudaFuncSetAttribute(roi_align_fwd_sweep7_kernel<XI, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      roi_align_fwd_sweep7_kernel<XI, 2><<<grid, kSweepThreads, smem, st>>>(feat, L, C, H, W, PH, p.Rr, p.St, p.span_max,
                                                                           p.nsteps, rec_stride, ws, out);
    } else {
      roi_sweep7_prep_kernel<1><<<F, 512, sizeof(int) * 3 * nb, st>>>(L, H, W, PH, scale, aligned, p.Rr, p.St, p.span_max,
                                                                     p.nsteps, rec_stride, ws);
      B2D_LAUNCHED();
      B2D_CUDA(cudaFuncSetAttribute(roi_align_fwd_sweep7_kernel<XI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
      roi_align_fwd_sweep7_kernel<XI, 1><<<grid, kSweepThreads, smem, st>>>(feat, L, C, H, W, PH, p.Rr, p.St, p.span_max,
                                                                           p.nsteps, rec_stride, ws, out);
    }
    B2D_LAUNCHED();
    return B2D_OK;
  }
  roi_sweep_prep_kernel<<<F, 256, sizeof(int) * 3 * nb, st>>>(L, H, W, PH, PW, scale, S, aligned, p.Rr, p.St, p.span_max,
                                                             p.nsteps, items_stride, ws);
  B2D_LAUNCHED();
  const size_t smem = p.ring_bytes + p.stage_bytes;
  B2D_CUDA(cudaFuncSetAttribute(roi_align_fwd_sweep_kernel<SMAX, XI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  roi_align_fwd_sweep_kernel<SMAX, XI><<<grid, kSweepThreads, smem, st>>>(
      feat, L, C, H, W, PH, PW, scale, S, aligned, p.Rr, p.St, p.span_max, p.nsteps, items_stride, ws, out);
  B2D_LAUNCHED();
  return B2D_OK;
}

size_t sweep_workspace_bytes(int F, int H, int n_list, int per_frame) {
  return carve_sweep(nullptr, F, n_list, per_frame, H).bytes;
}

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
