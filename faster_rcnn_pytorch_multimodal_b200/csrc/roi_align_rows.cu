// RoIAlign forward, "rows" kernel: production path for 7x7 pooling with sampling_ratio 1 or 2
// (cfg.POOLING_SIZE = 7, model/config.py:367) on feature maps whose rows fit a shared-memory ring
// (res101 C4 maps of KITTI / Waymo / BEV, FPN p3..p5).
//
// What bounds RoIAlign on B200.  The op moves 100 MB per Waymo frame (39 MB of features in, 60 MB of
// pooled features out) but evaluates 240 M bilinear taps, and with channels on the lanes every tap
// is one 128-byte shared-memory wavefront.  The stock formulation (16 taps per output) needs 784
// wavefronts per (roi, 32 channels); at 1 wavefront/clk/SM that alone is 26 us per frame against
// 15 us of HBM time.  So the kernel is organised to cut WAVEFRONTS, not bytes:
//
//   * a work item is (roi, a range of bin-rows [ph0, ph0+nph)) whose sample rows fit the resident
//     window; small RoIs are ONE item (nph = 7);
//   * the item walks its distinct feature rows once.  For each row it evaluates the 7 column
//     interpolants T[pw] = sum_ix hx*F[row][xlo] + lx*F[row][xlo+1] (28 taps, the x weights and
//     addresses stay in registers for the whole item) and scatters them into the nph x 7
//     accumulators with that row's weights: acc[p][pw] += wy[row][p] * T[pw].
//     A 5x5-pixel RoI costs 5 rows x 28 taps = 140 wavefronts instead of 784; only RoIs whose bins
//     are taller than two pixels still pay 16 taps per output (their samples share nothing).
//   * tap addresses are [column register + uniform row register + imm]: the row offset is made
//     warp-uniform with one CREDUX per row, so the inner loop has no address arithmetic at all.
//
// Data flow per CTA = (32-channel group, frame[, item split]): the CTA sweeps the feature rows top
// to bottom through a ring [slot][channel][pitch] (pitch = 1 mod 32 words, so 32 lanes reading the
// same pixel of 32 channels never conflict); rows of the next step are prefetched into registers
// while the current step computes, so every feature byte is read from HBM/L2 once per channel
// group.  A prep kernel turns the RoIs of a frame into self-contained item records bucketed by
// first row; warps claim items from a shared counter (dynamic balancing) and prefetch the next
// record while they compute.
#include "roi_common.cuh"

namespace b2d {

namespace rows {

constexpr int kWarps = 12;
constexpr int kThreads = kWarps * 32;
constexpr int kCh = 32;            // channels per CTA (lanes)
constexpr int kP = 7;              // PH = PW = 7
constexpr int kRecVec = 32;        // float4 per item record
constexpr int kRecBytes = kRecVec * 16;
constexpr int kMaxRows = 12;       // distinct feature rows per item
constexpr int kXVec = 7;           // float4 of column taps {xo_a, lx_a, xo_b, lx_b}
constexpr int kRowVec0 = 1 + kXVec;
constexpr int kMaxBlk = 64;        // ring blocks (mbarrier pairs)

// Ring geometry.  The ring holds nblk blocks of St rows.  Bucket k (items whose first row lies in
// block k) may touch blocks k .. k + nbk - 1; the other nblk - nbk blocks are slack: they are being
// refilled while slower warps still work on older buckets, so warps drift apart by up to that many
// blocks instead of meeting at a barrier every step.
struct Plan {
  int pitch;      // words per (slot, channel) row, = 1 mod 32 and > W
  int row_bytes;  // bytes per ring slot
  int Rr, St, nblk, nbk, span_max, nsteps;
  size_t smem;
  bool ok;
};

static Plan make_plan(int H, int W) {
  Plan p{};
  p.pitch = ((W + 1 + 30) / 32) * 32 + 1;            // smallest value = 1 (mod 32) that is >= W + 1
  p.row_bytes = kCh * p.pitch * 4;
  const size_t fixed = (size_t)kWarps * kRecBytes + (size_t)kWarps * kCh * kP * 4 + 64;
  const size_t budget = 227 * 1024 - 2048 - fixed;
  int Rr = (int)(budget / p.row_bytes);
  if (Rr < 6) { p.ok = false; return p; }
  if (Rr >= H) {
    p.Rr = H; p.St = H; p.nblk = 1; p.nbk = 1; p.span_max = H; p.nsteps = 1;
  } else {
    int St = Rr / 12;
    if (St < 1) St = 1;
    Rr -= Rr % St;                                   // blocks of St rows never wrap inside the ring
    const int nblk = Rr / St;
    const int slack = nblk >= 12 ? 3 : (nblk >= 8 ? 2 : 1);
    p.Rr = Rr; p.St = St; p.nblk = nblk; p.nbk = nblk - slack;
    p.span_max = (p.nbk - 1) * St + 1;
    p.nsteps = ceil_div(H, St);
  }
  if (p.span_max > kMaxRows) p.span_max = kMaxRows;
  p.smem = fixed + (size_t)p.Rr * p.row_bytes;
  p.ok = p.span_max >= 4 && p.nblk <= kMaxBlk;
  return p;
}

struct Ws {
  float4* records;        // [F][items_cap][kRecVec]
  int32_t* bucket_start;  // [F][nb + 2]
  float scale;
  int aligned;
  size_t bytes;
};

static Ws carve(void* base, int F, int per_frame, int H) {
  Ws w{};
  size_t off = 0;
  char* p = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  w.records = reinterpret_cast<float4*>(take((size_t)F * per_frame * kP * kRecBytes));
  w.bucket_start = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F * (H + 4)));
  w.bytes = off;
  return w;
}

// ------------------------------------------------------------------------------------------
// Items of one RoI.  Calls emit(ph0, nph, first_row, last_row, slow) for each item, in ph order.
template <int S, class Emit>
__device__ __forceinline__ void for_each_item(const RoiGeom& g, int H, int span_max, Emit emit) {
  int a = 0, cf = H, cl = -1;
  for (int ph = 0; ph < kP; ++ph) {
    int rf = H, rl = -1;
    for (int iy = 0; iy < S; ++iy) {
      const AxisTap t = axis_tap(g.start_h, g.bin_h, ph, iy, S, H);
      if (t.ok) {
        rf = min(rf, t.lo);
        rl = max(rl, t.hi);
      }
    }
    const int nf = min(cf, rf), nl = max(cl, rl);
    if (ph > a && nl >= 0 && nl - nf + 1 > span_max) {
      emit(a, ph - a, cf, cl, false);
      a = ph;
      cf = rf;
      cl = rl;
    } else {
      cf = nf;
      cl = nl;
    }
    if (cl >= 0 && cl - cf + 1 > span_max) {      // a single bin-row taller than the window: slow path
      emit(a, 1, cf, cl, true);
      a = ph + 1;
      cf = H;
      cl = -1;
    }
  }
  if (a < kP) emit(a, kP - a, cf, cl, false);
}

template <int S>
__global__ void __launch_bounds__(256)
prep_kernel(RoiList L, int H, int W, float scale, int aligned, int Rr, int St, int span_max, int nsteps,
            int row_bytes, int items_cap, Ws ws) {
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
  // pass A: bucket histogram
  for (int i = threadIdx.x; i < n_ent; i += blockDim.x) {
    const int e = first + i;
    const int r = L.ids ? L.ids[e] : e;
    const float* roi = L.rois + (size_t)r * 5;
    if (!L.seg_count && (int)roi[0] != f) continue;
    const float rr[5] = {roi[0], roi[1], roi[2], roi[3], roi[4]};
    const RoiGeom g = roi_geometry(rr, scale, kP, kP, S, aligned != 0);
    for_each_item<S>(g, H, span_max, [&](int, int, int cf, int cl, bool slow) {
      const int b = slow ? nsteps : (cl < 0 ? 0 : cf / St);
      atomicAdd(&cnt[b], 1);
    });
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
  float4* recs = ws.records + (size_t)f * items_cap * kRecVec;
  for (int i = threadIdx.x; i < n_ent; i += blockDim.x) {
    const int e = first + i;
    const int r = L.ids ? L.ids[e] : e;
    const float* roi = L.rois + (size_t)r * 5;
    if (!L.seg_count && (int)roi[0] != f) continue;
    const float rr[5] = {roi[0], roi[1], roi[2], roi[3], roi[4]};
    const RoiGeom g = roi_geometry(rr, scale, kP, kP, S, aligned != 0);
    const float inv_cnt = 1.0f / g.count;
    for_each_item<S>(g, H, span_max, [&](int ph0, int nph, int cf, int cl, bool slow) {
      const int b = slow ? nsteps : (cl < 0 ? 0 : cf / St);
      float4* rec = recs + (size_t)(offs[b] + atomicAdd(&fill[b], 1)) * kRecVec;
      // column taps: {byte offset of the lo column, weight of the hi column}; the hi column is lo + 1.
      // Invalid samples and the clamped right border point at / run into the zero pad column W.
      for (int v = 0; v < kXVec; ++v) {
        float q[4] = {0.f, 0.f, 0.f, 0.f};
        for (int h = 0; h < 2; ++h) {
          const int k = 2 * v + h;                       // column slot: pw = k / 2, ix = k % 2 (S == 2)
          int xo = W * 4;
          float lx = 0.0f;
          if (k < kP * S) {
            const AxisTap t = axis_tap(g.start_w, g.bin_w, S == 2 ? k / 2 : k, S == 2 ? k % 2 : 0, S, W);
            if (t.ok) {
              xo = t.lo * 4;
              lx = t.whi;
            }
          }
          q[2 * h] = __int_as_float(xo);
          q[2 * h + 1] = lx;
        }
        rec[1 + v] = make_float4(q[0], q[1], q[2], q[3]);
      }
      // distinct feature rows of the item and their weights per bin-row
      int nrows = 0;
      if (!slow) {
        int row_id[kMaxRows];
        float wy[kMaxRows][kP];
        for (int p = 0; p < nph; ++p) {
          for (int iy = 0; iy < S; ++iy) {
            const AxisTap t = axis_tap(g.start_h, g.bin_h, ph0 + p, iy, S, H);
            if (!t.ok) continue;
            for (int h = 0; h < 2; ++h) {
              const int y = h ? t.hi : t.lo;
              const float w = (h ? t.whi : t.wlo) * inv_cnt;
              if (h && t.hi == t.lo) continue;           // clamped bottom border: hi weight is 0
              int j = nrows - 1;
              while (j >= 0 && row_id[j] != y) --j;      // samples are monotone: found near the end
              if (j < 0) {
                j = nrows++;
                row_id[j] = y;
                for (int q = 0; q < kP; ++q) wy[j][q] = 0.0f;
              }
              wy[j][p] += w;
            }
          }
        }
        const int rv = nph > 3 ? 2 : 1;
        for (int j = 0; j < nrows; ++j) {
          const int off = (row_id[j] % Rr) * row_bytes;
          rec[kRowVec0 + j * rv] = make_float4(__int_as_float(off), wy[j][0], wy[j][1], wy[j][2]);
          if (rv == 2) rec[kRowVec0 + j * rv + 1] = make_float4(wy[j][3], wy[j][4], wy[j][5], wy[j][6]);
        }
      }
      const int code = ph0 | (nph << 4) | (nrows << 8) | ((slow ? 1 : 0) << 16);
      rec[0] = make_float4(__int_as_float(r), __int_as_float(code), __int_as_float(b), 0.f);
    });
  }
}

// zero rows of padded list entries (seg mode), one block per (entry, frame)
__global__ void __launch_bounds__(256) zero_pad_kernel(RoiList L, int per_roi, float* __restrict__ out) {
  const int f = blockIdx.y, ri = blockIdx.x;
  if (ri < L.seg_count[f]) return;
  const int e = f * L.seg_stride + ri;
  const int r = L.ids ? L.ids[e] : e;
  float* o = out + (size_t)r * per_roi;
  for (int i = threadIdx.x; i < per_roi; i += blockDim.x) o[i] = 0.0f;
}

// ------------------------------------------------------------------------------------------
// 4-byte asynchronous copy global -> shared (LDGSTS): the ring layout interleaves channels at word
// granularity, so neither 16-byte cp.async nor bulk TMA can write it; 4-byte copies keep the fill
// off the register file and off the issue slots of the compute code.
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
// arrive on `bar` once every cp.async this thread has issued so far has landed (the arrival is
// counted against the barrier's expected count: .noinc)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float lds_at(uint32_t addr) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));   // not volatile: taps of an item may reorder
  return v;
}

// One item: nph = NPH bin-rows of one RoI for this lane's channel.
template <int NPH, int S>
__device__ __forceinline__ void run_item(const float4* __restrict__ slot, int nrows, uint32_t lane_base,
                                         float* __restrict__ stage, int lane, float* __restrict__ o,
                                         const int (&ooff)[kP], unsigned omask) {
  constexpr int NX = kP * S;
  constexpr int RV = NPH > 3 ? 2 : 1;
  uint32_t xa[NX];
  float lx[NX], hx[NX];
#pragma unroll
  for (int v = 0; v < kXVec; ++v) {
    const float4 q = slot[1 + v];
    if (2 * v < NX) {
      xa[2 * v] = lane_base + (uint32_t)__float_as_int(q.x);
      lx[2 * v] = q.y;
      hx[2 * v] = 1.0f - q.y;
    }
    if (2 * v + 1 < NX) {
      xa[2 * v + 1] = lane_base + (uint32_t)__float_as_int(q.z);
      lx[2 * v + 1] = q.w;
      hx[2 * v + 1] = 1.0f - q.w;
    }
  }
  float acc[NPH][kP];
#pragma unroll
  for (int p = 0; p < NPH; ++p)
#pragma unroll
    for (int pw = 0; pw < kP; ++pw) acc[p][pw] = 0.0f;

  for (int i = 0; i < nrows; ++i) {
    const float4 e0 = slot[kRowVec0 + i * RV];
    float wy[kP];
    wy[0] = e0.y;
    wy[1] = e0.z;
    wy[2] = e0.w;
    if (RV == 2) {
      const float4 e1 = slot[kRowVec0 + i * RV + 1];
      wy[3] = e1.x;
      wy[4] = e1.y;
      wy[5] = e1.z;
      wy[6] = e1.w;
    }
    // the row offset is the same in every lane; the reduction tells ptxas so (CREDUX -> uniform
    // register), which lets every tap below use [column + uniform row + imm] addressing
    const uint32_t ro = __reduce_max_sync(0xffffffffu, (uint32_t)__float_as_int(e0.x));
#pragma unroll
    for (int pw = 0; pw < kP; ++pw) {
      float t;
      if (S == 2) {
        const float a = lds_at(xa[2 * pw] + ro), b = lds_at(xa[2 * pw] + ro + 4);
        const float c = lds_at(xa[2 * pw + 1] + ro), d = lds_at(xa[2 * pw + 1] + ro + 4);
        t = hx[2 * pw] * a;
        t = fmaf(lx[2 * pw], b, t);
        t = fmaf(hx[2 * pw + 1], c, t);
        t = fmaf(lx[2 * pw + 1], d, t);
      } else {
        const float a = lds_at(xa[pw] + ro), b = lds_at(xa[pw] + ro + 4);
        t = hx[pw] * a;
        t = fmaf(lx[pw], b, t);
      }
#pragma unroll
      for (int p = 0; p < NPH; ++p) acc[p][pw] = fmaf(wy[p], t, acc[p][pw]);
    }
  }
  // results: stage one bin-row [32 ch][7] at a time so that global stores run along (c, pw)
#pragma unroll
  for (int p = 0; p < NPH; ++p) {
#pragma unroll
    for (int pw = 0; pw < kP; ++pw) stage[lane * kP + pw] = acc[p][pw];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kP; ++j)
      if (omask & (1u << j)) o[p * kP + ooff[j]] = stage[lane + 32 * j];
    __syncwarp();
  }
}

template <int S>
__global__ void __launch_bounds__(kThreads, 1)
fwd_kernel(const float* __restrict__ feat, RoiList L, int C, int H, int W, int pitch, int St, int nblk, int nbk,
           int nsteps, int items_cap, Ws ws, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxBlk];   // block b landed (every thread arrives through cp.async)
  __shared__ __align__(8) uint64_t done_bar[kMaxBlk];   // every warp is past bucket j
  __shared__ int s_ctr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f = blockIdx.y;
  const int c0 = blockIdx.x * kCh;
  const int nch = min(kCh, C - c0);
  const int split = gridDim.z, part = blockIdx.z;
  constexpr int bins = kP * kP;
  const int nb = nsteps + 1;
  const int row_words = kCh * pitch;
  const int Rr = St * nblk;
  const float* fbase = feat + ((size_t)f * C + c0) * H * W;
  // dynamic shared: [record slots][staging tiles][ring]
  float4* slot = reinterpret_cast<float4*>(smem) + (size_t)warp * kRecVec;
  float* stage = smem + (size_t)kWarps * kRecBytes / 4 + (size_t)warp * kCh * kP;
  float* ring = smem + (size_t)kWarps * kRecBytes / 4 + (size_t)kWarps * kCh * kP + 16;
  const uint32_t ring_s = smem_u32(ring);
  if (tid == 0) {
    s_ctr = 0;
    for (int i = 0; i < nblk; ++i) {
      mbar_init(&full_bar[i], kThreads);
      mbar_init(&done_bar[i], kWarps);
    }
  }
  // pad columns x in [W, pitch) stay zero for the whole kernel (clamped / invalid taps read them)
  {
    const int padw = pitch - W;
    for (int i = tid; i < Rr * kCh * padw; i += kThreads) {
      const int rc = i / padw, x = W + (i - rc * padw);
      ring[(size_t)rc * pitch + x] = 0.0f;
    }
  }
  __syncthreads();

  // output scatter pattern of this lane: flat index lane + 32*j of a [32][7] tile -> channel c_j
  int ooff[kP];
  unsigned omask = 0u;
#pragma unroll
  for (int j = 0; j < kP; ++j) {
    const int idx = lane + 32 * j;
    const int c = idx / kP;
    ooff[j] = idx + c * (bins - kP);
    if (c < nch) omask |= 1u << j;
  }
  const uint32_t lane_base = ring_s + (uint32_t)lane * (uint32_t)pitch * 4u;
  const float4* recs = ws.records + (size_t)f * items_cap * kRecVec;
  const int n_items = ws.bucket_start[(size_t)f * (nb + 2) + nb];

  // work claiming: item index = part + split * (shared counter)
  auto claim = [&]() -> int {
    int v = 0;
    if (lane == 0) v = atomicAdd(&s_ctr, 1);
    v = __shfl_sync(0xffffffffu, v, 0);
    return part + split * v;
  };
  // ---------------- ring fill, cooperative: no producer warp.  Block b (St rows) may be written once
  // every warp has released bucket b - nblk.  Each warp issues ITS share of a block (channel-row
  // pairs warp, warp + 12, ...) as 4-byte cp.async and arrives on full_bar through them; it does so
  // whenever it looks (item boundaries and every wait loop), so loads never depend on a warp that
  // is itself waiting.  One warp cannot keep enough cp.async in flight to feed the SM; twelve can.
  int issued = 0;       // blocks whose share this warp has issued
  auto pump = [&]() {
    while (issued < nsteps) {
      const int b = issued;
      if (b >= nblk && !mbar_test(&done_bar[b % nblk], (uint32_t)(((b - nblk) / nblk) & 1))) break;
      const int y0 = b * St;
      const int prow = min(H, y0 + St) - y0;
      const uint32_t blk = ring_s + (uint32_t)((b % nblk) * St) * (uint32_t)(row_words * 4) + (uint32_t)lane * 4u;
      const float* src0 = fbase + (size_t)y0 * W + lane;
      for (int pr = warp; pr < kCh * prow; pr += kWarps) {
        const int c = pr & 31, dy = pr >> 5;
        if (c >= nch) continue;
        const float* src = src0 + ((size_t)c * H + dy) * W;
        const uint32_t dst = blk + (uint32_t)((dy * kCh + c) * pitch) * 4u;
        for (int x = lane, o = 0; x < W; x += 32, o += 32) cp_async4(dst + 4u * o, src + o);
      }
      cp_async_arrive(&full_bar[b % nblk]);
      ++issued;
    }
  };
  auto wait_on = [&](uint64_t* bar, uint32_t parity) {
    while (!mbar_test(bar, parity)) pump();
  };
  int cur = 0;          // buckets < cur are released by this warp
  int landed = 0;       // blocks < landed have been observed in the ring by this warp
  auto observe = [&](int upto) {
    for (; landed < upto; ++landed) wait_on(&full_bar[landed % nblk], (uint32_t)((landed / nblk) & 1));
  };
  // mbarrier parity waits are only meaningful within one phase of the barrier's current phase, so
  // every warp walks both barrier arrays strictly in order: it observes block j before it releases
  // bucket j (block j + nblk cannot land before that release, so full_bar is never two phases
  // ahead), and before arriving for bucket j it waits for the phase of bucket j - nblk (a warp that
  // skips far ahead on a sparse frame would otherwise be counted twice in that older phase).
  auto release = [&](int to) {
    for (int j = cur; j < to; ++j) {
      observe(j + 1);
      if (j >= nblk) wait_on(&done_bar[j % nblk], (uint32_t)(((j - nblk) / nblk) & 1));
      __syncwarp();
      if (lane == 0) mbar_arrive(&done_bar[j % nblk]);
    }
    cur = max(cur, to);
  };
  pump();
  int pending = claim();
  float4 rec_next = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pending < n_items) rec_next = __ldg(recs + (size_t)pending * kRecVec + lane);
  while (pending < n_items) {
    pump();
    slot[lane] = rec_next;
    __syncwarp();
    const int nxt = claim();
    if (nxt < n_items) rec_next = __ldg(recs + (size_t)nxt * kRecVec + lane);
    const float4 hdr = slot[0];
    const int r = __float_as_int(hdr.x), code = __float_as_int(hdr.y), bucket = __float_as_int(hdr.z);
    const int ph0 = code & 15, nph = (code >> 4) & 15, nrows = (code >> 8) & 255;
    float* o = out + ((size_t)r * C + c0) * bins + ph0 * kP;
    if (!(code >> 16)) {
      // release the buckets this warp has left behind, then make sure the item's blocks have landed
      release(bucket);
      observe(min(bucket + nbk, nsteps));
      switch (nph) {
        case 1: run_item<1, S>(slot, nrows, lane_base, stage, lane, o, ooff, omask); break;
        case 2: run_item<2, S>(slot, nrows, lane_base, stage, lane, o, ooff, omask); break;
        case 3: run_item<3, S>(slot, nrows, lane_base, stage, lane, o, ooff, omask); break;
        case 4: run_item<4, S>(slot, nrows, lane_base, stage, lane, o, ooff, omask); break;
        case 5: run_item<5, S>(slot, nrows, lane_base, stage, lane, o, ooff, omask); break;
        case 6: run_item<6, S>(slot, nrows, lane_base, stage, lane, o, ooff, omask); break;
        default: run_item<7, S>(slot, nrows, lane_base, stage, lane, o, ooff, omask); break;
      }
    } else {
      // bin-row taller than the resident window: taps straight from global memory (rare)
      const bool ch_ok = lane < nch;
      const float* roi = L.rois + (size_t)r * 5;
      const float rr[5] = {__ldg(roi), __ldg(roi + 1), __ldg(roi + 2), __ldg(roi + 3), __ldg(roi + 4)};
      const RoiGeom g = roi_geometry(rr, ws.scale, kP, kP, S, ws.aligned != 0);
      const float* plane = fbase + (size_t)(ch_ok ? lane : 0) * H * W;
      for (int pw = 0; pw < kP; ++pw) {
        float acc = 0.0f;
        for (int iy = 0; iy < S; ++iy) {
          const AxisTap ty = axis_tap(g.start_h, g.bin_h, ph0, iy, S, H);
          for (int ix = 0; ix < S; ++ix) {
            const AxisTap tx = axis_tap(g.start_w, g.bin_w, pw, ix, S, W);
            if (!(ty.ok && tx.ok)) continue;
            const float top = tx.wlo * __ldg(plane + ty.lo * W + tx.lo) + tx.whi * __ldg(plane + ty.lo * W + tx.hi);
            const float bot = tx.wlo * __ldg(plane + ty.hi * W + tx.lo) + tx.whi * __ldg(plane + ty.hi * W + tx.hi);
            acc += ty.wlo * top + ty.whi * bot;
          }
        }
        if (ch_ok) o[(size_t)lane * bins + pw] = acc / g.count;
      }
    }
    __syncwarp();
    pending = nxt;
  }
  // out of items: release every remaining bucket so the producer can finish
  release(nsteps);
  while (issued < nsteps) pump();
  asm volatile("cp.async.wait_all;" ::: "memory");
}

}  // namespace rows

size_t rows_workspace_bytes(int F, int H, int per_frame) { return rows::carve(nullptr, F, per_frame, H).bytes; }

// Returns B2D_ERR_UNSUPPORTED when this path does not apply (caller falls back).
int roi_align_forward_rows(int F, int C, int H, int W, const float* feat, const RoiList& L, int PH, int PW,
                           float scale, int S, int aligned, float* out, void* workspace, size_t workspace_bytes,
                           cudaStream_t st) {
  using namespace rows;
  if (PH != kP || PW != kP || S < 1 || S > 2 || L.n >= (1 << 27)) return B2D_ERR_UNSUPPORTED;
  const Plan p = make_plan(H, W);
  if (!p.ok) return B2D_ERR_UNSUPPORTED;
  const int per_frame = L.seg_count ? L.seg_stride : L.n;
  Ws ws = carve(workspace, F, per_frame, H);
  if (!workspace || workspace_bytes < ws.bytes) return B2D_ERR_UNSUPPORTED;
  ws.scale = scale;
  ws.aligned = aligned;
  const int items_cap = per_frame * kP;
  if (L.seg_count) {
    dim3 zg(L.seg_stride, F);
    zero_pad_kernel<<<zg, 256, 0, st>>>(L, C * PH * PW, out);
    B2D_LAUNCHED();
  }
  const int groups = ceil_div(C, kCh) * F;
  int split = 1;
  while (split < 4 && groups * split < 2 * kNumSMs) split *= 2;
  dim3 grid(ceil_div(C, kCh), F, split);
  const int nb = p.nsteps + 1;
#define B2D_ROWS(SS)                                                                                              \
  do {                                                                                                            \
    prep_kernel<SS><<<F, 256, sizeof(int) * 3 * nb, st>>>(L, H, W, scale, aligned, p.Rr, p.St, p.span_max,        \
                                                          p.nsteps, p.row_bytes, items_cap, ws);                  \
    B2D_LAUNCHED();                                                                                               \
    B2D_CUDA(cudaFuncSetAttribute(fwd_kernel<SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));     \
    fwd_kernel<SS><<<grid, kThreads, p.smem, st>>>(feat, L, C, H, W, p.pitch, p.St, p.nblk, p.nbk, p.nsteps,      \
                                                   items_cap, ws, out);                                           \
    B2D_LAUNCHED();                                                                                               \
  } while (0)
  if (S == 2) B2D_ROWS(2);
  else B2D_ROWS(1);
#undef B2D_ROWS
  return B2D_OK;
}

}  // namespace b2d
