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
//
// Data flow per CTA = (32-channel group, frame, band of rows - one band unless the call has few frames).  The CTA
// streams the feature rows of its band, 32 channels, top to bottom through a ring of nblk blocks of St rows
// (with few frames per call - the reference API issues ONE - the frame is cut into up to 4 bands so that 128
// CTAs work on it and each fills a quarter of the rows).  Lane c reads pixel
// (row, x) of channel c at word  slot*row_words + c*lane_stride + x  with an ODD lane_stride (W + 1 for
// even W), so the 32 lanes of a tap always hit 32 different banks.
//
//   * TMA fill (plane size % 4 == 0): two producer warps.  TMA cannot write that layout: bulk and tiled
//     copies land rows at 16-byte granularity (an odd pitch is impossible) and tiled box coordinates must
//     be 16-byte aligned in the innermost dimension too (an element skew per channel or per box raises an
//     illegal-instruction fault - measured twice).  So one tiled TMA per feature row (box = [32 planes][row],
//     starting on the aligned element at or before the row) lands in a small dense staging buffer and the
//     producers repack the staged row into its skewed ring slot with conflict-free LDS/STS - 2 wavefronts
//     per 32 elements.  (Register-staged LDG producers stall on address register reuse with only 6
//     scoreboards per warp; 4-byte cp.async from one warp sustains only ~8 copies in flight.)
//   * cp.async fill (other shapes): every warp issues its share of a block as 4-byte LDGSTS whenever
//     it looks (item boundaries and wait loops).
//
// Blocks are handed over with mbarriers only (full[b]: block landed; TMA variant: consumers publish the
// bucket they work in and the producers poll the minimum), there is no CTA-wide barrier after start-up:
// bucket k may touch blocks k .. k+nbk-1 and the remaining nblk-nbk blocks are slack, so warps drift
// apart by that many blocks.  A prep kernel turns the RoIs of a frame into self-contained item records
// sorted by first row; warps claim items from a shared counter and prefetch the next record while they
// compute.
//
// Output.  A whole-RoI item (nph = 7) of a full channel group is 6272 contiguous bytes of the output:
// the warp try-locks one of kPool shared tiles, writes its 49 values per lane there and ONE bulk copy
// (cp.async.bulk shared -> global) stores the slice - no LDS/STG round trip, half of all output bytes on
// the bench workload.  Everything else (items of RoIs split over several windows, ragged channel groups,
// the cp.async variant) is staged one bin-row [32 ch][7] at a time so that global stores run along (c, pw).
// Measured and rejected (round 2, B200, 64 frames): assembling split RoIs in their own output slice - every
// item stores its bin-rows transposed [bin][32 ch] with fully coalesced 128-byte stores, a per-RoI shared
// counter elects the last item, which reads the slice back from L2, transposes it through a pool tile and
// bulk-stores it in place.  The coalesced part stores alone run at 34.5 us/frame against 36.7 for the
// bin-row path, but the __threadfence() that must precede the count costs 7 us (it waits for the stores the
// warp has just issued) and the read-back + transpose 17-24 us (49 dependent L2 loads stall one of only 8
// consumer warps): 66 us/frame in total.  The scattered bin-row stores cost 2 us, not the 12 the round-1
// knock-out build suggested; what the kernel waits for is the shared-memory pipe (0.7 wavefronts/clk).
// Also measured and rejected: landing the TMA box in the ring slot itself and repacking it IN PLACE (no staging
// buffer: 13 ring rows instead of 11, window 10 + slack 3, every free slot with a copy in flight, the two
// producers on alternate rows): 37.2 us/frame against 35.8 - a row repacked by one warp takes twice as long to
// publish, and the deeper ring buys nothing because the consumers are not waiting for slack.
//
// Two things ptxas must be told (each cost 10 % when missed): (1) every branch on a warp-uniform value
// goes through a warp reduction (CREDUX -> uniform register); one branch it takes for divergent - the
// producer/consumer split on threadIdx.x >> 5, a plain `held >= 0` - and every tap of the consumers
// loses its [column + uniform row + imm] addressing to a per-tap IADD; (2) code size: the item code
// is instantiated for 2, 4 and 7 accumulator rows only - one variant per nph falls out of the instruction
// cache (4500+ instructions: 46 us/frame instead of 42).
#include <cuda.h>
#include <string.h>

#include "roi_common.cuh"

namespace b2d {

namespace rows {

// A/B build knobs (profiles/ only; the shipped library is built without any -D)
#ifndef B2D_STAGES
#define B2D_STAGES 2
#endif
#ifndef B2D_POOL
#define B2D_POOL 2
#endif
#ifndef B2D_SLACK
#define B2D_SLACK 1
#endif
#ifndef B2D_SLACK_BIG
#define B2D_SLACK_BIG 3
#endif
#ifndef B2D_WARPS
#define B2D_WARPS 12
#endif
#ifndef B2D_L2_AHEAD
#define B2D_L2_AHEAD 2
#endif
#define B2D_POLL_NS 64       // producers' back-off between two looks at the consumers' progress (256 / 1024 ns measured +0.4 % / +2.3 %)

#ifdef B2D_AB_COUNT
__device__ unsigned long long g_ab_count[8];
#define B2D_COUNT(i, v) do { if (lane == 0) atomicAdd(&g_ab_count[i], (unsigned long long)(v)); } while (0)
#else
#define B2D_COUNT(i, v) do { } while (0)
#endif

constexpr int kWarps = B2D_WARPS;          // 10 consumers + 2 producers (12 measured 1.2 % faster than 10: 35.8 vs 36.3 us/frame)
constexpr int kThreads = kWarps * 32;
constexpr int kCh = 32;            // channels per CTA (lanes).  (16-channel CTAs, measured in a skeleton build: the fill protocol costs
                                   // ~0.5 us per ROW whatever its size - 28.1 us per frame against 18.4)
constexpr int kP = 7;              // PH = PW = 7
constexpr int kRecVec = 32;        // float4 per item record
constexpr int kRecBytes = kRecVec * 16;
constexpr int kMaxRows = 10;       // distinct feature rows per item
constexpr int kXVec = 11;          // float4 holding 14 column taps {xo, hx, lx}
constexpr int kRowVec0 = 1 + kXVec;
constexpr int kMaxBlk = 64;        // ring blocks (mbarrier pairs)
constexpr int kMaxSplit = 4;       // CTAs sharing one (frame, channel group): each streams one band of rows
constexpr int kRowBlkShift = 18;   // row entries: ring byte offset (< 227 KB) in the low bits, block of the row above

constexpr int kPool = B2D_POOL;             // TMA fill: output tiles [32 ch][49] shared by the consumer warps (minimum)
constexpr int kMaxPool = 10;                // ... and as many as fit once the ring has kPoolRingRows rows (chosen by the plan);
                                            // one per consumer warp = private tiles: no locks, no wait after the store
constexpr int kPoolRingRows = 12;
// per consumer warp: its record slot (512 B), overlaid by its bin-row output tile [32 ch][7] (896 B) - the record is
// dead once the row loop of an item has finished
constexpr int kWarpAreaBytes = kCh * kP * 4 > kRecBytes ? kCh * kP * 4 : kRecBytes;
constexpr int kTileWords = (kCh * kP * kP + 31) / 32 * 32;      // (1568 for 32 channels; rounded so that what follows stays 128-byte aligned)
constexpr int kStages = B2D_STAGES;         // staging rows of the TMA fill
constexpr int kProducers = 2;               // producer warps of the TMA fill (each repacks 32 / kProducers channels)

struct Plan {
  int fill;         // 0: cooperative cp.async; 1: TMA + repack producer warp
  int lane_stride;  // words between channels of one slot (odd)
  int row_words;    // words per ring slot
  int Rr, St, nblk, nbk, span_max, span_whole, nsteps, npool;
  size_t smem;
  bool ok;
};

// Width of a staged row.  A tiled-TMA box must start on a 16-byte boundary of the plane, so when W is not a
// multiple of 4 the box starts up to 3 elements before the row and is 4 elements wider; the repack reads it
// with that per-row shift.
static int stage_width(int W) { return W % 4 == 0 ? W : (W + 3) / 4 * 4 + 4; }

static Plan make_plan(int H, int W, bool allow_tma) {
  Plan p{};
  p.fill = (allow_tma && (H * W) % 4 == 0 && stage_width(W) <= 256) ? 1 : 0;
  const int pitch = (W + 1) | 1;                     // smallest ODD value >= W + 1: any odd pitch maps 32 channels to 32 banks
  p.lane_stride = pitch;
  p.row_words = kCh * pitch;
  const size_t row_bytes = (size_t)p.row_words * 4;
  const size_t staging = p.fill ? (size_t)kStages * kCh * stage_width(W) * 4 : 0;
  const int nslot = p.fill ? kWarps - kProducers : kWarps;      // consumer warps own a record slot and a bin-row tile
  const size_t tile_bytes = (size_t)kTileWords * 4;
  const size_t budget = 227 * 1024 - 1280;       // 1.2 KB of static shared memory (barriers, locks, counters)
  // Pool tiles: every whole-RoI item that finds no free tile falls back to the scattered bin-row stores, so small
  // maps, whose ring does not need all of shared memory, get up to kMaxPool tiles (KITTI 24x78: -7 %, BEV 50x44:
  // -13 % with 6-8 tiles); a tile is only added while the ring keeps kPoolRingRows rows (or the whole map) -
  // on the 80x120 Waymo map a third tile costs a ring row and loses.
  p.npool = p.fill ? kPool : 0;
  auto fixed_for = [&](int npool) { return (size_t)nslot * kWarpAreaBytes + 256 + staging + (size_t)npool * tile_bytes; };
  while (p.fill && p.npool < kMaxPool) {
    const size_t f0 = fixed_for(p.npool), f1 = fixed_for(p.npool + 1);
    if (f1 + 6 * row_bytes > budget) break;
    const int rows0 = (int)((budget - f0) / row_bytes), rows1 = (int)((budget - f1) / row_bytes);
    // free (fits in the remainder below a whole ring row) or affordable
    if (rows1 < rows0 && rows1 < (H < kPoolRingRows ? H : kPoolRingRows)) break;
    ++p.npool;
  }
  const size_t fixed = fixed_for(p.npool);
  if (fixed + 6 * row_bytes > budget) { p.ok = false; return p; }
  int Rr = (int)((budget - fixed) / row_bytes);
  while (Rr > 6 && ((size_t)Rr * row_bytes) % 128 != 0) --Rr;      // what follows the ring stays 128-byte aligned (a no-op for 32 channels)
  if (Rr >= H) {
    p.Rr = H; p.St = H; p.nblk = 1; p.nbk = 1; p.span_max = H; p.nsteps = 1;
  } else {
    int St = Rr / 12;
    if (St < 1) St = 1;
    Rr -= Rr % St;                                   // blocks of St rows never wrap inside the ring
    const int nblk = Rr / St;
    const int slack = nblk >= 12 ? B2D_SLACK_BIG : (nblk >= 8 ? B2D_SLACK : 1);
    p.Rr = Rr; p.St = St; p.nblk = nblk; p.nbk = nblk - slack;
    p.span_max = (p.nbk - 1) * St + 1;
    p.nsteps = ceil_div(H, St);
  }
  if (p.span_max > kMaxRows) p.span_max = kMaxRows;
  // whole-RoI items may use all but two blocks of the ring (+ the one the producer fills)
  p.span_whole = p.nblk > 1 ? (p.nblk - 2 - 1) * p.St + 1 : p.span_max;
  if (p.span_whole > kMaxRows) p.span_whole = kMaxRows;
  if (p.span_whole < p.span_max) p.span_whole = p.span_max;
  p.smem = fixed + (size_t)p.Rr * row_bytes;
  p.ok = p.span_max >= 4 && p.nblk <= kMaxBlk;
  return p;
}

struct Ws {
  float4* records;        // [F][items_cap][kRecVec]: item q of list entry i at slot i * 7 + q (unsorted)
  int32_t* part_start;    // [F][kMaxSplit + 1]
  uint32_t* keys;         // [F][items_cap]: part * nb + bucket of the slot's item, ~0 for an unused slot
  int32_t* order;         // [F][items_cap]: slots grouped by part, sorted by first row within a part
  int32_t* ticket;        // [F]: prep CTAs of the frame that have finished (the last one sorts)
  float scale;
  int aligned;
  size_t bytes;
};

static Ws carve(void* base, int F, int per_frame) {
  Ws w{};
  size_t off = 0;
  char* p = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  w.records = reinterpret_cast<float4*>(take((size_t)F * per_frame * kP * kRecBytes));
  w.part_start = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F * (kMaxSplit + 1)));
  w.keys = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * (size_t)F * per_frame * kP));
  w.order = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F * per_frame * kP));
  w.ticket = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F));
  w.bytes = off;
  return w;
}

// ------------------------------------------------------------------------------------------
// Items of one RoI.  rf[ph] / rl[ph]: first / last feature row the samples of bin-row ph touch (rf > rl: none),
// all_ok: every sample row is valid.  Calls emit(ph0, nph, first_row, last_row, slow) for each item, in ph order;
// the items cover the 7 bin-rows exactly once.
// A RoI whose sample rows all fit in `span_whole` rows stays ONE item (it leaves through the bulk-store tile,
// which is worth a tighter window); otherwise it is cut into items of at most span_max rows.
template <class Emit>
__device__ __forceinline__ void for_each_item(const int (&rf)[kP], const int (&rl)[kP], bool all_ok, int H, int span_max,
                                              int span_whole, Emit emit) {
  if (span_whole > span_max && all_ok) {
    // (rows are monotone in ph: the first and the last bin-row bound the RoI; a RoI with invalid outer
    // samples is left to the general walk below)
    const int wf = min(rf[0], rf[kP - 1]), wl = max(rl[0], rl[kP - 1]);
    if (wl - wf + 1 > span_max && wl - wf + 1 <= span_whole) {
      emit(0, kP, wf, wl, false);
      return;
    }
  }
  int a = 0, cf = H, cl = -1;
#pragma unroll
  for (int ph = 0; ph < kP; ++ph) {
    const int nf = min(cf, rf[ph]), nl = max(cl, rl[ph]);
    if (ph > a && nl >= 0 && nl - nf + 1 > span_max) {
      emit(a, ph - a, cf, cl, false);
      a = ph;
      cf = rf[ph];
      cl = rl[ph];
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

// Record of one item (float4 units):
//   [0]       {roi row r, ph0 | nph << 4 | nrows << 8 | slow << 16, block of the first row (bucket) | block of the
//              last row << 16, 0}
//   [1..11]   14 column taps: 14 byte offsets of the lo column (the hi column is lo + 1), then the weights;
//             S == 2: per bin pair (2j, 2j+1) the weights of tap a / b of the first sample and c / d of the second
//             as aligned register pairs, then bin 6;  S == 1: {offset, hx, lx} triples
//   [12..]    per distinct feature row: {ring byte offset, wy[0..2]} (+ {wy[3..6]} when nph > 2);
//             wy[p] = weight of that row in bin-row ph0 + p, already divided by the sample count
// One WARP per list entry, kPrepWarps entries per CTA, any number of CTAs per frame.  The lanes of a warp hold the
// 14 row taps (lanes 0..13) and the 14 column taps (lanes 16..29) of the RoI - one axis_tap() each - and build the
// item records cooperatively: lane v writes float4 v of the 512-byte record, pulling what it needs from the tap
// lanes with shuffles.  Records are written unsorted (slot = entry * 7 + item), each with its sort key; the LAST
// CTA of a frame to finish (ticket) counting-sorts the frame's slots by (part, first row) into `order`.
// (One CTA per frame with a thread per RoI took 70 us at one frame per call and 250 us for 2000 RoIs: a single SM
// running ~10 000 serial instructions per RoI.)
constexpr int kPrepWarps = 8;
constexpr int kPrepThreads = kPrepWarps * 32;
constexpr int kPrepKeyCache = 4096;      // slots whose sort keys the sorting CTA keeps in shared memory (16 KB)

// word e of the column-tap block (11 float4 = 44 words) = field f of column tap k, packed k | f << 4
// (f: 0 byte offset of the lo column, 1 weight of the lo column, 2 weight of the hi column, 3 zero)
__host__ __device__ constexpr int xw_map(int S, int e) {
  if (S == 2) {
    if (e < 14) return e;
    if (e < 16) return 3 << 4;
    if (e < 40) {
      const int j = (e - 16) / 8, m = (e - 16) % 8;
      // wa = hx[4j], hx[4j+2]; wb = lx[4j], lx[4j+2]; wc = hx[4j+1], hx[4j+3]; wd = lx[4j+1], lx[4j+3]
      const int k = 4 * j + (m / 4) + 2 * (m % 2);
      const int f = ((m / 2) % 2) ? 2 : 1;
      return k | (f << 4);
    }
    return (12 + (e - 40) / 2) | ((((e - 40) % 2) ? 2 : 1) << 4);
  }
  if (e < 42) return ((e / 3) < 7 ? e / 3 : 0) | (((e / 3) < 7 ? e % 3 : 3) << 4);
  return 3 << 4;
}

template <int S>
__global__ void __launch_bounds__(kPrepThreads)
prep_kernel(RoiList L, int H, int W, float scale, int aligned, int Rr, int St, int span_max, int span_whole, int nsteps,
            int row_bytes, int per_frame, int split, int band_rows, Ws ws) {
  extern __shared__ int s_buckets[];   // last CTA of the frame: [split * nb] counts, offsets, fill
  __shared__ int s_last;
  constexpr int NT = kP * S;           // taps per axis
  pdl_trigger();
  pdl_wait();
  const int nb = nsteps + 1;
  const int nkey = split * nb;
  const int f = blockIdx.y;
  const int items_cap = per_frame * kP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int first = 0, n_ent = L.n;
  if (L.seg_count) {
    first = f * L.seg_stride;
    n_ent = L.seg_count[f];
  }
  float4* recs = ws.records + (size_t)f * items_cap * kRecVec;
  uint32_t* keys = ws.keys + (size_t)f * items_cap;
  const int i = blockIdx.x * kPrepWarps + warp;      // list entry of this warp
  if (i < per_frame) {
    int n_emitted = 0;
    bool mine = i < n_ent;
    int r = 0;
    float rr[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (mine) {
      const int e = first + i;
      r = L.ids ? L.ids[e] : e;
      const float* roi = L.rois + (size_t)r * 5;
#pragma unroll
      for (int c = 0; c < 5; ++c) rr[c] = __ldg(roi + c);
      if (!L.seg_count && (int)rr[0] != f) mine = false;
    }
    if (mine) {      // (warp-uniform)
      const RoiGeom g = roi_geometry(rr, scale, kP, kP, S, aligned != 0);
      const float inv_cnt = 1.0f / g.count;
      // this lane's tap: lanes [0, NT) rows, lanes [16, 16 + NT) columns; the others compute a dummy
      const bool is_x = lane >= 16;
      const int k_mine = min(is_x ? lane - 16 : lane, NT - 1);
      const AxisTap t = axis_tap(is_x ? g.start_w : g.start_h, is_x ? g.bin_w : g.bin_h, k_mine / S, k_mine % S, S,
                                 is_x ? W : H);
      const unsigned okm = __ballot_sync(0xFFFFFFFFu, t.ok);
      // row lanes carry weights already divided by the sample count
      const float wlo = is_x ? t.wlo : t.wlo * inv_cnt, whi = is_x ? t.whi : t.whi * inv_cnt;
      const unsigned yok = okm & ((1u << NT) - 1u);
      // row range of every bin-row (uniform in the warp)
      int rf[kP], rl[kP];
#pragma unroll
      for (int ph = 0; ph < kP; ++ph) {
        rf[ph] = H;
        rl[ph] = -1;
#pragma unroll
        for (int iy = 0; iy < S; ++iy) {
          const int k = ph * S + iy;
          const int lo_k = __shfl_sync(0xFFFFFFFFu, t.lo, k), hi_k = __shfl_sync(0xFFFFFFFFu, t.hi, k);
          if (yok & (1u << k)) {
            rf[ph] = min(rf[ph], lo_k);
            rl[ph] = max(rl[ph], hi_k);
          }
        }
      }
      const bool all_ok = yok == (1u << NT) - 1u;
      // float4 `lane` of the record, column-tap part (lanes 1..11), the same for every item of the RoI.  Invalid
      // samples carry zero weights and point at column 0 (always resident); at the clamped right border lx is 0
      // and the hi tap reads whatever follows the row (finite).
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
      {
        const float xoff = __int_as_float(t.ok ? t.lo * 4 : 0);
        float c[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int m = xw_map(S, min(max(4 * (lane - 1) + q, 0), 4 * kXVec - 1));
          const int src = 16 + (m & 15), fld = m >> 4;
          const float a0 = __shfl_sync(0xFFFFFFFFu, xoff, src), a1 = __shfl_sync(0xFFFFFFFFu, t.wlo, src),
                      a2 = __shfl_sync(0xFFFFFFFFu, t.whi, src);
          c[q] = fld == 0 ? a0 : (fld == 1 ? a1 : (fld == 2 ? a2 : 0.0f));
        }
        xv = make_float4(c[0], c[1], c[2], c[3]);
      }
      // the items of the RoI (at most 7), one per lane: ph0 | nph << 4 | slow << 8 | cf << 9 | cl << 20
      // (the walk is uniform in the warp; keeping the list in lanes avoids four inlined copies of the record code)
      int my_item = 0;
      for_each_item(rf, rl, all_ok, H, span_max, span_whole, [&](int ph0, int nph, int cf, int cl, bool slow) {
        const int packed = ph0 | (nph << 4) | ((slow ? 1 : 0) << 8) | (min(cf, 2047) << 9) | ((cl + 1) << 20);
        if (lane == n_emitted) my_item = packed;
        ++n_emitted;
      });
#pragma unroll 1
      for (int it = 0; it < n_emitted; ++it) {
        const int packed = __shfl_sync(0xFFFFFFFFu, my_item, it);
        const int ph0 = packed & 15, nph = (packed >> 4) & 15;
        const bool slow = (packed >> 8) & 1;
        const int cl = (packed >> 20) - 1;
        const int cf = cl < 0 ? H : (packed >> 9) & 2047;
        const int b = slow ? nsteps : (cl < 0 ? 0 : cf / St);
        // the CTA of band `part` streams rows [part * band_rows, (part + 1) * band_rows + window): an item belongs
        // to the band of its first row (items without a valid row go to band 0, slow items need no rows)
        const int part = cl < 0 ? 0 : min(split - 1, cf / band_rows);
        const int sl = i * kP + it;
        float4* rec = recs + (size_t)sl * kRecVec;
        // distinct feature rows of the item: the integers of [cf, cl] that one of its samples touches
        const bool in_item = !is_x && lane >= ph0 * S && lane < (ph0 + nph) * S && t.ok;
        unsigned touched = 0u;
        if (!slow && cl >= 0) {
          const unsigned mine_bits = in_item ? (1u << (t.lo - cf)) | (1u << (t.hi - cf)) : 0u;
          touched = __reduce_or_sync(0xFFFFFFFFu, mine_bits);
        }
        const int nrows = __popc(touched);
        const int rv = nph > 2 ? 2 : 1;                // matches the NPH variant the item runs on (2, 4 or 7)
        float4 out = xv;
        if (lane == 0) {
          const int code = ph0 | (nph << 4) | (nrows << 8) | ((slow ? 1 : 0) << 16);
          const int blocks = b | ((cl < 0 ? 0 : cl / St) << 16);
          out = make_float4(__int_as_float(r), __int_as_float(code), __int_as_float(blocks), 0.0f);
        }
        // lanes 12..31: row entry j = (lane - 12) / rv, half (lane - 12) % rv:
        //   half 0 {ring byte offset, wy[0..2]}, half 1 {wy[3..6]};  wy[p] = weight of the row in bin-row ph0 + p
        const int jl = max(lane - kRowVec0, 0);
        const int jrow = rv == 2 ? jl >> 1 : jl, half = rv == 2 ? jl & 1 : 0;
        int y = -1;                                    // the jrow-th touched row
#pragma unroll
        for (int bit = 0; bit < kMaxRows; ++bit)
          if (((touched >> bit) & 1u) && __popc(touched & ((1u << bit) - 1u)) == jrow) y = cf + bit;
        float wv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int p = half * 3 + c - (half ? 0 : 1);          // half 0: c = 1..3 -> p = 0..2; half 1: p = 3..6
          float w = 0.0f;
#pragma unroll
          for (int iy = 0; iy < S; ++iy) {
            const int k = min(max((ph0 + p) * S + iy, 0), NT - 1);
            const int lo_k = __shfl_sync(0xFFFFFFFFu, t.lo, k), hi_k = __shfl_sync(0xFFFFFFFFu, t.hi, k);
            const float wl_k = __shfl_sync(0xFFFFFFFFu, wlo, k), wh_k = __shfl_sync(0xFFFFFFFFu, whi, k);
            const bool use = p >= 0 && p < nph && (yok & (1u << k)) && y >= 0;
            if (use && lo_k == y) w += wl_k;
            if (use && hi_k == y && hi_k != lo_k) w += wh_k;    // clamped bottom border: hi == lo, its weight is 0
          }
          wv[c] = w;
        }
        if (lane >= kRowVec0) {
          int off = 0;
          if (y >= 0) off = ((Rr == H ? y : y % Rr) * row_bytes) | ((y / St) << kRowBlkShift);   // ring byte offset | block of the row
          out = half == 0 ? make_float4(__int_as_float(off), wv[1], wv[2], wv[3]) : make_float4(wv[0], wv[1], wv[2], wv[3]);
        }
        rec[lane] = out;
        if (lane == 0) keys[sl] = (uint32_t)(part * nb + b);
      }
    }
    if (lane >= n_emitted && lane < kP) keys[i * kP + lane] = 0xFFFFFFFFu;
  }
  // ---- the last CTA of the frame sorts the slots
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ws.ticket + f, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  int* cnt = s_buckets;
  int* offs = s_buckets + nkey;
  int* fill = s_buckets + 2 * nkey;
  for (int k = threadIdx.x; k < 3 * nkey; k += kPrepThreads) s_buckets[k] = 0;
  __syncthreads();
  // keys: global -> shared (independent loads, one L2 round trip), counted on the way
  uint32_t* skey = reinterpret_cast<uint32_t*>(s_buckets + 3 * nkey);
  const bool cached = items_cap <= kPrepKeyCache;
#pragma unroll 4
  for (int sl = threadIdx.x; sl < items_cap; sl += kPrepThreads) {
    const uint32_t key = __ldcg(keys + sl);
    if (cached) skey[sl] = key;
    if (key != 0xFFFFFFFFu) atomicAdd(&cnt[key], 1);
  }
  __syncthreads();
  if (warp == 0) {
    // exclusive prefix over the nkey counts: each lane sums a run of keys, a warp scan joins the runs
    const int per = (nkey + 31) / 32;
    const int k0 = min(lane * per, nkey), k1 = min(k0 + per, nkey);
    int sum = 0;
    for (int k = k0; k < k1; ++k) sum += cnt[k];
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += v;
    }
    int run = incl - sum;
    int32_t* ps = ws.part_start + (size_t)f * (kMaxSplit + 1);
    for (int k = k0; k < k1; ++k) {
      if (k % nb == 0) ps[k / nb] = run;
      offs[k] = run;
      run += cnt[k];
    }
    const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (lane >= split && lane <= kMaxSplit) ps[lane] = total;
  }
  __syncthreads();
  int32_t* order = ws.order + (size_t)f * items_cap;
#pragma unroll 4
  for (int sl = threadIdx.x; sl < items_cap; sl += kPrepThreads) {
    const uint32_t key = cached ? skey[sl] : __ldcg(keys + sl);
    if (key != 0xFFFFFFFFu) order[offs[key] + atomicAdd(&fill[key], 1)] = sl;
  }
}

// zero rows of padded list entries (seg mode), one block per (entry, frame)
__global__ void __launch_bounds__(256) zero_pad_kernel(RoiList L, int per_roi, float* __restrict__ out, int32_t* __restrict__ ticket,
                                                       int pdl) {
  if (pdl) {       // (one CTA per list entry, nearly all of which leave at once: with many frames the wait is not free)
    pdl_trigger();
    pdl_wait();
  }
  const int f = blockIdx.y, ri = blockIdx.x;
  if (ri == 0 && threadIdx.x == 0) ticket[f] = 0;       // (instead of a memset node, which would break the launch chain)
  if (ri < L.seg_count[f]) return;
  const int e = f * L.seg_stride + ri;
  const int r = L.ids ? L.ids[e] : e;
  float* o = out + (size_t)r * per_roi;
  for (int i = threadIdx.x; i < per_roi; i += blockDim.x) o[i] = 0.0f;
}

// ------------------------------------------------------------------------------------------
// 4-byte asynchronous copy global -> shared (LDGSTS), the fill of the non-TMA path.
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
// tiled 2-D bulk tensor copy global -> shared, completion counted in bytes on `bar`
// (box coordinates must be 16-byte aligned in the innermost dimension)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

// 1-D bulk copy shared -> global (16-byte aligned on both sides), tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Output tile pool of the TMA variant: kPool whole-RoI tiles [32 ch][49].  A whole-RoI item of a full channel
// group TRIES to take a tile (one pass over the locks, no waiting - a warp that blocked here would stop
// publishing its progress and could starve the fill the tile holders are waiting on): with a tile the 49 values
// of each lane go to shared memory once and leave with ONE bulk store of 6272 contiguous bytes; the tile stays
// locked until the copy has read it (released at the warp's next item).  Without one the item takes the
// bin-row path.
__device__ __forceinline__ int tile_try_acquire(int* locks, int npool, int warp, int lane) {
  unsigned t = 0;                          // tile index + 1, 0 = none
  if (lane == 0) {
    int i = warp % npool;
    for (int k = 0; k < npool && t == 0; ++k) {
      if (atomicCAS(&locks[i], 0, 1) == 0) t = (unsigned)i + 1u;
      if (++i == npool) i = 0;
    }
  }
  return (int)__reduce_max_sync(0xffffffffu, t) - 1;
}
__device__ __forceinline__ void tile_release(int* locks, int t, int lane) {
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    *reinterpret_cast<volatile int*>(&locks[t]) = 0;
  }
}

__device__ __forceinline__ float lds_at(uint32_t addr) {
#ifdef B2D_AB_NOTAPS
  return __uint_as_float(addr);       // A/B timing build: no tap loads
#endif
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));   // not volatile: taps of an item may reorder
  return v;
}

// One item: nph = NPH bin-rows of one RoI for this lane's channel.
// NPH is the variant (2, 4 or 7 accumulator rows), nph <= NPH the bin-rows the item really has: the rows in
// between carry zero weights and are not stored.
template <int NPH, int S, bool POOL, class Landed>
__device__ __forceinline__ void run_item(Landed&& wait_row, const float4* __restrict__ slot, int nrows, int nph, uint32_t lane_base,
                                         float* __restrict__ stage, int lane, float* __restrict__ o,
                                         const int (&ooff)[kP], unsigned omask,
                                         float* __restrict__ pool, int* locks, int npool, int warp, bool tile_out, int& held) {
  constexpr int NX = kP * S;
  constexpr int RV = NPH > 2 ? 2 : 1;
  constexpr bool PACK = S == 2;
  uint32_t xa[NX];
  float lx[NX], hx[NX];
  if (PACK) {
    // offsets only; the weights are read as pairs below
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float4 t = slot[1 + v];
      xa[4 * v] = lane_base + (uint32_t)__float_as_int(t.x);
      xa[4 * v + 1] = lane_base + (uint32_t)__float_as_int(t.y);
      if (v < 3) {
        xa[4 * v + 2] = lane_base + (uint32_t)__float_as_int(t.z);
        xa[4 * v + 3] = lane_base + (uint32_t)__float_as_int(t.w);
      }
    }
    const float4 l = slot[1 + 10];
    hx[12] = l.x, lx[12] = l.y, hx[13] = l.z, lx[13] = l.w;
  } else {
    float q[4 * kXVec];
#pragma unroll
    for (int v = 0; v < kXVec; ++v) {
      const float4 t = slot[1 + v];
      q[4 * v] = t.x;
      q[4 * v + 1] = t.y;
      q[4 * v + 2] = t.z;
      q[4 * v + 3] = t.w;
    }
#pragma unroll
    for (int k = 0; k < NX; ++k) {
      xa[k] = lane_base + (uint32_t)__float_as_int(q[3 * k]);
      hx[k] = q[3 * k + 1];
      lx[k] = q[3 * k + 2];
    }
  }
  // accumulators as register pairs (pw = 2j, 2j+1 in acc2[p][j]; pw = 6 in acc2[p][3].x): the contraction
  // below runs on packed fp32 FMAs (FFMA2, sm_100), which halves its issue slots
  float2 acc2[NPH][4];
#pragma unroll
  for (int p = 0; p < NPH; ++p)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[p][j] = make_float2(0.0f, 0.0f);
#define B2D_ACC(p, pw) (((pw) & 1) ? acc2[p][(pw) >> 1].y : acc2[p][(pw) >> 1].x)
  // x weights of the bin pairs (0,1), (2,3), (4,5): tap a, b of the first sample, c, d of the second
  float2 wa[PACK ? 3 : 1], wb[PACK ? 3 : 1], wc[PACK ? 3 : 1], wd[PACK ? 3 : 1];
  if (PACK) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float4 u = slot[1 + 4 + 2 * j], v = slot[1 + 5 + 2 * j];
      wa[j] = make_float2(u.x, u.y);
      wb[j] = make_float2(u.z, u.w);
      wc[j] = make_float2(v.x, v.y);
      wd[j] = make_float2(v.z, v.w);
    }
  }

  constexpr int NT = 2 * S * kP;      // taps per feature row
  auto load_row = [&](float (&t)[NT], uint32_t ro) {
#pragma unroll
    for (int pw = 0; pw < kP; ++pw) {
      if (S == 2) {
        t[4 * pw] = lds_at(xa[2 * pw] + ro);
        t[4 * pw + 1] = lds_at(xa[2 * pw] + ro + 4);
        t[4 * pw + 2] = lds_at(xa[2 * pw + 1] + ro);
        t[4 * pw + 3] = lds_at(xa[2 * pw + 1] + ro + 4);
      } else {
        t[2 * pw] = lds_at(xa[pw] + ro);
        t[2 * pw + 1] = lds_at(xa[pw] + ro + 4);
      }
    }
  };
  auto compute_row = [&](const float (&t)[NT], const float4& f0, const float4& f1) {
    float wy[kP];
    wy[0] = f0.y;
    wy[1] = f0.z;
    wy[2] = f0.w;
    if (RV == 2) {
      wy[3] = f1.x;
      if (NPH > 4) {
        wy[4] = f1.y;
        wy[5] = f1.z;
        wy[6] = f1.w;
      }
    }
    if (PACK) {
      // same operations per element as the scalar path (FMUL, then three FMAs; one FMA per accumulator)
      float2 T2[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T2[j] = __fmul2_rn(wa[j], make_float2(t[8 * j], t[8 * j + 4]));
        T2[j] = __ffma2_rn(wb[j], make_float2(t[8 * j + 1], t[8 * j + 5]), T2[j]);
        T2[j] = __ffma2_rn(wc[j], make_float2(t[8 * j + 2], t[8 * j + 6]), T2[j]);
        T2[j] = __ffma2_rn(wd[j], make_float2(t[8 * j + 3], t[8 * j + 7]), T2[j]);
      }
      float t6 = hx[12] * t[24];
      t6 = fmaf(lx[12], t[25], t6);
      t6 = fmaf(hx[13], t[26], t6);
      t6 = fmaf(lx[13], t[27], t6);
#pragma unroll
      for (int p = 0; p < NPH; ++p) {
        const float2 w2 = make_float2(wy[p], wy[p]);
#pragma unroll
        for (int j = 0; j < 3; ++j) acc2[p][j] = __ffma2_rn(w2, T2[j], acc2[p][j]);
        acc2[p][3].x = fmaf(wy[p], t6, acc2[p][3].x);
      }
      return;
    }
#pragma unroll
    for (int pw = 0; pw < kP; ++pw) {
      float v = hx[pw] * t[2 * pw];
      v = fmaf(lx[pw], t[2 * pw + 1], v);
#pragma unroll
      for (int p = 0; p < NPH; ++p) B2D_ACC(p, pw) = fmaf(wy[p], v, B2D_ACC(p, pw));
    }
  };
  // The row offset is the same in every lane; the reduction tells ptxas so (CREDUX -> uniform register),
  // which lets the taps use [column + uniform row + imm] addressing.  Reading row entries past the last
  // row stays inside shared memory; such entries are never used as addresses.
  auto row_off = [](const float4& f0) { return __reduce_max_sync(0xffffffffu, (uint32_t)__float_as_int(f0.x)); };
  // row entries are fetched one row ahead so that the LDS -> CREDUX -> tap-address chain of row i + 1
  // overlaps the taps of row i
  float4 e0 = slot[kRowVec0], e1 = slot[kRowVec0 + (RV == 2 ? 1 : 0)];
#pragma unroll 1
  for (int i = 0; i < nrows; ++i) {
#ifdef B2D_AB_NOROWS
    break;                            // A/B timing build: no row loop
#endif
    const float4 f0 = e0, f1 = e1;
    const uint32_t rv = row_off(f0);
    const uint32_t ro = rv & ((1u << kRowBlkShift) - 1u);
    wait_row((int)(rv >> kRowBlkShift));        // the row's block has landed (a test on a uniform counter unless it is new)
    e0 = slot[kRowVec0 + (i + 1) * RV];
    if (RV == 2) e1 = slot[kRowVec0 + (i + 1) * RV + 1];
    float t[NT];
    load_row(t, ro);
    compute_row(t, f0, f1);
  }
#ifdef B2D_AB_NOSTORE
  return;      // A/B timing build: results discarded
#endif
  if (POOL) {
    if (NPH == kP && tile_out && nph == kP) {
      // npool tiles for the 10 consumer warps.  With a tile per warp (small maps) the tile is private: no lock, and
      // the warp waits for its previous bulk store only here, an item later, when the copy has long read the tile.
      const bool priv = npool >= kWarps - kProducers;
      int t;
      if (priv) {
        t = warp;
        if (__reduce_max_sync(0xffffffffu, (unsigned)(held + 1)) != 0u) {
          if (lane == 0) bulk_wait_read();
          __syncwarp();
        }
      } else {
        t = tile_try_acquire(locks, npool, warp, lane);
      }
      B2D_COUNT(t >= 0 ? 0 : 1, 1);
      if (t >= 0) {
        float* tile = pool + (size_t)t * kTileWords;
#pragma unroll
        for (int p = 0; p < NPH; ++p)
#pragma unroll
          for (int pw = 0; pw < kP; ++pw) tile[lane * (kP * kP) + p * kP + pw] = B2D_ACC(p, pw);   // 49 words per lane: odd pitch
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk copy
        __syncwarp();
#ifndef B2D_AB_NOGLOBAL
        if (lane == 0) bulk_s2g(o, smem_u32(tile), (uint32_t)kTileWords * 4u);
#endif
        held = t;
        return;
      }
    }
  }
  // results: stage one bin-row [32 ch][7] at a time so that global stores run along (c, pw)
  B2D_COUNT(2, 1);
  B2D_COUNT(3, nph);
#pragma unroll
  for (int p = 0; p < NPH; ++p) {
    if (p >= nph) break;
    if (p > 0) __syncwarp();
#pragma unroll
    for (int pw = 0; pw < kP; ++pw) stage[lane * kP + pw] = B2D_ACC(p, pw);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kP; ++j)
#ifdef B2D_AB_NOGLOBAL
      if (omask & (1u << j)) asm volatile("" ::"f"(stage[lane + 32 * j]));     // A/B timing build: everything but the global stores
#else
      if (omask & (1u << j)) o[p * kP + ooff[j]] = stage[lane + 32 * j];
#endif
  }
  __syncwarp();
#undef B2D_ACC
}

struct KArgs {
  const float* feat;
  RoiList L;
  int C, H, W;
  int lane_stride, row_words, stage_w;
  int St, nblk, nbk, items_cap, band_rows, halo, npool;
  Ws ws;
  float* out;
};

template <int S, bool FILL>
__global__ void __launch_bounds__(kThreads, 1)
fwd_kernel(const __grid_constant__ KArgs a, const __grid_constant__ CUtensorMap tmap, const float* __restrict__ feat_g,
           const float4* __restrict__ records_g, const float* __restrict__ rois_g, float* __restrict__ out_g) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxBlk];   // block b landed
  __shared__ __align__(8) uint64_t done_bar[kMaxBlk];   // every consumer warp is past bucket j
  __shared__ __align__(8) uint64_t stg_bar[kStages];    // staging row landed (TMA bytes)
  __shared__ int s_ctr;
  __shared__ int s_tile_lock[kMaxPool];
  __shared__ int s_progress[kWarps];                    // TMA fill: bucket each consumer warp is working in
  constexpr int kConsumers = FILL ? kWarps - kProducers : kWarps;
  const int tid = threadIdx.x, lane = tid & 31;
  // warp-uniform for ptxas: with a per-thread warp index the producer branch below looks divergent and
  // every tap address of the consumers is rebuilt with an IADD instead of [column + uniform row] addressing
  const int warp = (int)__reduce_max_sync(0xffffffffu, (unsigned)(tid >> 5));
  const int f = blockIdx.y;
  const int c0 = blockIdx.x * kCh;
  const int C = a.C, H = a.H, W = a.W;
  const int nch = min(kCh, C - c0);
  const int part = blockIdx.z;
  constexpr int bins = kP * kP;
  const int St = a.St, nblk = a.nblk;
  // rows this CTA streams: its band plus the window of the last item that starts in it; block indices below are
  // RELATIVE to the band's first block (barrier slots and phases start at 0), ring slots stay absolute (the item
  // records address them as y % ring rows)
  const int y_begin = min(part * a.band_rows, H), y_end = min(H, y_begin + a.band_rows + a.halo);
  const int b0 = y_begin / St;
  const int nsteps = (y_end - y_begin + St - 1) / St;
  const int row_words = a.row_words;
  const float* fbase = feat_g + ((size_t)f * C + c0) * H * W;
  // dynamic shared: [ring (128-byte aligned)][record slots][bin-row tiles][pool][staging]
  // (pointer arithmetic on `smem` keeps the shared address space; a cast through an integer would
  // turn every slot / staging access into a generic load)
  float* ring = smem + (((128u - (smem_u32(smem) & 127u)) & 127u) >> 2);
  float* stage = ring + (size_t)St * nblk * row_words + (size_t)warp * (kWarpAreaBytes / 4);
  float4* slot = reinterpret_cast<float4*>(stage);
  float* pool = ring + (size_t)St * nblk * row_words + (size_t)kConsumers * (kWarpAreaBytes / 4);
  float* stg = pool + (FILL ? (size_t)a.npool * kTileWords : 0);
  const uint32_t ring_s = smem_u32(ring);
  if (tid < kWarps) s_progress[tid] = 0;
  if (tid < kMaxPool) s_tile_lock[tid] = 0;
  if (tid == 0) {
    s_ctr = 0;
    for (int i = 0; i < kStages; ++i) mbar_init(&stg_bar[i], 1);
    for (int i = 0; i < nblk; ++i) {
      mbar_init(&full_bar[i], FILL ? kProducers : kThreads);
      mbar_init(&done_bar[i], kConsumers);
    }
  }
  {
    // pad columns x in [W, pitch) are read by clamped taps with weight 0: keep them finite
    const int pitch = a.lane_stride, padw = pitch - W;
    for (int i = tid; i < St * nblk * kCh * padw; i += kThreads) {
      const int rc = i / padw, x = W + (i - rc * padw);
      ring[(size_t)rc * pitch + x] = 0.0f;
    }
  }
  pdl_wait();        // (few-frame calls chain their launches: the prologue above overlaps the prep kernel)
  __syncthreads();

  if (FILL && warp >= kConsumers) {
    // ---------------- producer: one tiled TMA per feature row (box = [32 planes][W]) into a dense
    // staging buffer, then repack the staged row into its skewed ring slot
    const int Ws = a.stage_w;                                     // staged row width (>= W, multiple of 4)
    const uint32_t row_tx = (uint32_t)kCh * (uint32_t)Ws * 4u;    // the box is always 32 planes (OOB planes / elements zero-filled)
    const int nchunk = (W + 31) / 32;
    const bool tail_ok = lane + 32 * (nchunk - 1) < W;
    const int plane0 = f * C + c0;
    const int pw_id = warp - kConsumers;                    // which producer
    constexpr int kChP = kCh / kProducers;                  // channels this producer repacks
    auto fetch = [&](int y) {        // row y of the 32 channels -> staging buffer (y - y_begin) % kStages
      if (pw_id == 0 && lane == 0) {
        const int sb = (y - y_begin) % kStages;
        uint64_t* bar = &stg_bar[sb];
        mbar_expect_tx(bar, row_tx);
        tma_load_2d(smem_u32(stg + (size_t)sb * kCh * Ws), &tmap, (y * W) & ~3, plane0, bar);
        if (B2D_L2_AHEAD > 0 && y + B2D_L2_AHEAD < y_end)      // pull a later row into L2: its TMA then pays L2 latency only
          asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&tmap),
                       "r"(((y + B2D_L2_AHEAD) * W) & ~3), "r"(plane0)
                       : "memory");
      }
    };
    if (B2D_L2_AHEAD > 0 && pw_id == 0 && lane == 0)
      for (int y = y_begin + kStages; y < y_begin + kStages + B2D_L2_AHEAD && y < y_end; ++y)
        asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&tmap), "r"((y * W) & ~3), "r"(plane0)
                     : "memory");
    for (int y = y_begin; y < y_begin + kStages && y < y_end; ++y) fetch(y);
    for (int y = y_begin; y < y_end; ++y) {
      const int babs = y / St, dy = y - babs * St;
      const int b = babs - b0;
      const int sb = (y - y_begin) % kStages;
      if (dy == 0 && b >= nblk) {
        // the slot is free once every consumer warp works in a bucket beyond b - nblk
        for (;;) {
          int pr = lane < kConsumers ? *reinterpret_cast<volatile int*>(&s_progress[lane]) : 0x7fffffff;
          pr = (int)__reduce_min_sync(0xffffffffu, (unsigned)pr);
          if (pr > b - nblk) break;
          __nanosleep(B2D_POLL_NS);
        }
      }
      mbar_wait(&stg_bar[sb], (uint32_t)(((y - y_begin) / kStages) & 1));
      const uint32_t sstep = (uint32_t)Ws * 4u, dstep = (uint32_t)a.lane_stride * 4u;
      const uint32_t src = smem_u32(stg + (size_t)sb * kCh * Ws) + (uint32_t)(lane + ((y * W) & 3)) * 4u +
                           (uint32_t)(pw_id * kChP) * sstep;
      const uint32_t dst = ring_s + (uint32_t)((babs % nblk) * St + dy) * (uint32_t)(row_words * 4) + (uint32_t)lane * 4u +
                           (uint32_t)(pw_id * kChP) * dstep;
      // batches of 8 channels x up to 4 chunks (32 values per lane).  One address register per channel, the chunk
      // offset as an immediate, the chunk predicates once per batch: the address arithmetic of a per-element
      // formulation was two thirds of the repack's instructions.
      const int ncg = (nchunk + 3) / 4, nbat = (kChP / 8) * ncg;
#define B2D_LD4(J, SJ)                                                                                         \
  if (p0) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[(J) * 4 + 0]) : "r"(SJ) : "memory");                 \
  if (p1) asm volatile("ld.shared.f32 %0, [%1+128];" : "=f"(v[(J) * 4 + 1]) : "r"(SJ) : "memory");             \
  if (p2) asm volatile("ld.shared.f32 %0, [%1+256];" : "=f"(v[(J) * 4 + 2]) : "r"(SJ) : "memory");             \
  if (p3) asm volatile("ld.shared.f32 %0, [%1+384];" : "=f"(v[(J) * 4 + 3]) : "r"(SJ) : "memory");
#define B2D_ST4(J, DJ)                                                                                         \
  if (p0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(DJ), "f"(v[(J) * 4 + 0]) : "memory");                   \
  if (p1) asm volatile("st.shared.f32 [%0+128], %1;" ::"r"(DJ), "f"(v[(J) * 4 + 1]) : "memory");               \
  if (p2) asm volatile("st.shared.f32 [%0+256], %1;" ::"r"(DJ), "f"(v[(J) * 4 + 2]) : "memory");               \
  if (p3) asm volatile("st.shared.f32 [%0+384], %1;" ::"r"(DJ), "f"(v[(J) * 4 + 3]) : "memory");
      for (int bi = 0; bi < nbat; ++bi) {
#ifdef B2D_AB_NOREPACK
        break;                          // A/B timing build: TMA + barrier protocol only
#endif
        const int c8 = (bi / ncg) * 8, cg = (bi - (bi / ncg) * ncg) * 4;
        const bool p0 = cg + 0 < nchunk - 1 || (cg + 0 == nchunk - 1 && tail_ok);
        const bool p1 = cg + 1 < nchunk - 1 || (cg + 1 == nchunk - 1 && tail_ok);
        const bool p2 = cg + 2 < nchunk - 1 || (cg + 2 == nchunk - 1 && tail_ok);
        const bool p3 = cg + 3 < nchunk - 1 || (cg + 3 == nchunk - 1 && tail_ok);
        float v[32];
        uint32_t sj = src + (uint32_t)c8 * sstep + 128u * (uint32_t)cg;
        uint32_t dj = dst + (uint32_t)c8 * dstep + 128u * (uint32_t)cg;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          B2D_LD4(j, sj)
          sj += sstep;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          B2D_ST4(j, dj)
          dj += dstep;
        }
      }
#undef B2D_LD4
#undef B2D_ST4
      __syncwarp();
      if (kProducers > 1) asm volatile("bar.sync 1, %0;" ::"n"(kProducers * 32) : "memory");   // all producers done with the buffer
      if (y + kStages < y_end) fetch(y + kStages);     // this staging buffer is free again
      if (dy == St - 1 || y == y_end - 1) {
        if (lane == 0) mbar_arrive(&full_bar[b % nblk]);
      }
    }
    return;
  }

  // ---------------- consumers
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
  const uint32_t lane_base = ring_s + (uint32_t)lane * (uint32_t)a.lane_stride * 4u;
  // bulk stores need 16-byte aligned RoI slices: C * 49 * 4 bytes per RoI -> C % 4 == 0
  const bool tile_out = FILL && nch == kCh && (C & 3) == 0 && (reinterpret_cast<uintptr_t>(out_g) & 15u) == 0;
  int held = -1;        // pool tile a bulk store of this warp may still be reading
  const float4* recs = records_g + (size_t)f * a.items_cap * kRecVec;
  const int32_t* order = a.ws.order + (size_t)f * a.items_cap;
  const int item0 = (int)__reduce_max_sync(0xffffffffu, (unsigned)a.ws.part_start[(size_t)f * (kMaxSplit + 1) + part]);
  const int n_items = (int)__reduce_max_sync(0xffffffffu, (unsigned)a.ws.part_start[(size_t)f * (kMaxSplit + 1) + part + 1]);

  // work claiming: item index = first item of this CTA's part + shared counter
  auto claim = [&]() -> int {
    unsigned v = 0;
    if (lane == 0) v = (unsigned)atomicAdd(&s_ctr, 1);
    v = __reduce_max_sync(0xffffffffu, v);        // broadcast that ptxas knows to be uniform
    return item0 + (int)v;
  };

  // cp.async fill, cooperative: block b may be written once every warp has released bucket b - nblk.
  // Each warp issues ITS share of a block (channel-row pairs warp, warp + 12, ...) and arrives on
  // full_bar through the copies; it does so whenever it looks (item boundaries and every wait
  // loop), so loads never depend on a warp that is itself waiting.  (One warp cannot keep enough
  // LDGSTS in flight to feed the SM.)
  int issued = 0;       // blocks whose share this warp has issued
  auto pump = [&]() {
    if (FILL) return;
    while (issued < nsteps) {
      const int b = issued;
      if (b >= nblk && !mbar_test(&done_bar[b % nblk], (uint32_t)(((b - nblk) / nblk) & 1))) break;
      const int y0 = (b + b0) * St;
      const int prow = min(y_end, y0 + St) - y0;
      const uint32_t blk = ring_s + (uint32_t)(((b + b0) % nblk) * St) * (uint32_t)(row_words * 4) + (uint32_t)lane * 4u;
      const float* src0 = fbase + (size_t)y0 * W + lane;
      for (int pr = warp; pr < kCh * prow; pr += kWarps) {
        const int c = pr & 31, dy = pr >> 5;
        if (c >= nch) continue;
        const float* src = src0 + ((size_t)c * H + dy) * W;
        const uint32_t dst = blk + (uint32_t)((dy * kCh + c) * a.lane_stride) * 4u;
        for (int x = lane, o = 0; x < W; x += 32, o += 32) cp_async4(dst + 4u * o, src + o);
      }
      cp_async_arrive(&full_bar[b % nblk]);
      ++issued;
    }
  };
  auto wait_on = [&](uint64_t* bar, uint32_t parity) {
    if (FILL) {
      mbar_wait(bar, parity);       // (a nanosleep back-off here measured +0.3 %)
    } else {
      while (!mbar_test(bar, parity)) pump();
    }
  };
  int cur = 0;          // buckets < cur are released by this warp
  int landed = 0;       // blocks < landed have been observed in the ring by this warp
  auto observe = [&](int upto) {
    // the single producer arrives block after block, so block b landed implies all earlier ones; a
    // parity test is valid as long as the block one phase earlier (b - nblk) is known to have landed
    if (FILL && upto > landed && upto - 1 - nblk < landed) landed = upto - 1;
    for (; landed < upto; ++landed) wait_on(&full_bar[landed % nblk], (uint32_t)((landed / nblk) & 1));
  };
  // TMA fill: an item does not wait for its window.  Each row entry names its block and the row loop waits for a
  // block the first time this warp meets it, so an item starts as soon as its first row is there and consumes the
  // rest as the fill delivers them (blocks land in order; `landed` walks them one by one, which also keeps every
  // parity test within one phase of the barrier).
  int land_slot = 0;
  uint32_t land_phase = 0;
  auto wait_row = [&](int blk_abs) {
    if (!FILL) return;
    const int blk = blk_abs - b0;
    while (landed <= blk) {
      mbar_wait(&full_bar[land_slot], land_phase);
      ++landed;
      if (++land_slot == nblk) {
        land_slot = 0;
        land_phase ^= 1u;
      }
    }
  };
  // mbarrier parity waits are only meaningful within one phase of the barrier's current phase, so
  // every warp walks both barrier arrays strictly in order: it observes block j before it releases
  // bucket j (block j + nblk cannot land before that release, so full_bar is never two phases
  // ahead), and before arriving for bucket j it waits for the phase of bucket j - nblk (a warp that
  // skips far ahead on a sparse frame would otherwise be counted twice in that older phase).
  auto release = [&](int to) {
    if (FILL) {
      // TMA fill: publish the bucket this warp now works in; the producer polls the minimum.  All
      // taps of earlier items have been consumed by the time this store issues.
      if (to > cur) {
        cur = to;
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          *reinterpret_cast<volatile int*>(&s_progress[warp]) = to;
        }
      }
      return;
    }
    for (int j = cur; j < to; ++j) {
      observe(j + 1);
      if (j >= nblk) wait_on(&done_bar[j % nblk], (uint32_t)(((j - nblk) / nblk) & 1));
      __syncwarp();
      if (lane == 0) mbar_arrive(&done_bar[j % nblk]);
    }
    cur = max(cur, to);
  };
  // Between items: give back the pool tile of the previous bulk store.
  auto between_items = [&]() {
    // (`held` is warp-uniform; the reduction tells ptxas so - a branch it takes for divergent costs the
    // consumers their [column + uniform row] tap addressing)
    if (FILL && a.npool < kConsumers && __reduce_max_sync(0xffffffffu, (unsigned)(held + 1)) != 0u) {
      // the bulk store issued by the previous item has (nearly always) read its tile by now
      if (lane == 0) bulk_wait_read();
      tile_release(s_tile_lock, held, lane);
      held = -1;
    }
  };
  pump();
  // Claimed item -> record slot through the frame's sorted `order` list.  Each warp keeps a 64-entry window of the
  // list in registers (lane l: entries win_base + l and win_base + 32 + l, the second half prefetched), so the slot
  // of a freshly claimed item is one shuffle away and its record load can be issued at once - claiming further
  // ahead instead lets a warp sit on an early item while the others run on, which holds the ring back (+2.5 %).
  // (loads are unconditional - clamped index - so that a prefetch is not followed by a select that would wait
  // for it; a frame without items leaves `order` unwritten: slots are clamped into the record array)
  const int last_item = max(n_items - 1, 0);
  const unsigned last_slot = (unsigned)max(a.items_cap - 1, 0);
  int win_base = item0;
  int win0 = __ldg(order + min(win_base + lane, last_item)), win1 = __ldg(order + min(win_base + 32 + lane, last_item));
  auto slot_of = [&](int v) -> unsigned {            // v: warp-uniform claimed index
    v = min(v, last_item);
    if (v >= win_base + 64) {                        // the other warps ran far ahead meanwhile (rare): reload
      win_base += (v - win_base) & ~31;
      win0 = __ldg(order + min(win_base + lane, last_item));
      win1 = __ldg(order + min(win_base + 32 + lane, last_item));
    } else if (v >= win_base + 32) {
      win_base += 32;
      win0 = win1;
      win1 = __ldg(order + min(win_base + 32 + lane, last_item));
    }
    return min((unsigned)__shfl_sync(0xffffffffu, win0, v - win_base), last_slot);
  };
  int pending = claim();
  float4 rec_next = __ldg(recs + (size_t)slot_of(pending) * kRecVec + lane);

  while (pending < n_items) {
    pump();
    between_items();
    slot[lane] = rec_next;
    __syncwarp();
    const int nxt = claim();
    rec_next = __ldg(recs + (size_t)slot_of(nxt) * kRecVec + lane);
    const float4 hdr = slot[0];
    // header fields are warp-uniform; the reductions make that visible to ptxas (uniform branches / loops)
    const int r = (int)__reduce_max_sync(0xffffffffu, (uint32_t)__float_as_int(hdr.x));
    const int code = (int)__reduce_max_sync(0xffffffffu, (uint32_t)__float_as_int(hdr.y));
    const int blocks = (int)__reduce_max_sync(0xffffffffu, (uint32_t)__float_as_int(hdr.z));
    const int ph0 = code & 15, nph = (code >> 4) & 15, nrows = (code >> 8) & 255;
    const int bucket = (blocks & 0xFFFF) - b0, last_blk = (blocks >> 16) - b0;
    float* o = out_g + ((size_t)r * C + c0) * bins + ph0 * kP;
    if (!((code >> 16) & 1)) {
      // release the buckets this warp has left behind, then make sure the item's blocks have landed
      // (only the blocks the item reads have to be there, not its whole window)
      release(bucket);
      if (!FILL) observe(min(last_blk + 1, nsteps));
#define B2D_RUN(N) run_item<N, S, FILL>(wait_row, slot, nrows, nph, lane_base, stage, lane, o, ooff, omask, pool, s_tile_lock, a.npool, warp, tile_out, held)
      if (nph <= 2) B2D_RUN(2);
      else if (nph <= 4) B2D_RUN(4);
      else B2D_RUN(7);
#undef B2D_RUN
    } else {
      // bin-row taller than the resident window: taps straight from global memory (rare)
      const bool ch_ok = lane < nch;
      const float* roi = rois_g + (size_t)r * 5;
      const float rr[5] = {__ldg(roi), __ldg(roi + 1), __ldg(roi + 2), __ldg(roi + 3), __ldg(roi + 4)};
      const RoiGeom g = roi_geometry(rr, a.ws.scale, kP, kP, S, a.ws.aligned != 0);
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
#ifndef B2D_AB_NOSTORE
        if (ch_ok) o[(size_t)lane * bins + pw] = acc / g.count;
#endif
      }
    }
    __syncwarp();
    pending = nxt;
  }
  // out of items: release every remaining bucket so the fill can finish
  release(nsteps);
  between_items();      // the last bulk store still reads shared memory
  if (FILL && a.npool >= kConsumers && held >= 0) {
    if (lane == 0) bulk_wait_read();
    __syncwarp();
  }
  if (!FILL) {
    while (issued < nsteps) pump();
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// feature planes as a 2-D tensor [F*C planes][H*W elements]; box = one row of 32 consecutive planes
static bool make_tmap(CUtensorMap* map, const float* feat, int F, int C, int H, int W) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || (reinterpret_cast<uintptr_t>(feat) & 15u) || ((size_t)H * W) % 4 != 0 || stage_width(W) > 256) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)H * W, (cuuint64_t)F * C};
  const cuuint64_t strides[1] = {(cuuint64_t)H * W * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)stage_width(W), (cuuint32_t)kCh};
  const cuuint32_t estr[2] = {1u, 1u};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(feat), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace rows

#ifdef B2D_AB_COUNT
extern "C" void b2d_ab_counters(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, rows::g_ab_count, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(rows::g_ab_count, z, sizeof(z));
  }
}
#endif

size_t rows_workspace_bytes(int F, int /*H*/, int per_frame) { return rows::carve(nullptr, F, per_frame).bytes; }

// Returns B2D_ERR_UNSUPPORTED when this path does not apply (caller falls back).
int roi_align_forward_rows(int F, int C, int H, int W, const float* feat, const RoiList& L, int PH, int PW,
                           float scale, int S, int aligned, bool coop_fill, float* out, void* workspace,
                           size_t workspace_bytes, cudaStream_t st) {
  using namespace rows;
  if (PH != kP || PW != kP || S < 1 || S > 2 || L.n >= (1 << 24) || H >= 2047) return B2D_ERR_UNSUPPORTED;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  const int per_frame = L.seg_count ? L.seg_stride : L.n;
  const Plan p = make_plan(H, W, !coop_fill && make_tmap(&tmap, feat, F, C, H, W));
  if (!p.ok || p.nsteps >= 0xFFFF) return B2D_ERR_UNSUPPORTED;
  Ws ws = carve(workspace, F, per_frame);
  if (!workspace || workspace_bytes < ws.bytes) return B2D_ERR_UNSUPPORTED;
  ws.scale = scale;
  ws.aligned = aligned;
  const int items_cap = per_frame * kP;
  const int groups = ceil_div(C, kCh) * F;
  int split = 1;
  while (split < kMaxSplit && groups * split < 2 * kNumSMs) split *= 2;
  // few frames per call: the launches are chained programmatically (every kernel waits for its predecessor on the
  // device), which hides the launch gaps - a fifth of a one-frame call
  const bool pdl = split > 1;
  if (L.seg_count) {
    dim3 zg(L.seg_stride, F);
    B2D_CUDA(launch_pdl(zero_pad_kernel, zg, dim3(256), 0, st, pdl, L, C * PH * PW, out, ws.ticket, pdl ? 1 : 0));
    B2D_LAUNCHED();
  } else {
    B2D_CUDA(cudaMemsetAsync(ws.ticket, 0, sizeof(int32_t) * (size_t)F, st));
  }
  dim3 grid(ceil_div(C, kCh), F, split);
  const int nb = p.nsteps + 1;
  // split > 1 (few frames): each of the `split` CTAs of a (frame, channel group) streams only its band of rows
  const int band_rows = ceil_div(ceil_div(H, split), p.St) * p.St;
  const int halo = (p.span_whole > p.span_max ? p.span_whole : p.span_max) - 1;
  KArgs a{feat, L, C, H, W, p.lane_stride, p.row_words, stage_width(W), p.St, p.nblk, p.nbk, items_cap, band_rows, halo, p.npool, ws, out};
#define B2D_ROWS(SS, FF)                                                                                          \
  do {                                                                                                            \
    B2D_CUDA(launch_pdl(prep_kernel<SS>, dim3(ceil_div(per_frame, kPrepWarps), F), dim3(kPrepThreads),             \
                        sizeof(int) * (3 * split * nb + (items_cap <= kPrepKeyCache ? items_cap : 0)), st, pdl, L, H, W, scale, \
                        aligned, p.Rr, p.St, p.span_max, p.span_whole, p.nsteps, p.row_words * 4, per_frame, split,   \
                        band_rows, ws));                                                                          \
    B2D_LAUNCHED();                                                                                               \
    B2D_CUDA(cudaFuncSetAttribute(fwd_kernel<SS, FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem)); \
    B2D_CUDA(launch_pdl(fwd_kernel<SS, FF>, grid, dim3(kThreads), p.smem, st, pdl, a, tmap, feat,                  \
                        (const float4*)ws.records, L.rois, out));                                                 \
    B2D_LAUNCHED();                                                                                               \
  } while (0)
#define B2D_ROWS_S(FF)           \
  do {                           \
    if (S == 2) B2D_ROWS(2, FF); \
    else B2D_ROWS(1, FF);        \
  } while (0)
  if (p.fill) B2D_ROWS_S(true);
  else B2D_ROWS_S(false);
#undef B2D_ROWS_S
#undef B2D_ROWS
  return B2D_OK;
}

}  // namespace b2d
