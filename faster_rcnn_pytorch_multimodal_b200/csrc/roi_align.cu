// RoIAlign forward / backward, NCHW fp32, torchvision.ops.roi_align semantics.
//
// Replaces Network._crop_pool_layer (lib/nets/network.py is missing from the reference
// snapshot) and torchvision.ops.roi_align at utils/torchpoolers.py:165,194.
// The sample arithmetic restates torchvision csrc/ops/cpu/roi_align_kernel.cpp
// (pre_calc_for_bilinear_interpolate) operation by operation, without FMA contraction.
//
// Forward, "plane-resident" kernel: one CTA owns CPB consecutive channel planes of one
// frame.  NCHW makes those planes ONE contiguous run in HBM, so a single elected thread
// pulls them into shared memory with 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx);
// every feature byte is read from HBM exactly once per frame.  The CTA then walks all
// RoIs of its frame; a thread owns one (roi, bin), keeps that bin's sample geometry in
// registers and reuses it for the CPB channels, so the 4-tap gathers hit shared memory.
// Planes that do not fit in shared memory (FPN p2/p3) take the gather kernel that reads
// through L1/L2 instead.
//
// Backward, "band-owned" kernel: one CTA owns (32 channels) x (a band of rows) of
// grad_feat as shared-memory accumulators laid out [row][x][33] (channel fastest, padded),
// one thread per channel.  A thread is the only writer of its channel plane, RoIs are
// visited in index order, so accumulation is deterministic with no atomics at all; each
// grad_feat element is written to HBM exactly once.
#include "roi_common.cuh"

namespace b2d {

size_t rows_workspace_bytes(int F, int H, int per_frame);
size_t bwd_rows_workspace_bytes(int n_list);
int roi_align_backward_rows(int F, int C, int H, int W, const float* grad_out, const RoiList& L, int PH, int PW,
                            float scale, int S, int aligned, int accumulate, float* grad_feat, void* workspace,
                            size_t workspace_bytes, cudaStream_t st);
int roi_align_forward_rows(int F, int C, int H, int W, const float* feat, const RoiList& L, int PH, int PW,
                           float scale, int S, int aligned, bool coop_fill, float* out, void* workspace,
                           size_t workspace_bytes, cudaStream_t st);

// ------------------------------------------------------------------------------------------
template <int CPB>
__global__ void __launch_bounds__(512, 1)
roi_align_fwd_planes_kernel(const float* __restrict__ feat, RoiList L, int C, int H, int W, int PH, int PW, float scale,
                            int sampling_ratio, int aligned, int use_bulk, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* planes = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bar;

  const int f = blockIdx.y;
  const int c0 = blockIdx.x * CPB;
  const int nch = min(CPB, C - c0);
  const int HW = H * W;
  const float* src = feat + ((size_t)f * C + c0) * HW;

  if (use_bulk) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      mbar_expect_tx(&bar, (uint32_t)(nch * HW * 4));
      for (int c = 0; c < nch; ++c) bulk_g2s(planes + (size_t)c * HW, src + (size_t)c * HW, (uint32_t)(HW * 4), &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
  } else {
    for (int i = threadIdx.x; i < nch * HW; i += blockDim.x) planes[i] = __ldg(src + i);
    __syncthreads();
  }

  const int bins = PH * PW;
  int first = 0, n_roi = L.n, n_valid = L.n;
  if (L.seg_count) {
    first = f * L.seg_stride;
    n_roi = L.seg_stride;
    n_valid = L.seg_count[f];
  }
  const int items = n_roi * bins;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int ri = it / bins;
    const int bin = it - ri * bins;
    const int e = first + ri;
    const int r = L.ids ? L.ids[e] : e;
    float acc[CPB];
#pragma unroll
    for (int c = 0; c < CPB; ++c) acc[c] = 0.0f;
    bool write = true;
    if (ri < n_valid) {
      const float* roi = L.rois + (size_t)r * 5;
      if (!L.seg_count && (int)roi[0] != f) {
        write = false;
      } else {
        const int ph = bin / PW, pw = bin - ph * PW;
        const RoiGeom g = roi_geometry(roi, scale, PH, PW, sampling_ratio, aligned != 0);
        for (int iy = 0; iy < g.grid_h; ++iy) {
          const AxisTap ty = axis_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H);
          for (int ix = 0; ix < g.grid_w; ++ix) {
            const AxisTap tx = axis_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W);
            if (!(ty.ok && tx.ok)) continue;
            const float w1 = fmul(ty.wlo, tx.wlo), w2 = fmul(ty.wlo, tx.whi);
            const float w3 = fmul(ty.whi, tx.wlo), w4 = fmul(ty.whi, tx.whi);
            const int o1 = ty.lo * W + tx.lo, o2 = ty.lo * W + tx.hi, o3 = ty.hi * W + tx.lo, o4 = ty.hi * W + tx.hi;
#pragma unroll
            for (int c = 0; c < CPB; ++c) {
              if (c < nch) {
                const float* p = planes + c * HW;
                const float v = fadd(fadd(fadd(fmul(w1, p[o1]), fmul(w2, p[o2])), fmul(w3, p[o3])), fmul(w4, p[o4]));
                acc[c] = fadd(acc[c], v);
              }
            }
          }
        }
#pragma unroll
        for (int c = 0; c < CPB; ++c) acc[c] = fdiv(acc[c], g.count);
      }
    }
    if (write) {
      float* o = out + ((size_t)r * C + c0) * bins + bin;
#pragma unroll
      for (int c = 0; c < CPB; ++c)
        if (c < nch) o[(size_t)c * bins] = acc[c];
    }
  }
}

// Gather kernel for planes that do not fit in shared memory: one thread per output element.
__global__ void __launch_bounds__(256)
roi_align_fwd_gather_kernel(const float* __restrict__ feat, RoiList L, int F, int C, int H, int W, int PH, int PW,
                            float scale, int sampling_ratio, int aligned, float* __restrict__ out) {
  const int bins = PH * PW;
  const long long total = (long long)L.n * C * bins;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int bin = (int)(idx % bins);
    const int c = (int)((idx / bins) % C);
    const int e = (int)(idx / ((long long)bins * C));
    const int r = L.ids ? L.ids[e] : e;
    float val = 0.0f;
    bool valid = true;
    int f = 0;
    if (L.seg_count) {
      f = e / L.seg_stride;
      valid = (e - f * L.seg_stride) < L.seg_count[f];
    }
    if (valid) {
      const float* roi = L.rois + (size_t)r * 5;
      if (!L.seg_count) f = (int)roi[0];
      if (f < 0 || f >= F) continue;
      const float* p = feat + ((size_t)f * C + c) * H * W;
      const int ph = bin / PW, pw = bin - ph * PW;
      const RoiGeom g = roi_geometry(roi, scale, PH, PW, sampling_ratio, aligned != 0);
      for (int iy = 0; iy < g.grid_h; ++iy) {
        const AxisTap ty = axis_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H);
        for (int ix = 0; ix < g.grid_w; ++ix) {
          const AxisTap tx = axis_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W);
          if (!(ty.ok && tx.ok)) continue;
          const float w1 = fmul(ty.wlo, tx.wlo), w2 = fmul(ty.wlo, tx.whi);
          const float w3 = fmul(ty.whi, tx.wlo), w4 = fmul(ty.whi, tx.whi);
          const float v = fadd(fadd(fadd(fmul(w1, __ldg(p + ty.lo * W + tx.lo)), fmul(w2, __ldg(p + ty.lo * W + tx.hi))),
                                    fmul(w3, __ldg(p + ty.hi * W + tx.lo))),
                               fmul(w4, __ldg(p + ty.hi * W + tx.hi)));
          val = fadd(val, v);
        }
      }
      val = fdiv(val, g.count);
    }
    out[((size_t)r * C + c) * bins + bin] = val;
  }
}

// Multi-level form of the gather kernel (MultiScaleRoIAlign on one frame's RoIs): the level of every RoI picks
// the feature map, its size and its scale, so a whole FPN crop is ONE launch with no per-level index lists.
constexpr int kMaxLevels = 8;
struct LevelSet {
  const float* feat[kMaxLevels];
  int H[kMaxLevels], W[kMaxLevels];
  float scale[kMaxLevels];
};

__global__ void __launch_bounds__(256)
roi_align_fwd_gather_levels_kernel(LevelSet S, int num_levels, const float* __restrict__ rois,
                                   const int32_t* __restrict__ levels, int R, int F, int C, int PH, int PW,
                                   int sampling_ratio, int aligned, float* __restrict__ out) {
  const int bins = PH * PW;
  const long long total = (long long)R * C * bins;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int bin = (int)(idx % bins);
    const int c = (int)((idx / bins) % C);
    const int r = (int)(idx / ((long long)bins * C));
    const float* roi = rois + (size_t)r * 5;
    const int f = (int)roi[0];
    const int lv = levels[r];
    float val = 0.0f;
    if (f >= 0 && f < F && lv >= 0 && lv < num_levels) {
      const int H = S.H[lv], W = S.W[lv];
      const float* p = S.feat[lv] + ((size_t)f * C + c) * H * W;
      const int ph = bin / PW, pw = bin - ph * PW;
      const RoiGeom g = roi_geometry(roi, S.scale[lv], PH, PW, sampling_ratio, aligned != 0);
      for (int iy = 0; iy < g.grid_h; ++iy) {
        const AxisTap ty = axis_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H);
        for (int ix = 0; ix < g.grid_w; ++ix) {
          const AxisTap tx = axis_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W);
          if (!(ty.ok && tx.ok)) continue;
          const float w1 = fmul(ty.wlo, tx.wlo), w2 = fmul(ty.wlo, tx.whi);
          const float w3 = fmul(ty.whi, tx.wlo), w4 = fmul(ty.whi, tx.whi);
          const float v = fadd(fadd(fadd(fmul(w1, __ldg(p + ty.lo * W + tx.lo)), fmul(w2, __ldg(p + ty.lo * W + tx.hi))),
                                    fmul(w3, __ldg(p + ty.hi * W + tx.lo))),
                               fmul(w4, __ldg(p + ty.hi * W + tx.hi)));
          val = fadd(val, v);
        }
      }
      val = fdiv(val, g.count);
    }
    out[idx] = val;
  }
}

// ------------------------------------------------------------------------------------------
// Backward.  Shared accumulators acc[(y - y0) * W + x][33], thread = channel.
constexpr int kBwdCh = 32;
constexpr int kBwdPad = 33;

__global__ void __launch_bounds__(256, 1)
roi_align_bwd_band_kernel(const float* __restrict__ grad_out, RoiList L, int C, int H, int W, int PH, int PW,
                          float scale, int sampling_ratio, int aligned, int band_rows, int accumulate,
                          float* __restrict__ grad_feat) {
  extern __shared__ __align__(16) float acc[];
  const int f = blockIdx.z;
  const int c0 = blockIdx.y * kBwdCh;
  const int y0 = blockIdx.x * band_rows;
  const int y1 = min(H, y0 + band_rows);
  const int rows = y1 - y0;
  const int tid = threadIdx.x;
  const int n_acc = rows * W * kBwdPad;
  for (int i = tid; i < n_acc; i += blockDim.x) acc[i] = 0.0f;
  __syncthreads();

  // 256 threads = 8 sub-bands x 32 channels; each warp owns a contiguous slice of the band's
  // rows so no two threads ever touch the same accumulator.
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int nwarp = blockDim.x >> 5;
  const int per = (rows + nwarp - 1) / nwarp;
  const int wy0 = y0 + warp * per;
  const int wy1 = min(y1, wy0 + per);
  const int c = c0 + lane;
  const bool ch_ok = c < C;
  const int bins = PH * PW;

  int first = 0, n_roi = L.n;
  if (L.seg_count) {
    first = f * L.seg_stride;
    n_roi = L.seg_count[f];
  }
  if (wy0 < wy1) {
    for (int ri = 0; ri < n_roi; ++ri) {
      const int e = first + ri;
      const int r = L.ids ? L.ids[e] : e;
      const float* roi = L.rois + (size_t)r * 5;
      if (!L.seg_count && (int)__ldg(roi) != f) continue;
      const float rr[5] = {__ldg(roi), __ldg(roi + 1), __ldg(roi + 2), __ldg(roi + 3), __ldg(roi + 4)};
      const RoiGeom g = roi_geometry(rr, scale, PH, PW, sampling_ratio, aligned != 0);
      // quick reject: rows touched by this RoI
      {
        const float ya = g.start_h, yb = fadd(g.start_h, fmul((float)PH, g.bin_h));
        if (yb < (float)wy0 - 1.0f || ya > (float)wy1) continue;
      }
      const float* go = grad_out + ((size_t)r * C + c) * bins;
      for (int ph = 0; ph < PH; ++ph) {
        for (int iy = 0; iy < g.grid_h; ++iy) {
          const AxisTap ty = axis_tap(g.start_h, g.bin_h, ph, iy, g.grid_h, H);
          if (!ty.ok) continue;
          const bool lo_in = ty.lo >= wy0 && ty.lo < wy1;
          const bool hi_in = ty.hi >= wy0 && ty.hi < wy1;
          if (!(lo_in || hi_in)) continue;
          float* row_lo = acc + (size_t)(ty.lo - y0) * W * kBwdPad + lane;
          float* row_hi = acc + (size_t)(ty.hi - y0) * W * kBwdPad + lane;
          for (int pw = 0; pw < PW; ++pw) {
            const float gv = ch_ok ? fdiv(__ldg(go + ph * PW + pw), g.count) : 0.0f;
            for (int ix = 0; ix < g.grid_w; ++ix) {
              const AxisTap tx = axis_tap(g.start_w, g.bin_w, pw, ix, g.grid_w, W);
              if (!tx.ok) continue;
              if (lo_in) {
                row_lo[tx.lo * kBwdPad] += gv * ty.wlo * tx.wlo;
                row_lo[tx.hi * kBwdPad] += gv * ty.wlo * tx.whi;
              }
              if (hi_in) {
                row_hi[tx.lo * kBwdPad] += gv * ty.whi * tx.wlo;
                row_hi[tx.hi * kBwdPad] += gv * ty.whi * tx.whi;
              }
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // transposed write-out: lanes sweep x so global stores are coalesced; smem reads have stride 33.
  const int n_px = rows * W;
  for (int cc = warp; cc < kBwdCh; cc += nwarp) {
    const int ch = c0 + cc;
    if (ch >= C) break;
    float* dst = grad_feat + ((size_t)f * C + ch) * H * W + (size_t)y0 * W;
    for (int p = lane; p < n_px; p += 32) {
      const float v = acc[(size_t)p * kBwdPad + cc];
      dst[p] = accumulate ? dst[p] + v : v;
    }
  }
}

__global__ void __launch_bounds__(256) fpn_level_kernel(int n, const float* __restrict__ boxes, int k_min, int k_max,
                                                        float s0, int lvl0, float eps, int32_t* __restrict__ levels) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* b = boxes + (size_t)i * 4;
  const float area = fmul(fsub(b[2], b[0]), fsub(b[3], b[1]));   // torchvision box_area
  const float s = __fsqrt_rn(area);
  float lvl = floorf(fadd(fadd((float)lvl0, log2f(fdiv(s, s0))), eps));   // torchpoolers.py:48
  lvl = fminf(fmaxf(lvl, (float)k_min), (float)k_max);
  if (lvl != lvl) lvl = (float)k_min;  // NaN area: torch's int cast is undefined; pin to k_min
  levels[i] = (int)lvl - k_min;
}

static size_t fwd_smem_budget() { return 200 * 1024; }

template <int CPB>
static int launch_fwd_planes(int F, int C, int H, int W, const float* feat, const RoiList& L, int PH, int PW,
                             float scale, int sr, int aligned, int use_bulk, float* out, cudaStream_t st) {
  const size_t smem = (size_t)CPB * H * W * sizeof(float);
  B2D_CUDA(cudaFuncSetAttribute(roi_align_fwd_planes_kernel<CPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  dim3 grid(ceil_div(C, CPB), F);
  roi_align_fwd_planes_kernel<CPB><<<grid, 512, smem, st>>>(feat, L, C, H, W, PH, PW, scale, sr, aligned, use_bulk, out);
  B2D_LAUNCHED();
  return B2D_OK;
}

}  // namespace b2d

using namespace b2d;

extern "C" size_t b2d_roi_align_workspace_bytes(int F, int /*C*/, int H, int /*W*/, int num_rois, int per_frame) {
  if (F <= 0 || H <= 0 || num_rois <= 0) return 0;
  if (per_frame <= 0 || per_frame > num_rois) per_frame = num_rois;
  const size_t b = rows_workspace_bytes(F, H, per_frame), c = bwd_rows_workspace_bytes(num_rois);
  return b > c ? b : c;
}

// Rows of RoIs whose batch index is outside [0, F) are written by no CTA of the frame-parallel kernels:
// they come back as zeros (plain [R,5] lists only; padded per-frame segments carry no batch index).
__global__ void __launch_bounds__(256) zero_foreign_rows_kernel(RoiList L, int F, int per_roi, float* __restrict__ out) {
  const int e = blockIdx.x;
  const int r = L.ids ? L.ids[e] : e;
  const float b = L.rois[(size_t)r * 5];
  if (b >= 0.0f && (int)b < F) return;
  float* o = out + (size_t)r * per_roi;
  for (int i = threadIdx.x; i < per_roi; i += blockDim.x) o[i] = 0.0f;
}

static int roi_align_forward_impl(int F, int C, int H, int W, const float* feat, const float* rois, int num_rois,
                                  const int32_t* roi_ids, int n_roi_ids, const int32_t* seg_count, int seg_stride,
                                  int PH, int PW, float spatial_scale, int sampling_ratio, int aligned, int route,
                                  float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (F <= 0 || C <= 0 || H <= 0 || W <= 0 || PH <= 0 || PW <= 0 || !feat || !out) return B2D_ERR_INVALID_ARG;
  if (route < B2D_ROI_ROUTE_AUTO || route > B2D_ROI_ROUTE_GATHER) return B2D_ERR_INVALID_ARG;
  RoiList L{rois, roi_ids, roi_ids ? n_roi_ids : num_rois, seg_count, seg_stride};
  if (seg_count && (seg_stride <= 0 || (long long)seg_stride * F > L.n)) return B2D_ERR_INVALID_ARG;
  if (L.n <= 0) return B2D_OK;
  if (!rois) return B2D_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  const long long total = (long long)L.n * C * PH * PW;
  if (!seg_count) {
    zero_foreign_rows_kernel<<<L.n, 256, 0, st>>>(L, F, C * PH * PW, out);
    B2D_LAUNCHED();
  }
  // Few RoIs on few channel groups (one FPN level of one frame: ~75 RoIs x 256 channels): the streaming
  // kernels would run 8-32 CTAs and pull the whole level through shared memory for a handful of RoIs
  // (p3 160x240: 325 us); one thread per output straight from L2 takes 20-45 us.  The gather kernel costs
  // ~30 ps per output, the streaming kernels ~2.5 ps at full occupancy.
  const bool sparse = route == B2D_ROI_ROUTE_GATHER ||
                      (route == B2D_ROI_ROUTE_AUTO && total <= (1LL << 22) && (long long)ceil_div(C, 32) * F < kNumSMs / 2);
  // production path: the 7x7 "rows" kernel (roi_align_rows.cu); needs the workspace
  if (C >= 16 && !sparse && route != B2D_ROI_ROUTE_PLANES) {
    const int rc = roi_align_forward_rows(F, C, H, W, feat, L, PH, PW, spatial_scale, sampling_ratio, aligned,
                                          route == B2D_ROI_ROUTE_ROWS_COOP, out, workspace, workspace_bytes, st);
    if (rc != B2D_ERR_UNSUPPORTED) return rc;
  }
  const size_t plane = (size_t)H * W * sizeof(float);
  int cpb = (int)(fwd_smem_budget() / plane);
  if (cpb > 8) cpb = 8;
  if (cpb > C) cpb = C;
  if (cpb >= 1 && !sparse) {
    const int use_bulk = ((H * W) % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat) & 15u) == 0);
    switch (cpb) {
      case 1: return launch_fwd_planes<1>(F, C, H, W, feat, L, PH, PW, spatial_scale, sampling_ratio, aligned, use_bulk, out, st);
      case 2: return launch_fwd_planes<2>(F, C, H, W, feat, L, PH, PW, spatial_scale, sampling_ratio, aligned, use_bulk, out, st);
      case 3: return launch_fwd_planes<3>(F, C, H, W, feat, L, PH, PW, spatial_scale, sampling_ratio, aligned, use_bulk, out, st);
      case 4: return launch_fwd_planes<4>(F, C, H, W, feat, L, PH, PW, spatial_scale, sampling_ratio, aligned, use_bulk, out, st);
      case 5: return launch_fwd_planes<5>(F, C, H, W, feat, L, PH, PW, spatial_scale, sampling_ratio, aligned, use_bulk, out, st);
      case 6: return launch_fwd_planes<6>(F, C, H, W, feat, L, PH, PW, spatial_scale, sampling_ratio, aligned, use_bulk, out, st);
      case 7: return launch_fwd_planes<7>(F, C, H, W, feat, L, PH, PW, spatial_scale, sampling_ratio, aligned, use_bulk, out, st);
      default: return launch_fwd_planes<8>(F, C, H, W, feat, L, PH, PW, spatial_scale, sampling_ratio, aligned, use_bulk, out, st);
    }
  }
  long long blocks = (total + 255) / 256;
  if (blocks > 32LL * kNumSMs) blocks = 32LL * kNumSMs;
  roi_align_fwd_gather_kernel<<<(int)blocks, 256, 0, st>>>(feat, L, F, C, H, W, PH, PW, spatial_scale, sampling_ratio,
                                                           aligned, out);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_roi_align_forward(int F, int C, int H, int W, const float* feat, const float* rois, int num_rois,
                                     const int32_t* roi_ids, int n_roi_ids, const int32_t* seg_count, int seg_stride,
                                     int PH, int PW, float spatial_scale, int sampling_ratio, int aligned, float* out,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  return roi_align_forward_impl(F, C, H, W, feat, rois, num_rois, roi_ids, n_roi_ids, seg_count, seg_stride, PH, PW,
                                spatial_scale, sampling_ratio, aligned, B2D_ROI_ROUTE_AUTO, out, workspace, workspace_bytes,
                                stream);
}

extern "C" int b2d_roi_align_forward_route(int F, int C, int H, int W, const float* feat, const float* rois, int num_rois,
                                           const int32_t* roi_ids, int n_roi_ids, const int32_t* seg_count,
                                           int seg_stride, int PH, int PW, float spatial_scale, int sampling_ratio,
                                           int aligned, int route, float* out, void* workspace, size_t workspace_bytes,
                                           void* stream) {
  return roi_align_forward_impl(F, C, H, W, feat, rois, num_rois, roi_ids, n_roi_ids, seg_count, seg_stride, PH, PW,
                                spatial_scale, sampling_ratio, aligned, route, out, workspace, workspace_bytes, stream);
}

extern "C" int b2d_roi_align_forward_levels(int num_levels, int F, int C, const float* const* feats, const int32_t* heights,
                                            const int32_t* widths, const float* scales, const float* rois,
                                            const int32_t* levels, int num_rois, int PH, int PW, int sampling_ratio,
                                            int aligned, float* out, void* stream) {
  if (num_levels <= 0 || num_levels > kMaxLevels || F <= 0 || C <= 0 || PH <= 0 || PW <= 0 || !feats || !heights ||
      !widths || !scales || !out)
    return B2D_ERR_INVALID_ARG;
  if (num_rois <= 0) return B2D_OK;
  if (!rois || !levels) return B2D_ERR_INVALID_ARG;
  LevelSet S{};
  for (int i = 0; i < num_levels; ++i) {
    if (!feats[i] || heights[i] <= 0 || widths[i] <= 0) return B2D_ERR_INVALID_ARG;
    S.feat[i] = feats[i];
    S.H[i] = heights[i];
    S.W[i] = widths[i];
    S.scale[i] = scales[i];
  }
  const long long total = (long long)num_rois * C * PH * PW;
  long long blocks = (total + 255) / 256;
  if (blocks > 32LL * kNumSMs) blocks = 32LL * kNumSMs;
  roi_align_fwd_gather_levels_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(S, num_levels, rois, levels, num_rois, F, C,
                                                                              PH, PW, sampling_ratio, aligned, out);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_roi_align_backward(int F, int C, int H, int W, const float* grad_out, const float* rois,
                                      int num_rois, const int32_t* roi_ids, int n_roi_ids, const int32_t* seg_count,
                                      int seg_stride, int PH, int PW, float spatial_scale, int sampling_ratio,
                                      int aligned, int accumulate, float* grad_feat, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  if (F <= 0 || C <= 0 || H <= 0 || W <= 0 || PH <= 0 || PW <= 0 || !grad_feat) return B2D_ERR_INVALID_ARG;
  RoiList L{rois, roi_ids, roi_ids ? n_roi_ids : num_rois, seg_count, seg_stride};
  if (seg_count && (seg_stride <= 0 || (long long)seg_stride * F > L.n)) return B2D_ERR_INVALID_ARG;
  if (L.n > 0 && (!rois || !grad_out)) return B2D_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  {
    // production path: table-driven one-warp-per-row kernel (roi_align_bwd_rows.cu); needs the workspace
    const int rc = roi_align_backward_rows(F, C, H, W, grad_out, L, PH, PW, spatial_scale, sampling_ratio, aligned,
                                           accumulate, grad_feat, workspace, workspace_bytes, st);
    if (rc != B2D_ERR_UNSUPPORTED) return rc;
  }
  const size_t row_bytes = (size_t)W * kBwdPad * sizeof(float);
  int band = (int)((200 * 1024) / row_bytes);
  if (band < 1) return B2D_ERR_UNSUPPORTED;   // W > ~1500: needs column tiling
  if (band > H) band = H;
  // balance the bands
  const int n_band = ceil_div(H, band);
  band = ceil_div(H, n_band);
  const size_t smem = (size_t)band * row_bytes;
  B2D_CUDA(cudaFuncSetAttribute(roi_align_bwd_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(n_band, ceil_div(C, kBwdCh), F);
  roi_align_bwd_band_kernel<<<grid, 256, smem, st>>>(grad_out, L, C, H, W, PH, PW, spatial_scale, sampling_ratio, aligned,
                                                     band, accumulate, grad_feat);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_fpn_level_map(int n, const float* boxes, int k_min, int k_max, float canonical_scale,
                                 int canonical_level, float eps, int32_t* levels, void* stream) {
  if (n < 0 || (n > 0 && (!boxes || !levels))) return B2D_ERR_INVALID_ARG;
  if (n == 0) return B2D_OK;
  fpn_level_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(n, boxes, k_min, k_max, canonical_scale,
                                                                    canonical_level, eps, levels);
  B2D_LAUNCHED();
  return B2D_OK;
}
