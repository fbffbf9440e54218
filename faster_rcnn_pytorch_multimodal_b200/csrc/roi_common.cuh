// Shared RoIAlign geometry (torchvision csrc/ops/cpu/roi_align_kernel.cpp restated per axis).
#pragma once
#include "common.cuh"

namespace b2d {

struct RoiGeom {
  float start_w, start_h, bin_w, bin_h;
  int grid_w, grid_h;
  float count;
};

__device__ __forceinline__ RoiGeom roi_geometry(const float* __restrict__ roi, float scale, int PH, int PW,
                                                int sampling_ratio, bool aligned) {
  RoiGeom g;
  const float off = aligned ? 0.5f : 0.0f;
  g.start_w = fsub(fmul(roi[1], scale), off);
  g.start_h = fsub(fmul(roi[2], scale), off);
  const float end_w = fsub(fmul(roi[3], scale), off);
  const float end_h = fsub(fmul(roi[4], scale), off);
  float rw = fsub(end_w, g.start_w);
  float rh = fsub(end_h, g.start_h);
  if (!aligned) {
    rw = fmaxf(rw, 1.0f);
    rh = fmaxf(rh, 1.0f);
  }
  g.bin_h = fdiv(rh, (float)PH);
  g.bin_w = fdiv(rw, (float)PW);
  g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(fdiv(rh, (float)PH));
  g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(fdiv(rw, (float)PW));
  g.count = (float)max(g.grid_h * g.grid_w, 1);
  return g;
}

struct AxisTap {
  int lo, hi;     // pixel indices
  float wlo, whi; // weights of lo / hi (hy, ly in torchvision's naming)
  bool ok;
};

__device__ __forceinline__ AxisTap axis_tap(float start, float bin, int p, int i, int grid, int limit) {
  AxisTap t;
  float c = fadd(fadd(start, fmul((float)p, bin)), fdiv(fmul(fadd((float)i, 0.5f), bin), (float)grid));
  t.ok = !(c < -1.0f || c > (float)limit);
  if (c <= 0.0f) c = 0.0f;
  int lo = (int)c;
  int hi;
  if (lo >= limit - 1) {
    hi = lo = limit - 1;
    c = (float)lo;
  } else {
    hi = lo + 1;
  }
  const float l = fsub(c, (float)lo);
  t.lo = lo;
  t.hi = hi;
  t.whi = l;
  t.wlo = fsub(1.0f, l);
  if (!t.ok) {
    t.lo = t.hi = 0;
    t.wlo = t.whi = 0.0f;
  }
  return t;
}

struct RoiList {
  const float* rois;        // [R,5]
  const int32_t* ids;       // optional indirection
  int n;                    // list length (R or n_ids)
  const int32_t* seg_count; // optional [F]: frame f owns list entries [f*seg_stride, +seg_count[f])
  int seg_stride;
};


}  // namespace b2d
