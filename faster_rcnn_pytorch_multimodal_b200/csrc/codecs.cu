// Box codecs, IoU, anchors and the rotated-box AABB.  All elementwise / small; the point of
// these kernels is to replace the reference's ~20 ATen launches per call with one launch whose
// arithmetic is rounded exactly like the ATen op chain (no FMA contraction).
//
// Reference: model/bbox_transform.py:16-49,52-70,75-105,132-233,235-257; utils/bbox.py:5-33,
// 296-336; layer_utils/snippets.py:13-40; layer_utils/generate_3d_anchors.py:47-118.
#include "common.cuh"

namespace b2d {

// utils/bbox.py:22-32
__device__ __forceinline__ float iou_plus1(const float bx1, const float by1, const float bx2, const float by2,
                                           const float b_area, const float qx1, const float qy1, const float qx2,
                                           const float qy2, const float q_area) {
  const float iw = fmaxf(fadd(fsub(fminf(bx2, qx2), fmaxf(bx1, qx1)), 1.0f), 0.0f);
  const float ih = fmaxf(fadd(fsub(fminf(by2, qy2), fmaxf(by1, qy1)), 1.0f), 0.0f);
  const float inter = fmul(iw, ih);
  const float ua = fsub(fadd(b_area, q_area), inter);
  return fdiv(inter, ua);
}
__device__ __forceinline__ float area_plus1(float x1, float y1, float x2, float y2) {
  return fmul(fadd(fsub(x2, x1), 1.0f), fadd(fsub(y2, y1), 1.0f));
}

__global__ void __launch_bounds__(256) bbox_overlaps_kernel(int n, int k, const float* __restrict__ boxes, int bs,
                                                            const float* __restrict__ query, int qs,
                                                            float* __restrict__ out) {
  const long long total = (long long)n * k;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / k), j = (int)(idx - (long long)i * k);
    const float* b = boxes + (size_t)i * bs;
    const float* q = query + (size_t)j * qs;
    const float b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2), b3 = __ldg(b + 3);
    const float q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
    out[idx] = iou_plus1(b0, b1, b2, b3, area_plus1(b0, b1, b2, b3), q0, q1, q2, q3, area_plus1(q0, q1, q2, q3));
  }
}

// model/bbox_transform.py:52-70
__device__ __forceinline__ void encode4(const float* e, const float* g, float* t) {
  const float ew = fadd(fsub(e[2], e[0]), 1.0f), eh = fadd(fsub(e[3], e[1]), 1.0f);
  const float diag = __fsqrt_rn(fadd(fmul(ew, ew), fmul(eh, eh)));
  const float ecx = fadd(e[0], fmul(0.5f, ew)), ecy = fadd(e[1], fmul(0.5f, eh));
  const float gw = fadd(fsub(g[2], g[0]), 1.0f), gh = fadd(fsub(g[3], g[1]), 1.0f);
  const float gcx = fadd(g[0], fmul(0.5f, gw)), gcy = fadd(g[1], fmul(0.5f, gh));
  t[0] = fdiv(fsub(gcx, ecx), diag);
  t[1] = fdiv(fsub(gcy, ecy), diag);
  t[2] = logf(fdiv(gw, ew));
  t[3] = logf(fdiv(gh, eh));
}

__global__ void __launch_bounds__(256) bbox_transform_kernel(int n, const float* __restrict__ ex, int es,
                                                             const float* __restrict__ gt, int gs,
                                                             float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float e[4], g[4], t[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    e[c] = ex[(size_t)i * es + c];
    g[c] = gt[(size_t)i * gs + c];
  }
  encode4(e, g, t);
#pragma unroll
  for (int c = 0; c < 4; ++c) out[(size_t)i * 4 + c] = t[c];
}

// model/bbox_transform.py:75-105 (+ optional clip :252-255)
__global__ void __launch_bounds__(256) bbox_transform_inv_kernel(int n, int k, const float* __restrict__ boxes, int bs,
                                                                 const float* __restrict__ deltas, int use_scale,
                                                                 float scale, int clip, const float* __restrict__ info,
                                                                 float* __restrict__ out) {
  const long long total = (long long)n * k;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int i = (int)(idx / k);
  float b[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    b[c] = boxes[(size_t)i * bs + c];
    if (use_scale) b[c] = fdiv(b[c], scale);
  }
  const float* d = deltas + idx * 4;
  const float w = fadd(fsub(b[2], b[0]), 1.0f), h = fadd(fsub(b[3], b[1]), 1.0f);
  const float diag = __fsqrt_rn(fadd(fmul(w, w), fmul(h, h)));
  const float cx = fadd(b[0], fmul(0.5f, w)), cy = fadd(b[1], fmul(0.5f, h));
  const float pcx = fadd(fmul(d[0], diag), cx), pcy = fadd(fmul(d[1], diag), cy);
  const float pw = fmul(expf(d[2]), w), ph = fmul(expf(d[3]), h);
  float o[4] = {fsub(pcx, fmul(0.5f, pw)), fsub(pcy, fmul(0.5f, ph)), fadd(pcx, fmul(0.5f, pw)),
                fadd(pcy, fmul(0.5f, ph))};
  if (clip) {
    const float xl = info[0], xh = fsub(info[1], 1.0f), yl = info[2], yh = fsub(info[3], 1.0f);
    o[0] = clampf(o[0], xl, xh);
    o[1] = clampf(o[1], yl, yh);
    o[2] = clampf(o[2], xl, xh);
    o[3] = clampf(o[3], yl, yh);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) out[idx * 4 + c] = o[c];
}

__global__ void __launch_bounds__(256) clip_boxes_kernel(long long total, const float* __restrict__ boxes,
                                                         const float* __restrict__ info, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx & 3);
  const float lo = (c & 1) ? info[2] : info[0];
  const float hi = fsub((c & 1) ? info[3] : info[1], 1.0f);
  out[idx] = clampf(boxes[idx], lo, hi);
}

// model/bbox_transform.py:16-49
__device__ __forceinline__ void lidar_encode(const float* roi, const float* anc, const float* gt, float* t) {
  const float rl = fadd(fsub(roi[2], roi[0]), 1.0f), rw = fadd(fsub(roi[3], roi[1]), 1.0f);
  const float eh = anc[5];
  const float cx = fadd(roi[0], fdiv(rl, 2.0f)), cy = fadd(roi[1], fdiv(rw, 2.0f));
  const float diag = __fsqrt_rn(fadd(fmul(rl, rl), fmul(rw, rw)));
  t[0] = fdiv(fsub(gt[0], cx), diag);
  t[1] = fdiv(fsub(gt[1], cy), diag);
  t[2] = fdiv(fsub(gt[2], anc[2]), eh);
  t[3] = logf(fdiv(gt[3], rl));
  t[4] = logf(fdiv(gt[4], rw));
  t[5] = logf(fdiv(gt[5], eh));
  t[6] = gt[6];
}

__global__ void __launch_bounds__(256) lidar_transform_kernel(int n, const float* __restrict__ rois, int rs,
                                                              const float* __restrict__ anchors,
                                                              const float* __restrict__ gt, int gs,
                                                              float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r[4], a[7], g[7], t[7];
#pragma unroll
  for (int c = 0; c < 4; ++c) r[c] = rois[(size_t)i * rs + c];
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    a[c] = anchors[(size_t)i * 7 + c];
    g[c] = gt[(size_t)i * gs + c];
  }
  lidar_encode(r, a, g, t);
#pragma unroll
  for (int c = 0; c < 7; ++c) out[(size_t)i * 7 + c] = t[c];
}

// model/bbox_transform.py:174-233 (mode 0) and :132-169 (mode 1)
__global__ void __launch_bounds__(256) lidar_transform_inv_kernel(int n, int k, const float* __restrict__ rois, int rs,
                                                                  const float* __restrict__ boxes,
                                                                  const float* __restrict__ deltas, int mode,
                                                                  float* __restrict__ out) {
  const long long total = (long long)n * k;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int i = (int)(idx / k);
  const float* r = rois + (size_t)i * rs;
  const float* b = boxes + (size_t)i * 7;
  const float* d = deltas + idx * 7;
  float* o = out + idx * 7;
  const float rl = fadd(fsub(r[2], r[0]), 1.0f), rw = fadd(fsub(r[3], r[1]), 1.0f);
  const float hh = b[5];
  if (mode == 0) {
    const float cx = fadd(r[0], fdiv(rl, 2.0f)), cy = fadd(r[1], fdiv(rw, 2.0f)), cz = b[2];
    const float diag = __fsqrt_rn(fadd(fmul(rl, rl), fmul(rw, rw)));
    o[0] = fadd(fmul(d[0], diag), cx);
    o[1] = fadd(fmul(d[1], diag), cy);
    o[2] = fadd(fmul(d[2], hh), cz);
    o[3] = fmul(expf(d[3]), rl);
    o[4] = fmul(expf(d[4]), rw);
    o[5] = fmul(expf(d[5]), hh);
    o[6] = d[6];
  } else {
    const float v[7] = {fmul(d[0], rl), fmul(d[1], rw), fmul(d[2], hh), fsub(expf(d[3]), 1.0f),
                        fsub(expf(d[4]), 1.0f), fsub(expf(d[5]), 1.0f), d[6]};
#pragma unroll
    for (int c = 0; c < 7; ++c) o[c] = fmul(v[c], v[c]);
  }
}

// utils/bbox.py:296-336 (fp32 torch flavour)
__global__ void __launch_bounds__(256) bbaa_kernel(int n, const float* __restrict__ boxes, int clip, float width,
                                                   float height, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* b = boxes + (size_t)i * 7;
  const float c = cosf(b[6]), s = sinf(b[6]);
  const float hl = fdiv(b[3], 2.0f), hw = fdiv(b[4], 2.0f);
  // row j of M = [[c, s], [-s, c]] times (+-hl, +-hw)
  const float m[2][2] = {{c, s}, {-s, c}};
  float lo[2], hi[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float a0 = fmul(m[j][0], -hl), b0 = fmul(m[j][0], hl);
    const float a1 = fmul(m[j][1], -hw), b1 = fmul(m[j][1], hw);
    lo[j] = fadd(fadd(fminf(a0, b0), fminf(a1, b1)), b[j]);
    hi[j] = fadd(fadd(fmaxf(a0, b0), fmaxf(a1, b1)), b[j]);
  }
  float o[4] = {lo[0], lo[1], hi[0], hi[1]};
  if (clip) {
    o[0] = clampf(o[0], 0.0f, fsub(width, 1.0f));
    o[2] = clampf(o[2], 0.0f, fsub(width, 1.0f));
    o[1] = clampf(o[1], 0.0f, fsub(height, 1.0f));
    o[3] = clampf(o[3], 0.0f, fsub(height, 1.0f));
  }
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) out[(size_t)i * 4 + cc] = o[cc];
}

// layer_utils/snippets.py:25-37: anchors[(h*W + w)*A + a] = fp32(base[a] + shift), fp64 add
constexpr int kMaxBase = 64;
struct BaseAnchors {
  double v[kMaxBase * 4];
};
__global__ void __launch_bounds__(256) anchors_kernel(int H, int W, int stride, int A, BaseAnchors base,
                                                      float* __restrict__ out) {
  const long long total = (long long)H * W * A * 4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx & 3);
    const long long n = idx >> 2;
    const int a = (int)(n % A);
    const long long loc = n / A;
    const int w = (int)(loc % W), h = (int)(loc / W);
    const double shift = (c & 1) ? (double)h * stride : (double)w * stride;
    out[idx] = (float)(base.v[a * 4 + c] + shift);
  }
}

}  // namespace b2d

using namespace b2d;

static int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = 64LL * kNumSMs;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

extern "C" int b2d_bbox_overlaps(int n, int k, const float* boxes, int box_stride, const float* query,
                                 int query_stride, float* overlaps, void* stream) {
  if (n < 0 || k < 0) return B2D_ERR_INVALID_ARG;
  if (n == 0 || k == 0) return B2D_OK;
  if (!boxes || !query || !overlaps || box_stride < 4 || query_stride < 4) return B2D_ERR_INVALID_ARG;
  bbox_overlaps_kernel<<<grid_for((long long)n * k), 256, 0, as_stream(stream)>>>(n, k, boxes, box_stride, query,
                                                                                 query_stride, overlaps);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_bbox_transform(int n, const float* ex, int es, const float* gt, int gs, float* targets,
                                  void* stream) {
  if (n < 0) return B2D_ERR_INVALID_ARG;
  if (n == 0) return B2D_OK;
  if (!ex || !gt || !targets || es < 4 || gs < 4) return B2D_ERR_INVALID_ARG;
  bbox_transform_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(n, ex, es, gt, gs, targets);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_bbox_transform_inv(int n, int k, const float* boxes, int box_stride, const float* deltas,
                                      int use_scale, float scale, int clip, const float* info, float* out,
                                      void* stream) {
  if (n < 0 || k < 0) return B2D_ERR_INVALID_ARG;
  if (n == 0 || k == 0) return B2D_OK;
  if (!boxes || !deltas || !out || box_stride < 4 || (clip && !info)) return B2D_ERR_INVALID_ARG;
  const long long total = (long long)n * k;
  bbox_transform_inv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
      n, k, boxes, box_stride, deltas, use_scale, scale, clip, info, out);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_clip_boxes(int n, int k, const float* boxes, const float* info, float* out, void* stream) {
  if (n < 0 || k < 0) return B2D_ERR_INVALID_ARG;
  if (n == 0 || k == 0) return B2D_OK;
  if (!boxes || !info || !out) return B2D_ERR_INVALID_ARG;
  const long long total = (long long)n * k * 4;
  clip_boxes_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(total, boxes, info, out);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_lidar_bbox_transform(int n, const float* ex_rois, int roi_stride, const float* ex_anchors,
                                        const float* gt_rois, int gt_stride, float* targets, void* stream) {
  if (n < 0) return B2D_ERR_INVALID_ARG;
  if (n == 0) return B2D_OK;
  if (!ex_rois || !ex_anchors || !gt_rois || !targets || roi_stride < 4 || gt_stride < 7) return B2D_ERR_INVALID_ARG;
  lidar_transform_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(n, ex_rois, roi_stride, ex_anchors, gt_rois,
                                                                          gt_stride, targets);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_lidar_bbox_transform_inv(int n, int k, const float* rois, int roi_stride, const float* boxes,
                                            const float* deltas, int mode, float* out, void* stream) {
  if (n < 0 || k < 0 || mode < 0 || mode > 1) return B2D_ERR_INVALID_ARG;
  if (n == 0 || k == 0) return B2D_OK;
  if (!rois || !boxes || !deltas || !out || roi_stride < 4) return B2D_ERR_INVALID_ARG;
  const long long total = (long long)n * k;
  lidar_transform_inv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(n, k, rois, roi_stride,
                                                                                            boxes, deltas, mode, out);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_bbaa(int n, const float* boxes7, int clip, float width, float height, float* aabb, void* stream) {
  if (n < 0) return B2D_ERR_INVALID_ARG;
  if (n == 0) return B2D_OK;
  if (!boxes7 || !aabb) return B2D_ERR_INVALID_ARG;
  bbaa_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(n, boxes7, clip, width, height, aabb);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_generate_anchors(int height, int width, int feat_stride, int num_base, const double* base_host,
                                    float* anchors, void* stream) {
  if (height <= 0 || width <= 0 || num_base <= 0 || !base_host || !anchors) return B2D_ERR_INVALID_ARG;
  if (num_base > kMaxBase) return B2D_ERR_UNSUPPORTED;
  BaseAnchors b;
  for (int i = 0; i < num_base * 4; ++i) b.v[i] = base_host[i];
  const long long total = (long long)height * width * num_base * 4;
  anchors_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(height, width, feat_stride, num_base, b, anchors);
  B2D_LAUNCHED();
  return B2D_OK;
}
