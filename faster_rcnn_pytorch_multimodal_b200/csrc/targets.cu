// Training-target assignment: anchor_target_layer_torch and proposal_target_layer.
//
// Reference: layer_utils/anchor_target_layer.py:22-165 (+ _unmap :335, _compute_targets :361),
// layer_utils/proposal_target_layer.py:22-262, utils/bbox.py:5-33, model/bbox_transform.py:16-70.
//
// The reference materialises the dense [N_inside, G] IoU matrix (38 MB at Waymo sizes) and
// compacts with ~12 boolean-mask indexings, each a host sync.  Here the IoU matrix never
// exists: pass 1 keeps per-anchor max/argmax and per-GT max, pass 2 re-evaluates the G IoUs
// of an anchor to test "equals some GT's max" (bit-identical arithmetic, so equality is exact).
// Both layers are split at their randperm()/randint() draws: phase 1 ends with ordered fg/bg
// index lists + counts, the host draws the permutations from torch's generator exactly as the
// reference does (sizes are data dependent), phase 2 consumes them.
#include "common.cuh"

namespace b2d {

constexpr int kMaxGt = 256;   // GT boxes staged in shared memory per frame

__device__ __forceinline__ float area_p1(const float* b) {
  return fmul(fadd(fsub(b[2], b[0]), 1.0f), fadd(fsub(b[3], b[1]), 1.0f));
}
__device__ __forceinline__ float iou_p1(const float* b, float b_area, const float* q, float q_area) {
  const float iw = fmaxf(fadd(fsub(fminf(b[2], q[2]), fmaxf(b[0], q[0])), 1.0f), 0.0f);
  const float ih = fmaxf(fadd(fsub(fminf(b[3], q[3]), fmaxf(b[1], q[1])), 1.0f), 0.0f);
  const float inter = fmul(iw, ih);
  return fdiv(inter, fsub(fadd(b_area, q_area), inter));
}

struct AnchorWs {
  float* max_ov;       // [F][N]
  int32_t* argmax;     // [F][N]
  int8_t* label;       // [F][N]   -2 = outside the frame, else -1/0/1
  uint32_t* gt_max;    // [F][kMaxGt] float bits (IoU >= 0, so unsigned order == float order)
  int32_t* fg_list;    // [F][N] flat anchor indices with label 1, ascending
  int32_t* bg_list;    // [F][N] flat anchor indices with label 0, ascending
  size_t bytes;
};

static AnchorWs carve_anchor(void* base, int F, int N) {
  AnchorWs w;
  size_t off = 0;
  char* p = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  w.max_ov = reinterpret_cast<float*>(take(sizeof(float) * (size_t)F * N));
  w.argmax = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F * N));
  w.label = reinterpret_cast<int8_t*>(take((size_t)F * N));
  w.gt_max = reinterpret_cast<uint32_t*>(take(sizeof(uint32_t) * (size_t)F * kMaxGt));
  w.fg_list = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F * N));
  w.bg_list = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F * N));
  w.bytes = off;
  return w;
}

// pass 1: inside filter (:37-42), per-anchor max/argmax over GT (:64-66), per-GT max (:67-69)
__global__ void __launch_bounds__(256) anchor_iou_kernel(int N, int max_gt, const float* __restrict__ anchors,
                                                         const float* __restrict__ gt_all,
                                                         const int32_t* __restrict__ num_gt,
                                                         const float* __restrict__ info_all, AnchorWs w) {
  __shared__ float s_gt[kMaxGt * 4];
  __shared__ float s_ga[kMaxGt];
  __shared__ unsigned int s_gmax[kMaxGt];
  const int f = blockIdx.y;
  const int G = min(num_gt[f], max_gt);
  const float* gt = gt_all + (size_t)f * max_gt * 5;
  const float* info = info_all + f * 7;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float q[4] = {gt[g * 5], gt[g * 5 + 1], gt[g * 5 + 2], gt[g * 5 + 3]};
    s_gt[g * 4] = q[0]; s_gt[g * 4 + 1] = q[1]; s_gt[g * 4 + 2] = q[2]; s_gt[g * 4 + 3] = q[3];
    s_ga[g] = area_p1(q);
    s_gmax[g] = 0u;
  }
  __syncthreads();
  const float x_lo = info[0], x_hi = info[1], y_lo = info[2], y_hi = info[3];
  const int stride = gridDim.x * blockDim.x;
  const int iters = (N + stride - 1) / stride;
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it, n += stride) {
    bool inside = false;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < N) {
      const float4 v = *reinterpret_cast<const float4*>(anchors + (size_t)n * 4);
      a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
      inside = (a[0] >= x_lo) && (a[1] >= y_lo) && (a[2] < x_hi) && (a[3] < y_hi);
    }
    const float a_area = area_p1(a);
    float best = -INFINITY;
    int arg = 0;
    for (int g = 0; g < G; ++g) {
      float ov = 0.0f;
      if (inside) {
        ov = iou_p1(a, a_area, s_gt + g * 4, s_ga[g]);
        if (ov > best) { best = ov; arg = g; }   // first maximum wins, like torch.argmax
      }
      // per-GT max over inside anchors: warp max, then one shared atomic per warp
      unsigned int bits = inside ? __float_as_uint(fmaxf(ov, 0.0f)) : 0u;
      bits = __reduce_max_sync(0xFFFFFFFFu, bits);
      if ((threadIdx.x & 31) == 0 && bits > s_gmax[g]) atomicMax(&s_gmax[g], bits);
    }
    if (n < N) {
      w.max_ov[(size_t)f * N + n] = inside ? best : 0.0f;
      w.argmax[(size_t)f * N + n] = arg;
      w.label[(size_t)f * N + n] = inside ? (int8_t)-1 : (int8_t)-2;
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x)
    if (s_gmax[g]) atomicMax(&w.gt_max[(size_t)f * kMaxGt + g], s_gmax[g]);
}

// pass 2: labels before subsampling (:74-89)
__global__ void __launch_bounds__(256) anchor_label_kernel(int N, int max_gt, const float* __restrict__ anchors,
                                                           const float* __restrict__ gt_all,
                                                           const int32_t* __restrict__ num_gt, float neg_ov,
                                                           float pos_ov, int clobber, AnchorWs w) {
  __shared__ float s_gt[kMaxGt * 4];
  __shared__ float s_ga[kMaxGt];
  __shared__ float s_gmax[kMaxGt];
  const int f = blockIdx.y;
  const int G = min(num_gt[f], max_gt);
  const float* gt = gt_all + (size_t)f * max_gt * 5;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float q[4] = {gt[g * 5], gt[g * 5 + 1], gt[g * 5 + 2], gt[g * 5 + 3]};
    s_gt[g * 4] = q[0]; s_gt[g * 4 + 1] = q[1]; s_gt[g * 4 + 2] = q[2]; s_gt[g * 4 + 3] = q[3];
    s_ga[g] = area_p1(q);
    // clamp(gt_max, eps, inf)  (:71)
    s_gmax[g] = fmaxf(__uint_as_float(w.gt_max[(size_t)f * kMaxGt + g]), 1.1920928955078125e-07f);
  }
  __syncthreads();
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    int8_t lab = w.label[(size_t)f * N + n];
    if (lab == -2) continue;
    const float4 v = *reinterpret_cast<const float4*>(anchors + (size_t)n * 4);
    const float a[4] = {v.x, v.y, v.z, v.w};
    const float a_area = area_p1(a);
    const float mx = w.max_ov[(size_t)f * N + n];
    bool is_gt_best = false;
    for (int g = 0; g < G; ++g) is_gt_best |= (iou_p1(a, a_area, s_gt + g * 4, s_ga[g]) == s_gmax[g]);
    lab = -1;
    if (!clobber && mx < neg_ov) lab = 0;
    if (is_gt_best) lab = 1;
    if (mx >= pos_ov) lab = 1;
    if (clobber && mx < neg_ov) lab = 0;
    w.label[(size_t)f * N + n] = lab;
  }
}

// ordered compaction of the fg / bg anchor lists, one CTA per frame; every thread owns 16 consecutive anchors
// per round (16 384 per round: 15 rounds at Waymo size instead of 235 rounds of 1024 with 4 barriers each)
__global__ void __launch_bounds__(1024) anchor_lists_kernel(int N, AnchorWs w, int32_t* __restrict__ counts) {
  constexpr int kPer = 16;
  __shared__ int s_warp[3][32];
  __shared__ int s_base[3];
  __shared__ int s_tot[3];
  const int f = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 3) s_base[threadIdx.x] = 0;
  __syncthreads();
  const int8_t* label = w.label + (size_t)f * N;
  for (int base = 0; base < N; base += 1024 * kPer) {
    const int n0 = base + threadIdx.x * kPer;
    int8_t lab[kPer];
    int c[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      lab[j] = n0 + j < N ? label[n0 + j] : (int8_t)-2;
      c[0] += lab[j] != -2;
      c[1] += lab[j] == 1;
      c[2] += lab[j] == 0;
    }
    int incl[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      incl[q] = c[q];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, incl[q], d);
        if (lane >= d) incl[q] += t;
      }
      if (lane == 31) s_warp[q][warp] = incl[q];
    }
    __syncthreads();
    if (warp < 3) {
      const int v = s_warp[warp][lane];
      int in2 = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, in2, d);
        if (lane >= d) in2 += t;
      }
      s_warp[warp][lane] = in2 - v;   // exclusive prefix of the per-warp totals
      if (lane == 31) s_tot[warp] = in2;
    }
    __syncthreads();
    int o1 = s_base[1] + s_warp[1][warp] + incl[1] - c[1];
    int o2 = s_base[2] + s_warp[2][warp] + incl[2] - c[2];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      if (lab[j] == 1) w.fg_list[(size_t)f * N + o1++] = n0 + j;
      if (lab[j] == 0) w.bg_list[(size_t)f * N + o2++] = n0 + j;
    }
    __syncthreads();
    if (threadIdx.x < 3) s_base[threadIdx.x] += s_tot[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x < 3) counts[f * 4 + threadIdx.x] = s_base[threadIdx.x];
  if (threadIdx.x == 3) counts[f * 4 + 3] = 0;
}

__global__ void __launch_bounds__(256) anchor_disable_kernel(int N, int disable_stride,
                                                             const int64_t* __restrict__ fg_disable,
                                                             const int32_t* __restrict__ n_fg_disable,
                                                             const int64_t* __restrict__ bg_disable,
                                                             const int32_t* __restrict__ n_bg_disable, AnchorWs w) {
  const int f = blockIdx.y;
  const int nf = n_fg_disable ? n_fg_disable[f] : 0;
  const int nb = n_bg_disable ? n_bg_disable[f] : 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf + nb; i += gridDim.x * blockDim.x) {
    if (i < nf) {
      const int pos = (int)fg_disable[(size_t)f * disable_stride + i];
      w.label[(size_t)f * N + w.fg_list[(size_t)f * N + pos]] = -1;       // :97-98
    } else {
      const int pos = (int)bg_disable[(size_t)f * disable_stride + (i - nf)];
      w.label[(size_t)f * N + w.bg_list[(size_t)f * N + pos]] = -1;       // :106-107
    }
  }
}

// targets (:110), weights (:113-132), unmap (:137-142) and the output layouts (:145-164)
__global__ void __launch_bounds__(256) anchor_emit_kernel(int N, int max_gt, int A, int HW,
                                                          const float* __restrict__ anchors,
                                                          const float* __restrict__ gt_all,
                                                          const int32_t* __restrict__ counts,
                                                          const int32_t* __restrict__ n_fg_disable,
                                                          const int32_t* __restrict__ n_bg_disable,
                                                          const float* __restrict__ inside_w4, float positive_weight,
                                                          AnchorWs w, float* __restrict__ labels,
                                                          float* __restrict__ targets, float* __restrict__ inside_w,
                                                          float* __restrict__ outside_w) {
  const int f = blockIdx.y;
  const float* gt = gt_all + (size_t)f * max_gt * 5;
  const int n_fg = counts[f * 4 + 1] - (n_fg_disable ? n_fg_disable[f] : 0);
  const int n_bg = counts[f * 4 + 2] - (n_bg_disable ? n_bg_disable[f] : 0);
  float pos_w, neg_w;
  if (positive_weight < 0.0f) {
    pos_w = neg_w = (float)(1.0 / (double)(n_fg + n_bg));            // :119-125
  } else {
    pos_w = fdiv(positive_weight, (float)n_fg);
    neg_w = fdiv(fsub(1.0f, positive_weight), (float)n_bg);
  }
  const float iw[4] = {inside_w4[0], inside_w4[1], inside_w4[2], inside_w4[3]};
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    const int8_t lab = w.label[(size_t)f * N + n];
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (lab != -2) {
      const float4 v = *reinterpret_cast<const float4*>(anchors + (size_t)n * 4);
      const float e[4] = {v.x, v.y, v.z, v.w};
      const float* g = gt + (size_t)w.argmax[(size_t)f * N + n] * 5;
      // bbox_transform (model/bbox_transform.py:52-70)
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
    const float lab_f = lab == -2 ? -1.0f : (float)lab;
    const float ow = lab == 1 ? pos_w : (lab == 0 ? neg_w : 0.0f);
    // labels: [F, A, H, W]; n = loc*A + a
    const int loc = n / A, a = n - loc * A;
    labels[((size_t)f * A + a) * HW + loc] = lab_f;
    float4* tp = reinterpret_cast<float4*>(targets + ((size_t)f * N + n) * 4);
    float4* ip = reinterpret_cast<float4*>(inside_w + ((size_t)f * N + n) * 4);
    float4* op = reinterpret_cast<float4*>(outside_w + ((size_t)f * N + n) * 4);
    *tp = make_float4(t[0], t[1], t[2], t[3]);
    *ip = lab == 1 ? make_float4(iw[0], iw[1], iw[2], iw[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    *op = make_float4(ow, ow, ow, ow);
  }
}

// ------------------------------------------------------------------------------------------
// proposal_target_layer
__global__ void __launch_bounds__(1024) roi_assign_kernel(int R, int G, const float* __restrict__ rois,
                                                          const float* __restrict__ gt, float fg_thresh, float bg_hi,
                                                          float bg_lo, int bg_mode, float* __restrict__ max_ov,
                                                          int32_t* __restrict__ assign, int32_t* __restrict__ fg_list,
                                                          int32_t* __restrict__ bg_list, int32_t* __restrict__ counts) {
  __shared__ int s_warp[2][32];
  __shared__ int s_base[2];
  __shared__ int s_tot[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 2) s_base[threadIdx.x] = 0;
  __syncthreads();
  for (int base = 0; base < R; base += 1024) {
    const int r = base + threadIdx.x;
    bool is_fg = false, is_bg = false;
    if (r < R) {
      const float b[4] = {rois[(size_t)r * 5 + 1], rois[(size_t)r * 5 + 2], rois[(size_t)r * 5 + 3],
                          rois[(size_t)r * 5 + 4]};
      const float ba = area_p1(b);
      float best = -INFINITY;
      int arg = 0;
      for (int g = 0; g < G; ++g) {
        const float q[4] = {__ldg(gt + g * 5), __ldg(gt + g * 5 + 1), __ldg(gt + g * 5 + 2), __ldg(gt + g * 5 + 3)};
        const float ov = iou_p1(b, ba, q, area_p1(q));
        if (ov > best) { best = ov; arg = g; }
      }
      max_ov[r] = best;
      assign[r] = arg;
      is_fg = best >= fg_thresh;                                           // :200
      is_bg = bg_mode ? (best < bg_hi && best >= bg_lo) : false;            // :203-204
    }
    const bool flag[2] = {is_fg, is_bg};
    int rank[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const unsigned bal = __ballot_sync(0xFFFFFFFFu, flag[q]);
      rank[q] = __popc(bal & ((1u << lane) - 1u));
      if (lane == 0) s_warp[q][warp] = __popc(bal);
    }
    __syncthreads();
    if (warp < 2) {
      const int v = s_warp[warp][lane];
      int incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
      }
      s_warp[warp][lane] = incl - v;
      if (lane == 31) s_tot[warp] = incl;
    }
    __syncthreads();
    if (is_fg) fg_list[s_base[0] + s_warp[0][warp] + rank[0]] = r;
    if (is_bg) bg_list[s_base[1] + s_warp[1][warp] + rank[1]] = r;
    __syncthreads();
    if (threadIdx.x < 2) s_base[threadIdx.x] += s_tot[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x < 2) counts[threadIdx.x] = s_base[threadIdx.x];
}

__global__ void __launch_bounds__(256) roi_targets_kernel(int S, int fg_count, const int64_t* __restrict__ keep,
                                                          const float* __restrict__ rois,
                                                          const float* __restrict__ scores,
                                                          const float* __restrict__ a3d, const float* __restrict__ gt,
                                                          const float* __restrict__ gt8,
                                                          const int32_t* __restrict__ assign, int K, int E,
                                                          int normalize, const float* __restrict__ means,
                                                          const float* __restrict__ stds, float* __restrict__ labels,
                                                          float* __restrict__ out_rois, float* __restrict__ out_a3d,
                                                          float* __restrict__ out_scores, float* __restrict__ targets,
                                                          float* __restrict__ inside_w, float* __restrict__ outside_w) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int r = (int)keep[s];
  const int g = assign[r];
  const float lab = s < fg_count ? gt[(size_t)g * 5 + 4] : 0.0f;             // :239-242
  labels[s] = lab;
  float roi[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    roi[c] = rois[(size_t)r * 5 + c];
    out_rois[(size_t)s * 5 + c] = roi[c];
  }
  out_scores[s] = scores[r];
  float anc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (a3d) {
#pragma unroll
    for (int c = 0; c < 7; ++c) anc[c] = a3d[(size_t)r * 7 + c];
  }
  if (out_a3d) {
#pragma unroll
    for (int c = 0; c < 7; ++c) out_a3d[(size_t)s * 7 + c] = anc[c];
  }
  float t[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (E == 4) {                                                             // _compute_targets :151-164
    const float* e = roi + 1;
    const float* q = gt + (size_t)g * 5;
    const float ew = fadd(fsub(e[2], e[0]), 1.0f), eh = fadd(fsub(e[3], e[1]), 1.0f);
    const float diag = __fsqrt_rn(fadd(fmul(ew, ew), fmul(eh, eh)));
    const float ecx = fadd(e[0], fmul(0.5f, ew)), ecy = fadd(e[1], fmul(0.5f, eh));
    const float gw = fadd(fsub(q[2], q[0]), 1.0f), gh = fadd(fsub(q[3], q[1]), 1.0f);
    const float gcx = fadd(q[0], fmul(0.5f, gw)), gcy = fadd(q[1], fmul(0.5f, gh));
    t[0] = fdiv(fsub(gcx, ecx), diag);
    t[1] = fdiv(fsub(gcy, ecy), diag);
    t[2] = logf(fdiv(gw, ew));
    t[3] = logf(fdiv(gh, eh));
  } else {                                                                  // _compute_lidar_targets :134-149
    const float* e = roi + 1;
    const float* q = gt8 + (size_t)g * 8;
    const float rl = fadd(fsub(e[2], e[0]), 1.0f), rw = fadd(fsub(e[3], e[1]), 1.0f);
    const float eh = anc[5];
    const float cx = fadd(e[0], fdiv(rl, 2.0f)), cy = fadd(e[1], fdiv(rw, 2.0f));
    const float diag = __fsqrt_rn(fadd(fmul(rl, rl), fmul(rw, rw)));
    t[0] = fdiv(fsub(q[0], cx), diag);
    t[1] = fdiv(fsub(q[1], cy), diag);
    t[2] = fdiv(fsub(q[2], anc[2]), eh);
    t[3] = logf(fdiv(q[3], rl));
    t[4] = logf(fdiv(q[4], rw));
    t[5] = logf(fdiv(q[5], eh));
    t[6] = q[6];
  }
  if (normalize) {
    for (int c = 0; c < E; ++c) t[c] = fdiv(fsub(t[c], means[c]), stds[c]);
  }
  const int KE = K * E;
  const int cls = (int)lab;                                                  // .long() truncation, :88-96
  for (int c = 0; c < KE; ++c) {
    const bool hit = lab > 0.0f && c >= cls * E && c < (cls + 1) * E;
    targets[(size_t)s * KE + c] = hit ? t[c - cls * E] : 0.0f;
    inside_w[(size_t)s * KE + c] = hit ? 1.0f : 0.0f;
    outside_w[(size_t)s * KE + c] = hit ? 1.0f : 0.0f;                       // (inside > 0).float(), :59
  }
}

}  // namespace b2d

using namespace b2d;

extern "C" size_t b2d_anchor_target_workspace_bytes(int F, int N, int max_gt) {
  if (F <= 0 || N <= 0 || max_gt < 0) return 0;
  return carve_anchor(nullptr, F, N).bytes;
}

extern "C" int b2d_anchor_target_phase1(int F, int N, int max_gt, const float* anchors, const float* gt_boxes,
                                        const int32_t* num_gt, const float* info, float neg_overlap, float pos_overlap,
                                        int clobber, int32_t* counts, void* workspace, size_t workspace_bytes,
                                        void* stream) {
  if (F <= 0 || N <= 0 || !anchors || !gt_boxes || !num_gt || !info || !counts) return B2D_ERR_INVALID_ARG;
  if (max_gt <= 0 || max_gt > kMaxGt) return B2D_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(anchors) & 15u) != 0) return B2D_ERR_INVALID_ARG;
  AnchorWs w = carve_anchor(workspace, F, N);
  if (!workspace || workspace_bytes < w.bytes) return B2D_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  B2D_CUDA(cudaMemsetAsync(w.gt_max, 0, sizeof(uint32_t) * (size_t)F * kMaxGt, st));
  int gx = ceil_div(N, 256);
  if (gx > 4 * kNumSMs) gx = 4 * kNumSMs;
  dim3 grid(gx, F);
  anchor_iou_kernel<<<grid, 256, 0, st>>>(N, max_gt, anchors, gt_boxes, num_gt, info, w);
  B2D_LAUNCHED();
  anchor_label_kernel<<<grid, 256, 0, st>>>(N, max_gt, anchors, gt_boxes, num_gt, neg_overlap, pos_overlap, clobber, w);
  B2D_LAUNCHED();
  anchor_lists_kernel<<<F, 1024, 0, st>>>(N, w, counts);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_anchor_target_phase2(int F, int N, int max_gt, int A, int height, int width, const float* anchors,
                                        const float* gt_boxes, const int32_t* counts, const int64_t* fg_disable,
                                        const int32_t* n_fg_disable, const int64_t* bg_disable,
                                        const int32_t* n_bg_disable, int disable_stride, const float* inside_weights4,
                                        float positive_weight, float* labels, float* bbox_targets, float* inside_w,
                                        float* outside_w, void* workspace, size_t workspace_bytes, void* stream) {
  if (F <= 0 || N <= 0 || A <= 0 || (long long)height * width * A != N || !anchors || !gt_boxes || !counts ||
      !inside_weights4 || !labels || !bbox_targets || !inside_w || !outside_w)
    return B2D_ERR_INVALID_ARG;
  if (max_gt <= 0 || max_gt > kMaxGt) return B2D_ERR_UNSUPPORTED;
  if ((n_fg_disable && !fg_disable) || (n_bg_disable && !bg_disable)) return B2D_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(anchors) & 15u) || (reinterpret_cast<uintptr_t>(bbox_targets) & 15u) ||
      (reinterpret_cast<uintptr_t>(inside_w) & 15u) || (reinterpret_cast<uintptr_t>(outside_w) & 15u))
    return B2D_ERR_INVALID_ARG;
  AnchorWs w = carve_anchor(workspace, F, N);
  if (!workspace || workspace_bytes < w.bytes) return B2D_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  if ((n_fg_disable || n_bg_disable) && disable_stride > 0) {
    dim3 g1(ceil_div(2 * disable_stride, 256), F);
    anchor_disable_kernel<<<g1, 256, 0, st>>>(N, disable_stride, fg_disable, n_fg_disable, bg_disable, n_bg_disable, w);
    B2D_LAUNCHED();
  }
  int gx = ceil_div(N, 256);
  if (gx > 4 * kNumSMs) gx = 4 * kNumSMs;
  dim3 grid(gx, F);
  anchor_emit_kernel<<<grid, 256, 0, st>>>(N, max_gt, A, height * width, anchors, gt_boxes, counts, n_fg_disable,
                                           n_bg_disable, inside_weights4, positive_weight, w, labels, bbox_targets,
                                           inside_w, outside_w);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_proposal_target_phase1(int R, int G, const float* rois, const float* gt_boxes, float fg_thresh,
                                          float bg_hi, float bg_lo, int bg_mode, float* max_overlap,
                                          int32_t* assignment, int32_t* fg_list, int32_t* bg_list, int32_t* counts,
                                          void* stream) {
  if (R < 0 || G <= 0 || !gt_boxes || !counts) return B2D_ERR_INVALID_ARG;
  if (R > 0 && (!rois || !max_overlap || !assignment || !fg_list || !bg_list)) return B2D_ERR_INVALID_ARG;
  roi_assign_kernel<<<1, 1024, 0, as_stream(stream)>>>(R, G, rois, gt_boxes, fg_thresh, bg_hi, bg_lo, bg_mode,
                                                       max_overlap, assignment, fg_list, bg_list, counts);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_proposal_target_phase2(int S, int fg_count, const int64_t* keep_inds, const float* rois,
                                          const float* scores, const float* anchors_3d, const float* gt_boxes,
                                          const float* true_gt_boxes, const int32_t* assignment, int K, int E,
                                          int normalize, const float* means, const float* stds, float* labels,
                                          float* out_rois, float* out_a3d, float* out_scores, float* bbox_targets,
                                          float* inside_w, float* outside_w, void* stream) {
  if (S < 0 || K <= 0 || (E != 4 && E != 7)) return B2D_ERR_INVALID_ARG;
  if (S == 0) return B2D_OK;
  if (!keep_inds || !rois || !scores || !gt_boxes || !assignment || !labels || !out_rois || !out_scores ||
      !bbox_targets || !inside_w || !outside_w)
    return B2D_ERR_INVALID_ARG;
  if (E == 7 && (!anchors_3d || !true_gt_boxes)) return B2D_ERR_INVALID_ARG;
  if (normalize && (!means || !stds)) return B2D_ERR_INVALID_ARG;
  roi_targets_kernel<<<ceil_div(S, 256), 256, 0, as_stream(stream)>>>(
      S, fg_count, keep_inds, rois, scores, anchors_3d, gt_boxes, true_gt_boxes, assignment, K, E, normalize, means,
      stds, labels, out_rois, out_a3d, out_scores, bbox_targets, inside_w, outside_w);
  B2D_LAUNCHED();
  return B2D_OK;
}
