// Result post-processing and detection-vs-ground-truth matching.
//
//   eval_match_kernel   the confidence-ordered greedy matching of datasets/waymo_eval.py:120-215 (the same loop is in
//                       kitti_eval.py:125-205 and cadc_eval.py:124-200): every detection of a class, in descending
//                       confidence, against the ground truth of ITS frame - overlap with the don't-care boxes, overlap
//                       with the class's boxes (max + first argmax), then true positive / duplicate / low-overlap false
//                       positive under the hit, ignore and difficulty flags.  Frames never interact, detections of one
//                       frame do (the hit flags): one WARP per frame walks the frame's detections in order, lanes
//                       stride over the ground-truth boxes.  Overlaps are computed in float64 with one rounding per
//                       numpy operation (no FMA contraction), so max / argmax and the threshold decisions are those
//                       of the reference's numpy code on the same boxes.
//   voxel_grid_to_pc_kernel   utils/bbox.py:140-162 (bbox_voxel_grid_to_pc), applied in model/test.py:224 to the
//                       detections of a lidar frame before they are stacked into all_boxes.
//
// The overlap function itself lived in the reference's missing utils/eval_utils.py; what is computed
// here is stated in oracle/eval_oracle.py `iou` (PASCAL-VOC overlap with the fork's +1 convention for image boxes,
// axis-aligned footprint for 'bev_aa').
#include "common.cuh"

namespace b2d {

struct EvalArgs {
  const double* det_boxes;          // [n_det, E] in confidence order
  const int32_t* rec_det_offset;    // [n_rec + 1]
  const int32_t* rec_det_index;     // positions (confidence order) of the detections of each frame, ascending
  const int32_t* gt_offset;         // [n_rec + 1]
  const double* gt_boxes;           // [sum G, E]
  const int32_t* gt_flags;          // [sum G]: bit 0 ignore, bits 8.. difficulty
  const int32_t* dc_offset;         // [n_rec + 1]
  const double* dc_boxes;           // [sum D, E]
  int n_rec, E, mode, ignore_dc;
  double ovthresh, ovthresh_dc;
  int32_t* code;                    // [n_det]: 0 nothing recorded, 1 tp, 2 duplicate fp, 3 low-overlap fp
  double* ovmax;                    // [n_det]
  int32_t* jmax;                    // [n_det]
  int32_t* difficulty;              // [n_det]: difficulty of the matched box (codes 1, 2), else -1
  unsigned char* hit;               // [sum G] scratch
};

__device__ __forceinline__ void eval_corners(const double* b, int mode, double (&c)[4]) {
  if (mode == 0) {
    c[0] = b[0], c[1] = b[1], c[2] = b[2], c[3] = b[3];
  } else {
    const double hl = __ddiv_rn(b[3], 2.0), hw = __ddiv_rn(b[4], 2.0);
    c[0] = __dsub_rn(b[0], hl), c[1] = __dsub_rn(b[1], hw), c[2] = __dadd_rn(b[0], hl), c[3] = __dadd_rn(b[1], hw);
  }
}

// inters / ((d area) + (g area) - inters), every numpy operation rounded once
__device__ __forceinline__ double eval_iou(const double (&g)[4], const double (&d)[4], double one) {
  const double iw = fmax(__dadd_rn(__dsub_rn(fmin(g[2], d[2]), fmax(g[0], d[0])), one), 0.0);
  const double ih = fmax(__dadd_rn(__dsub_rn(fmin(g[3], d[3]), fmax(g[1], d[1])), one), 0.0);
  const double inters = __dmul_rn(iw, ih);
  const double ad = __dmul_rn(__dadd_rn(__dsub_rn(d[2], d[0]), one), __dadd_rn(__dsub_rn(d[3], d[1]), one));
  const double ag = __dmul_rn(__dadd_rn(__dsub_rn(g[2], g[0]), one), __dadd_rn(__dsub_rn(g[3], g[1]), one));
  return __ddiv_rn(inters, __dsub_rn(__dadd_rn(ad, ag), inters));
}

__global__ void __launch_bounds__(128) eval_match_kernel(const EvalArgs a) {
  const int rec = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (rec >= a.n_rec) return;
  const int d0 = a.rec_det_offset[rec], d1 = a.rec_det_offset[rec + 1];
  const int g0 = a.gt_offset[rec], G = a.gt_offset[rec + 1] - g0;
  const int c0 = a.dc_offset[rec], D = a.dc_offset[rec + 1] - c0;
  const double one = a.mode == 0 ? 1.0 : 0.0;
  const double ninf = -__longlong_as_double(0x7FF0000000000000ll);
  for (int g = lane; g < G; g += 32) a.hit[g0 + g] = 0;
  __syncwarp();
  for (int q = d0; q < d1; ++q) {
    const int pos = a.rec_det_index[q];
    double d[4];
    eval_corners(a.det_boxes + (size_t)pos * a.E, a.mode, d);
    // overlap with the don't-care boxes: max only (ovmax_dc starts at 0, waymo_eval.py:162)
    double dc = ninf;
    if (a.ignore_dc)
      for (int j = lane; j < D; j += 32) {
        double gb[4];
        eval_corners(a.dc_boxes + (size_t)(c0 + j) * a.E, a.mode, gb);
        dc = fmax(dc, eval_iou(gb, d, one));
      }
    // overlap with the class's boxes: max and FIRST argmax
    double best = ninf;
    int bj = 0x7FFFFFFF;
    for (int j = lane; j < G; j += 32) {
      double gb[4];
      eval_corners(a.gt_boxes + (size_t)(g0 + j) * a.E, a.mode, gb);
      const double ov = eval_iou(gb, d, one);
      if (ov > best) best = ov, bj = j;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      const double ob = __shfl_xor_sync(0xFFFFFFFFu, best, s);
      const int oj = __shfl_xor_sync(0xFFFFFFFFu, bj, s);
      const double od = __shfl_xor_sync(0xFFFFFFFFu, dc, s);
      if (ob > best || (ob == best && oj < bj)) best = ob, bj = oj;
      dc = fmax(dc, od);
    }
    if (lane == 0) {
      const double ovmax_dc = (a.ignore_dc && D > 0) ? dc : 0.0;
      const int jm = G > 0 ? (bj == 0x7FFFFFFF ? 0 : bj) : 0;
      int code = 0, diff = -1;
      if (G > 0 && best > a.ovthresh && ovmax_dc < a.ovthresh_dc) {
        const int fl = a.gt_flags[g0 + jm];
        if (!(fl & 1)) {
          diff = fl >> 8;
          if (!a.hit[g0 + jm]) {
            a.hit[g0 + jm] = 1;
            code = 1;
          } else {
            code = 2;
          }
        }
      } else if (G > 0 && ovmax_dc < a.ovthresh_dc) {
        code = 3;
      }
      a.code[pos] = code;
      a.ovmax[pos] = G > 0 ? best : ninf;
      a.jmax[pos] = jm;
      a.difficulty[pos] = diff;
    }
    __syncwarp();
  }
}

// utils/bbox.py:140-162; fx, fy, x0, y0 are computed by the caller in fp32 exactly as numpy does
__global__ void __launch_bounds__(256) voxel_grid_to_pc_kernel(int n, int stride, float fx, float fy, float x0, float y0,
                                                               int aabb, float* __restrict__ boxes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* b = boxes + (size_t)i * stride;
  b[0] = fadd(fmul(b[0], fx), x0);
  b[1] = fadd(fmul(b[1], fy), y0);
  if (aabb) {
    b[2] = fadd(fmul(b[2], fx), x0);
    b[3] = fadd(fmul(b[3], fy), y0);
  } else {
    b[3] = fmul(b[3], fx);
    b[4] = fmul(b[4], fy);
  }
}

}  // namespace b2d

using namespace b2d;

extern "C" int b2d_eval_match(int n_det, int n_rec, int box_elem, int mode, const double* det_boxes,
                              const int32_t* rec_det_offset, const int32_t* rec_det_index, const int32_t* gt_offset,
                              const double* gt_boxes, const int32_t* gt_flags, const int32_t* dc_offset,
                              const double* dc_boxes, double ovthresh, double ovthresh_dc, int ignore_dc, int32_t* code,
                              double* ovmax, int32_t* jmax, int32_t* difficulty, unsigned char* hit_scratch, void* stream) {
  if (n_det < 0 || n_rec < 0 || (mode != 0 && mode != 1)) return B2D_ERR_INVALID_ARG;
  if ((mode == 0 && box_elem < 4) || (mode == 1 && box_elem < 5)) return B2D_ERR_INVALID_ARG;
  if (n_det == 0 || n_rec == 0) return B2D_OK;
  if (!det_boxes || !rec_det_offset || !rec_det_index || !gt_offset || !dc_offset || !code || !ovmax || !jmax || !difficulty)
    return B2D_ERR_INVALID_ARG;
  EvalArgs a{det_boxes, rec_det_offset, rec_det_index, gt_offset, gt_boxes, gt_flags, dc_offset, dc_boxes, n_rec, box_elem,
             mode, ignore_dc, ovthresh, ovthresh_dc, code, ovmax, jmax, difficulty, hit_scratch};
  eval_match_kernel<<<ceil_div(n_rec, 4), 128, 0, as_stream(stream)>>>(a);
  B2D_LAUNCHED();
  return B2D_OK;
}

extern "C" int b2d_bbox_voxel_grid_to_pc(int n, int row_stride, float fx, float fy, float x0, float y0, int aabb,
                                         float* boxes, void* stream) {
  if (n < 0 || row_stride < (aabb ? 4 : 5)) return B2D_ERR_INVALID_ARG;
  if (n == 0) return B2D_OK;
  if (!boxes) return B2D_ERR_INVALID_ARG;
  voxel_grid_to_pc_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(n, row_stride, fx, fy, x0, y0, aabb, boxes);
  B2D_LAUNCHED();
  return B2D_OK;
}
