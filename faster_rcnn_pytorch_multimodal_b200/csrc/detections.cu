// Final per-class detection filter, batched over frames and classes on the device.
//
// Replaces nms_hstack_torch / filter_and_draw_prep (utils/filter_predictions.py:45-130) and the
// max-dets filter of the test loop (model/test.py:213-221).  The reference does, per class c >= 1
// and per frame, a boolean-mask compaction, a torchvision NMS and 2-10 device->host copies; here one
// CTA per (frame, class) produces the padded record the end-of-stream gather ships:
//   1. inds   = rois with cls_score[:, c] > thresh                              (:46)
//   2. order  = inds by descending score (ties: lower roi index first)
//   3. boxes  = pred_boxes[inds, c*E:(c+1)*E]; mode 0 (image): clamped to the frame (:82-91)
//               mode 1 (lidar): NMS runs on the un-rotated AABB of (xc, yc, l, w)   (:58-62)
//               mode 2: image boxes as they are (nms_hstack_torch called on its own)
//   4. greedy NMS, torchvision semantics (nms.cu)                               (:67,69)
//   5. max_dets: keep scores >= the max_dets-th best kept score                 (test.py:213-221)
//   6. rows [box(E), score] in descending score order, source roi index, and the gathered
//      per-roi / per-class-box uncertainty rows                                 (:23-43,113-124)
#include "common.cuh"

namespace b2d {

constexpr int kDetThreads = 256;      // up to 1024 RoIs per frame (the test configurations: 300)
constexpr int kDetThreadsBig = 1024;  // more (cfg.TRAIN.RPN_POST_NMS_TOP_N = 2000 RoIs through filter_and_draw_prep)
constexpr int kDetMaxRois = 4096;     // shared-memory staging: 29 bytes per (padded) RoI, 116 KB at the limit

__device__ __forceinline__ bool det_iou_exceeds(const float4 a, const float4 b, const float thr_f) {
  const float w = fmaxf(0.0f, fsub(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  const float h = fmaxf(0.0f, fsub(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  const float inter = fmul(w, h);
  if (!(inter > 0.0f) && thr_f >= 0.0f) return false;  // iou is 0, -0 or NaN: never > thr
  const float area_a = fmul(fsub(a.z, a.x), fsub(a.w, a.y));
  const float area_b = fmul(fsub(b.z, b.x), fsub(b.w, b.y));
  const float iou = fdiv(inter, fsub(fadd(area_a, area_b), inter));
  return iou > thr_f;
}

struct DetArgs {
  const float* cls_score;   // [F, R, K]
  const float* pred_boxes;  // [F, R, K*E]
  const int32_t* num_rois;  // [F] or null
  const float* info;        // [F, 7]
  const float* uc_row;      // [F, R, Urow] or null
  const float* uc_cls;      // [F, R, Ucls, K*E] or null
  int R, K, E, Urow, Ucls, lidar, max_dets, max_out;
  float thresh, nms_thr_f;
  float* dets;              // [F, K, max_out, E+1]
  int32_t* det_roi;         // [F, K, max_out]
  float* out_uc_row;        // [F, K, max_out, Urow] or null
  float* out_uc_cls;        // [F, K, max_out, Ucls*E] or null
  int32_t* counts;          // [F, K]
};

static size_t det_smem_bytes(int n_pad) { return (size_t)n_pad * (16 + 8 + 4 + 1); }

__global__ void __launch_bounds__(kDetThreadsBig) final_detections_kernel(DetArgs a, int n_pad) {
  extern __shared__ __align__(16) unsigned char det_smem[];
  float4* s_box = reinterpret_cast<float4*>(det_smem);                                     // NMS boxes in sorted order
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(s_box + n_pad);
  int* s_keep = reinterpret_cast<int*>(s_key + n_pad);
  unsigned char* s_dead = reinterpret_cast<unsigned char*>(s_keep + n_pad);
  __shared__ int s_n, s_m;
  const int kDetThreads = (int)blockDim.x;
  const int c = blockIdx.x + 1, f = blockIdx.y;
  const int tid = threadIdx.x;
  const int R = a.R, K = a.K, E = a.E;
  const int n = a.num_rois ? min(a.num_rois[f], R) : R;
  const float* score = a.cls_score + (size_t)f * R * K + c;
  const float* boxes = a.pred_boxes + (size_t)f * R * K * E + (size_t)c * E;
  if (tid == 0) s_n = 0;
  __syncthreads();
  // 1-2. candidates as composite keys (score desc, index asc); non-candidates sort last (key 0)
  int local = 0;
  for (int i = tid; i < n_pad; i += kDetThreads) {
    unsigned long long key = 0ull;
    if (i < n) {
      const float s = score[(size_t)i * K];
      if (s > a.thresh) {
        key = composite_key(score_key(s), (uint32_t)i);
        ++local;
      }
    }
    s_key[i] = key;
  }
  if (local) atomicAdd(&s_n, local);
  __syncthreads();
  const int m0 = s_n;
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int p = tid; p < (n_pad >> 1); p += kDetThreads) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const int q = i | j;
        const unsigned long long x = s_key[i], y = s_key[q];
        const bool desc = (i & k) == 0;
        if ((x < y) == desc) {
          s_key[i] = y;
          s_key[q] = x;
        }
      }
      __syncthreads();
    }
  }
  // 3. NMS boxes of the sorted candidates
  const float* inf = a.info + (size_t)f * 7;
  const float fw = fsub(inf[1], inf[0]), fh = fsub(inf[3], inf[2]), scale = inf[6];
  const float xmax = fsub(fdiv(fw, scale), 1.0f), ymax = fsub(fdiv(fh, scale), 1.0f);
  for (int i = tid; i < m0; i += kDetThreads) {
    const float* b = boxes + (size_t)composite_index(s_key[i]) * K * E;
    float4 nb;
    if (a.lidar == 1) {
      nb = make_float4(fsub(b[0], fdiv(b[3], 2.0f)), fsub(b[1], fdiv(b[4], 2.0f)), fadd(b[0], fdiv(b[3], 2.0f)),
                       fadd(b[1], fdiv(b[4], 2.0f)));
    } else if (a.lidar == 2) {
      nb = make_float4(b[0], b[1], b[2], b[3]);
    } else {
      // torch.clamp_min / clamp_max propagate NaN
      nb = make_float4(b[0] != b[0] ? b[0] : fmaxf(b[0], 0.0f), b[1] != b[1] ? b[1] : fmaxf(b[1], 0.0f),
                       b[2] != b[2] ? b[2] : fminf(b[2], xmax), b[3] != b[3] ? b[3] : fminf(b[3], ymax));
    }
    s_box[i] = nb;
    s_dead[i] = 0;
  }
  if (tid == 0) s_m = 0;
  __syncthreads();
  // 4. greedy sweep: the next live candidate is kept and suppresses everything after it
  for (int i = 0; i < m0; ++i) {
    if (s_dead[i]) continue;                 // uniform: shared flag, barrier below orders the writes
    if (tid == 0) s_keep[s_m++] = i;
    const float4 bi = s_box[i];
    for (int j = i + 1 + tid; j < m0; j += kDetThreads)
      if (!s_dead[j] && det_iou_exceeds(bi, s_box[j], a.nms_thr_f)) s_dead[j] = 1;
    __syncthreads();
  }
  __syncthreads();
  // 5. max_dets: a prefix of the (score-sorted) kept list, extended over ties of the cut-off score
  int m = s_m;
  if (a.max_dets > 0 && m > a.max_dets) {
    const unsigned long long cut = s_key[s_keep[a.max_dets - 1]] >> 32;
    int mm = a.max_dets;
    while (mm < m && (s_key[s_keep[mm]] >> 32) == cut) ++mm;
    m = mm;
  }
  if (m > a.max_out) m = a.max_out;
  // 6. records
  const size_t slot0 = ((size_t)f * K + c) * a.max_out;
  if (tid == 0) a.counts[f * K + c] = m;
  for (int t = tid; t < a.max_out; t += kDetThreads) {
    float* d = a.dets + (slot0 + t) * (E + 1);
    if (t < m) {
      const int i = s_keep[t];
      const int roi = (int)composite_index(s_key[i]);
      const float* b = boxes + (size_t)roi * K * E;
      if (a.lidar == 1) {
        for (int e = 0; e < E; ++e) d[e] = b[e];
      } else {
        const float4 nb = s_box[i];
        d[0] = nb.x;
        d[1] = nb.y;
        d[2] = nb.z;
        d[3] = nb.w;
        for (int e = 4; e < E; ++e) d[e] = b[e];
      }
      d[E] = score[(size_t)roi * K];
      a.det_roi[slot0 + t] = roi;
      if (a.out_uc_row)
        for (int u = 0; u < a.Urow; ++u)
          a.out_uc_row[(slot0 + t) * a.Urow + u] = a.uc_row[((size_t)f * R + roi) * a.Urow + u];
      if (a.out_uc_cls)
        for (int u = 0; u < a.Ucls; ++u)
          for (int e = 0; e < E; ++e)
            a.out_uc_cls[((slot0 + t) * a.Ucls + u) * E + e] =
                a.uc_cls[(((size_t)f * R + roi) * a.Ucls + u) * K * E + (size_t)c * E + e];
    } else {
      for (int e = 0; e <= E; ++e) d[e] = 0.0f;
      a.det_roi[slot0 + t] = -1;
      if (a.out_uc_row)
        for (int u = 0; u < a.Urow; ++u) a.out_uc_row[(slot0 + t) * a.Urow + u] = 0.0f;
      if (a.out_uc_cls)
        for (int u = 0; u < a.Ucls * E; ++u) a.out_uc_cls[(slot0 + t) * a.Ucls * E + u] = 0.0f;
    }
  }
}

// class 0 (background) rows: zero counts
__global__ void final_detections_bg_kernel(int F, int K, int32_t* counts) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < F) counts[f * K] = 0;
}

}  // namespace b2d

using namespace b2d;

extern "C" int b2d_final_detections(int F, int R, int K, int E, const float* cls_score, const float* pred_boxes,
                                    const int32_t* num_rois, const float* info, int lidar, float score_thresh,
                                    double nms_thresh, int max_dets, int max_out, const float* uc_row, int n_uc_row,
                                    const float* uc_cls, int n_uc_cls, float* dets, int32_t* det_roi,
                                    float* out_uc_row, float* out_uc_cls, int32_t* counts, void* stream) {
  if (F <= 0 || R < 0 || K < 1 || E < 4 || max_out <= 0 || !counts || !dets || !det_roi || !info)
    return B2D_ERR_INVALID_ARG;
  if (R > kDetMaxRois) return B2D_ERR_UNSUPPORTED;
  if (lidar < 0 || lidar > 2 || (lidar == 1 && E < 5)) return B2D_ERR_INVALID_ARG;
  if ((n_uc_row > 0) != (uc_row && out_uc_row) || (n_uc_cls > 0) != (uc_cls && out_uc_cls)) return B2D_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  final_detections_bg_kernel<<<ceil_div(F, 128), 128, 0, st>>>(F, K, counts);
  B2D_LAUNCHED();
  if (K == 1) return B2D_OK;
  if (R > 0 && (!cls_score || !pred_boxes)) return B2D_ERR_INVALID_ARG;
  DetArgs a{cls_score, pred_boxes, num_rois, info, uc_row, uc_cls, R, K, E, n_uc_row, n_uc_cls, lidar, max_dets, max_out,
            score_thresh, float_floor_of(nms_thresh), dets, det_roi, n_uc_row > 0 ? out_uc_row : nullptr,
            n_uc_cls > 0 ? out_uc_cls : nullptr, counts};
  dim3 grid(K - 1, F);
  int n_pad = 1;
  while (n_pad < R) n_pad <<= 1;
  const size_t smem = det_smem_bytes(n_pad);
  if (smem > 48 * 1024)
    B2D_CUDA(cudaFuncSetAttribute(final_detections_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  final_detections_kernel<<<grid, R > 1024 ? kDetThreadsBig : kDetThreads, smem, st>>>(a, n_pad);
  B2D_LAUNCHED();
  return B2D_OK;
}
