// Tail of the detection head over the MC-dropout stack: ONE launch that turns the T stacked
// head passes of a batch of frames into what b2d_final_detections consumes.
//
// Reference pieces restated (the method that strings them together, Network.test_frame, lives in the missing
// lib/nets/network.py - SURVEY.md F1 - so the composition below is [INFERRED]; each piece is the reference's):
//   * per-class de-normalisation  bbox_pred * stds + means      model/config.py:219-223 (train_val.py applies the
//                                                                same tile at save time)
//   * mean over the T samples and compute_bbox_var              utils/loss_utils.py:114-120 (single-pass formula)
//   * lidar_3d_bbox_transform_inv / bbox_transform_inv + clip   model/bbox_transform.py:174-233 / :75-105,235-257
//   * lidar_3d_uncertainty_transform_inv (aleatoric + epistemic) model/bbox_transform.py:132-169
//   * softmax, mean class probability, categorical_entropy,
//     categorical_mutual_information                            utils/loss_utils.py:122-141
// Layouts: bbox_pred [T, F, R, K*E], cls_score [T, F, R, K] (T head passes over the same F*R RoIs, stacked),
// rois [F, R, 5], anchors_3d [F, R, 7], info [F, 7]; outputs [F, R, ...] exactly as b2d_final_detections reads
// them (pred_boxes [F,R,K*E], cls probabilities [F,R,K], per-roi and per-class-box uncertainty columns).
// The first `n_box_blocks` CTAs own one (f, r, class-box element) each; the rest own one (f, r) of the class part.
#include "common.cuh"

namespace b2d {

constexpr int kMaxElem = 8;

struct HeadTailArgs {
  int F, T, R, K, E, mode, use_scale, clip, n_box_blocks;
  float mean[kMaxElem], stdv[kMaxElem];
  const float *bbox_pred, *cls_score, *rois, *anchors_3d, *info, *a_var_in;
  float *boxes, *probs, *e_var, *a_var, *entropy, *mutual_info;
};

__global__ void __launch_bounds__(256) head_tail_kernel(const HeadTailArgs a) {
  const int K = a.K, E = a.E, T = a.T;
  const int KE = K * E;
  if ((int)blockIdx.x < a.n_box_blocks) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)a.F * a.R * KE;
    if (idx >= total) return;
    const int j = (int)(idx % KE);
    const long long fr = idx / KE;            // f * R + r
    const int f = (int)(fr / a.R);
    const int e = j % E;
    // de-normalised samples, summed in sample order like torch.sum over dim 0
    float s = 0.0f, s2 = 0.0f;
    const float sd = a.stdv[e], mu = a.mean[e];
    for (int t = 0; t < T; ++t) {
      const float v = fadd(fmul(__ldg(a.bbox_pred + (size_t)t * total + idx), sd), mu);
      s = fadd(s, v);
      s2 = fadd(s2, fmul(v, v));
    }
    const float d = fdiv(s, (float)T);        // mean prediction of element e of class j / E
    float var = 0.0f;
    if (T > 1) var = fmaxf(fdiv(fadd(s2, fdiv(-fmul(s, s), (float)T)), (float)(T - 1)), 0.0f);
    const float* roi = a.rois + fr * 5 + 1;
    float r0 = roi[0], r1 = roi[1], r2 = roi[2], r3 = roi[3];
    if (a.use_scale) {
      const float sc = a.info[f * 7 + 6];
      r0 = fdiv(r0, sc), r1 = fdiv(r1, sc), r2 = fdiv(r2, sc), r3 = fdiv(r3, sc);
    }
    const float av = a.a_var_in ? __ldg(a.a_var_in + idx) : 0.0f;
    float box, ev = var, aa = av;
    if (a.mode == 1) {
      // lidar: the codec is built on the RoI's AABB extent and the anchor's z / height
      const float rl = fadd(fsub(r2, r0), 1.0f), rw = fadd(fsub(r3, r1), 1.0f);
      const float* anc = a.anchors_3d + fr * 7;
      const float hh = anc[5];
      float ext;                              // the extent element e scales with
      if (e == 0 || e == 3) ext = rl;
      else if (e == 1 || e == 4) ext = rw;
      else ext = hh;
      if (e < 2) {
        const float diag = __fsqrt_rn(fadd(fmul(rl, rl), fmul(rw, rw)));
        const float c = e == 0 ? fadd(r0, fdiv(rl, 2.0f)) : fadd(r1, fdiv(rw, 2.0f));
        box = fadd(fmul(d, diag), c);
      } else if (e == 2) {
        box = fadd(fmul(d, hh), anc[2]);
      } else if (e < 6) {
        box = fmul(expf(d), ext);
      } else {
        box = d;
      }
      // lidar_3d_uncertainty_transform_inv: (x, y, z) scale with the extent, (l, w, h) go through exp(.) - 1, squared
      auto uc = [&](float u) {
        const float v = e < 3 ? fmul(u, ext) : (e < 6 ? fsub(expf(u), 1.0f) : u);
        return fmul(v, v);
      };
      ev = uc(var);
      aa = uc(av);
    } else {
      // image: bbox_transform_inv of class box j / 4 (needs the class's four deltas) + clip_boxes
      const float w = fadd(fsub(r2, r0), 1.0f), h = fadd(fsub(r3, r1), 1.0f);
      const float diag = __fsqrt_rn(fadd(fmul(w, w), fmul(h, h)));
      // centre / size deltas of this class box: element e pairs with e ^ 2 (x <-> w, y <-> h)
      float s_o = 0.0f;
      const int eo = e ^ 2;
      const float sdo = a.stdv[eo], muo = a.mean[eo];
      const long long idx_o = idx - e + eo;
      for (int t = 0; t < T; ++t) s_o = fadd(s_o, fadd(fmul(__ldg(a.bbox_pred + (size_t)t * total + idx_o), sdo), muo));
      const float d_o = fdiv(s_o, (float)T);
      const float dc = e < 2 ? d : d_o, ds = e < 2 ? d_o : d;      // centre delta, size delta of this axis
      const bool xaxis = (e & 1) == 0;
      const float len = xaxis ? w : h;
      const float ctr = fadd(xaxis ? r0 : r1, fmul(0.5f, len));
      const float pc = fadd(fmul(dc, diag), ctr);
      const float half = fmul(0.5f, fmul(expf(ds), len));
      box = e < 2 ? fsub(pc, half) : fadd(pc, half);
      if (a.clip) {
        const float* inf = a.info + f * 7;
        const float lo = xaxis ? inf[0] : inf[2], hi = fsub(xaxis ? inf[1] : inf[3], 1.0f);
        box = clampf(box, lo, hi);
      }
      // the image flavour of the uncertainty transform is unusable in the reference: raw variances
    }
    a.boxes[idx] = box;
    if (a.e_var) a.e_var[idx] = ev;
    if (a.a_var) a.a_var[idx] = aa;
    return;
  }
  // ---- class part: one thread per (f, r)
  const long long i = (long long)(blockIdx.x - a.n_box_blocks) * blockDim.x + threadIdx.x;
  const long long n = (long long)a.F * a.R;
  if (i >= n) return;
  float mean[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) mean[c] = 0.0f;
  float neg_ent_sum = 0.0f;  // sum_t sum_c p log2 p
  for (int t = 0; t < T; ++t) {
    const float* z = a.cls_score + ((size_t)t * n + i) * K;
    float mx = -INFINITY;
    for (int c = 0; c < K; ++c) mx = fmaxf(mx, __ldg(z + c));
    float den = 0.0f;
    for (int c = 0; c < K; ++c) den += expf(__ldg(z + c) - mx);
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      if (c < K) {
        const float p = expf(__ldg(z + c) - mx) / den;
        mean[c] += p;
        acc += p * log2f(p);
      }
    }
    neg_ent_sum += acc;
  }
  float total = 0.0f;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    if (c < K) {
      const float p = mean[c] / (float)T;
      a.probs[i * K + c] = p;
      total += p * log2f(p);
    }
  }
  total = -total;
  if (a.entropy) a.entropy[i] = total;
  if (a.mutual_info) a.mutual_info[i] = neg_ent_sum / (float)T + total;
}

}  // namespace b2d

using namespace b2d;

extern "C" int b2d_head_tail_decode(int num_frames, int T, int R, int K, int E, const float* bbox_pred,
                                    const float* cls_score, const float* rois, const float* anchors_3d,
                                    const float* info, const float* a_bbox_var_in, const float* means_host,
                                    const float* stds_host, int mode, int use_scale, int clip, float* boxes,
                                    float* probs, float* e_bbox_var, float* a_bbox_var, float* entropy,
                                    float* mutual_info, void* stream) {
  if (num_frames <= 0 || T <= 0 || R < 0 || K <= 0 || E <= 0 || mode < 0 || mode > 1) return B2D_ERR_INVALID_ARG;
  if (R == 0) return B2D_OK;
  if (!bbox_pred || !cls_score || !rois || !info || !boxes || !probs || !means_host || !stds_host)
    return B2D_ERR_INVALID_ARG;
  if ((mode == 1 && (E != 7 || !anchors_3d)) || (mode == 0 && E != 4)) return B2D_ERR_INVALID_ARG;
  if (K > 16) return B2D_ERR_UNSUPPORTED;
  if (a_bbox_var && !a_bbox_var_in) return B2D_ERR_INVALID_ARG;
  HeadTailArgs a{};
  a.F = num_frames, a.T = T, a.R = R, a.K = K, a.E = E, a.mode = mode, a.use_scale = use_scale, a.clip = clip;
  for (int e = 0; e < E; ++e) {
    a.mean[e] = means_host[e];
    a.stdv[e] = stds_host[e];
  }
  a.bbox_pred = bbox_pred, a.cls_score = cls_score, a.rois = rois, a.anchors_3d = anchors_3d, a.info = info;
  a.a_var_in = a_bbox_var_in;
  a.boxes = boxes, a.probs = probs, a.e_var = e_bbox_var, a.a_var = a_bbox_var, a.entropy = entropy;
  a.mutual_info = mutual_info;
  const long long n_box = (long long)num_frames * R * K * E, n_cls = (long long)num_frames * R;
  a.n_box_blocks = (int)((n_box + 255) / 256);
  const int blocks = a.n_box_blocks + (int)((n_cls + 255) / 256);
  head_tail_kernel<<<blocks, 256, 0, as_stream(stream)>>>(a);
  B2D_LAUNCHED();
  return B2D_OK;
}
