// Library-level entry points: version, status strings, launch counter, and the end-to-end
// host-buffer pipeline (bench.py `e2e`).
#include "common.cuh"

namespace b2d {
std::atomic<uint64_t> g_launches{0};
thread_local int tl_cuda_error = 0;
}  // namespace b2d

using namespace b2d;

extern "C" int b2d_abi_version(void) { return 1; }

extern "C" const char* b2d_status_string(int status) {
  switch (status) {
    case B2D_OK: return "ok";
    case B2D_ERR_INVALID_ARG: return "invalid argument";
    case B2D_ERR_WORKSPACE: return "workspace missing or too small";
    case B2D_ERR_CUDA: return "CUDA runtime error";
    case B2D_ERR_UNSUPPORTED: return "shape outside the compiled limits";
    default: return "unknown status";
  }
}

extern "C" int b2d_last_cuda_error(void) { return tl_cuda_error; }
extern "C" uint64_t b2d_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

namespace {
struct PipeWs {
  float *cls, *pred, *info, *feat, *rois, *scores, *pooled;
  int32_t* num_out;
  void* prop_ws;
  size_t prop_ws_bytes;
  size_t bytes;
};

PipeWs carve_pipe(void* base, int F, int n_loc, int A, int C, int H, int W, int pre, int post, int P) {
  PipeWs w;
  const int N = n_loc * A;
  const int k = (pre > 0 && pre < N) ? pre : N;
  const int mo = (post > 0 && post < k) ? post : k;
  size_t off = 0;
  char* p = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  w.cls = reinterpret_cast<float*>(take(sizeof(float) * (size_t)F * n_loc * 2 * A));
  w.pred = reinterpret_cast<float*>(take(sizeof(float) * (size_t)F * N * 4));
  w.info = reinterpret_cast<float*>(take(sizeof(float) * (size_t)F * 7));
  w.feat = reinterpret_cast<float*>(take(sizeof(float) * (size_t)F * C * H * W));
  w.rois = reinterpret_cast<float*>(take(sizeof(float) * (size_t)F * mo * 5));
  w.scores = reinterpret_cast<float*>(take(sizeof(float) * (size_t)F * mo));
  w.num_out = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (size_t)F));
  w.pooled = reinterpret_cast<float*>(take(sizeof(float) * (size_t)F * mo * C * P * P));
  w.prop_ws_bytes = b2d_proposal_workspace_bytes(F, n_loc, A, pre, post);
  w.prop_ws = take(w.prop_ws_bytes);
  w.bytes = off;
  return w;
}
}  // namespace

extern "C" size_t b2d_pipeline_device_bytes(int F, int n_loc, int A, int C, int H, int W, int pre, int post, int P) {
  if (F <= 0 || n_loc <= 0 || A <= 0 || C <= 0 || H <= 0 || W <= 0 || P <= 0) return 0;
  return carve_pipe(nullptr, F, n_loc, A, C, H, W, pre, post, P).bytes;
}

extern "C" int b2d_proposal_crop_host(int F, int n_loc, int A, int C, int H, int W, const float* cls_prob_host,
                                      const float* bbox_pred_host, const float* info_host, const float* anchors_dev,
                                      const float* feat_host, int pre, int post, double nms_thresh, int P,
                                      float spatial_scale, int sampling_ratio, float* rois_host, float* scores_host,
                                      int32_t* num_out_host, float* pooled_host, void* device_ws,
                                      size_t device_ws_bytes, void* stream) {
  if (F <= 0 || !cls_prob_host || !bbox_pred_host || !info_host || !anchors_dev || !feat_host || !rois_host ||
      !scores_host || !num_out_host || !pooled_host)
    return B2D_ERR_INVALID_ARG;
  PipeWs w = carve_pipe(device_ws, F, n_loc, A, C, H, W, pre, post, P);
  if (!device_ws || device_ws_bytes < w.bytes) return B2D_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int N = n_loc * A;
  const int k = (pre > 0 && pre < N) ? pre : N;
  const int mo = (post > 0 && post < k) ? post : k;
  B2D_CUDA(cudaMemcpyAsync(w.cls, cls_prob_host, sizeof(float) * (size_t)F * n_loc * 2 * A, cudaMemcpyHostToDevice, st));
  B2D_CUDA(cudaMemcpyAsync(w.pred, bbox_pred_host, sizeof(float) * (size_t)F * N * 4, cudaMemcpyHostToDevice, st));
  B2D_CUDA(cudaMemcpyAsync(w.info, info_host, sizeof(float) * (size_t)F * 7, cudaMemcpyHostToDevice, st));
  B2D_CUDA(cudaMemcpyAsync(w.feat, feat_host, sizeof(float) * (size_t)F * C * H * W, cudaMemcpyHostToDevice, st));
  int rc = b2d_proposal(F, n_loc, A, w.cls, w.pred, w.info, anchors_dev, nullptr, pre, post, nms_thresh, 1, w.rois,
                        w.scores, nullptr, nullptr, w.num_out, w.prop_ws, w.prop_ws_bytes, stream);
  if (rc != B2D_OK) return rc;
  rc = b2d_roi_align_forward(F, C, H, W, w.feat, w.rois, F * mo, nullptr, 0, w.num_out, mo, P, P, spatial_scale,
                             sampling_ratio, 0, w.pooled, nullptr, 0, stream);
  if (rc != B2D_OK) return rc;
  B2D_CUDA(cudaMemcpyAsync(rois_host, w.rois, sizeof(float) * (size_t)F * mo * 5, cudaMemcpyDeviceToHost, st));
  B2D_CUDA(cudaMemcpyAsync(scores_host, w.scores, sizeof(float) * (size_t)F * mo, cudaMemcpyDeviceToHost, st));
  B2D_CUDA(cudaMemcpyAsync(num_out_host, w.num_out, sizeof(int32_t) * (size_t)F, cudaMemcpyDeviceToHost, st));
  B2D_CUDA(cudaMemcpyAsync(pooled_host, w.pooled, sizeof(float) * (size_t)F * mo * C * P * P, cudaMemcpyDeviceToHost, st));
  B2D_CUDA(cudaStreamSynchronize(st));
  return B2D_OK;
}
