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
  size_t prop_ws_bytes;   // per frame
  void* roi_ws;
  size_t roi_ws_bytes;    // per frame
  size_t bytes;
};

// Every region is laid out frame-major so that frame f can be processed on its own.
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
  w.prop_ws_bytes = align_up(b2d_proposal_workspace_bytes(1, n_loc, A, pre, post), 256);
  w.prop_ws = take(w.prop_ws_bytes * F);
  w.roi_ws_bytes = align_up(b2d_roi_align_workspace_bytes(1, C, H, W, mo, mo), 256);
  w.roi_ws = take(w.roi_ws_bytes * F);
  w.bytes = off;
  return w;
}
}  // namespace

extern "C" size_t b2d_pipeline_device_bytes(int F, int n_loc, int A, int C, int H, int W, int pre, int post, int P) {
  if (F <= 0 || n_loc <= 0 || A <= 0 || C <= 0 || H <= 0 || W <= 0 || P <= 0) return 0;
  return carve_pipe(nullptr, F, n_loc, A, C, H, W, pre, post, P).bytes;
}

// Frames are independent, so the call is a three-stage pipeline over single frames: H2D copies on
// one side stream, the kernels on the caller's stream, D2H copies on a second side stream.  PCIe is
// full duplex, so frame f+1 uploads while frame f computes and frame f-1 downloads.
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
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_start = nullptr, ev_in = nullptr, ev_done = nullptr;
  int rc = B2D_OK;
  auto fail = [&](cudaError_t e) {
    if (e != cudaSuccess && rc == B2D_OK) rc = cuda_fail(e);
    return e != cudaSuccess;
  };
  if (fail(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)) ||
      fail(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking)) ||
      fail(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming)) ||
      fail(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming)) ||
      fail(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming))) {
    // fall through to the cleanup below
  } else {
    // the side streams start after whatever the caller already queued on `stream`
    fail(cudaEventRecord(ev_start, st));
    fail(cudaStreamWaitEvent(s_in, ev_start, 0));
    fail(cudaStreamWaitEvent(s_out, ev_start, 0));
    const size_t n_cls = (size_t)n_loc * 2 * A, n_pred = (size_t)N * 4, n_feat = (size_t)C * H * W;
    const size_t n_pool = (size_t)mo * C * P * P;
    for (int f = 0; f < F && rc == B2D_OK; ++f) {
      fail(cudaMemcpyAsync(w.cls + f * n_cls, cls_prob_host + f * n_cls, sizeof(float) * n_cls, cudaMemcpyHostToDevice, s_in));
      fail(cudaMemcpyAsync(w.pred + f * n_pred, bbox_pred_host + f * n_pred, sizeof(float) * n_pred, cudaMemcpyHostToDevice, s_in));
      fail(cudaMemcpyAsync(w.info + f * 7, info_host + f * 7, sizeof(float) * 7, cudaMemcpyHostToDevice, s_in));
      fail(cudaMemcpyAsync(w.feat + f * n_feat, feat_host + f * n_feat, sizeof(float) * n_feat, cudaMemcpyHostToDevice, s_in));
      fail(cudaEventRecord(ev_in, s_in));
      fail(cudaStreamWaitEvent(st, ev_in, 0));
      if (rc != B2D_OK) break;
      // col0 of the RoIs is the frame index within this call; the frame is processed alone, so its
      // kernels see frame index 0 and the batch index is patched in through batch_index_stride
      int r2 = b2d_proposal(1, n_loc, A, w.cls + f * n_cls, w.pred + f * n_pred, w.info + f * 7, anchors_dev, nullptr,
                            pre, post, nms_thresh, 0, w.rois + (size_t)f * mo * 5, w.scores + (size_t)f * mo, nullptr,
                            nullptr, w.num_out + f, static_cast<char*>(w.prop_ws) + f * w.prop_ws_bytes, w.prop_ws_bytes,
                            stream);
      if (r2 == B2D_OK)
        r2 = b2d_roi_align_forward(1, C, H, W, w.feat + f * n_feat, w.rois + (size_t)f * mo * 5, mo, nullptr, 0,
                                   w.num_out + f, mo, P, P, spatial_scale, sampling_ratio, 0, w.pooled + f * n_pool,
                                   static_cast<char*>(w.roi_ws) + f * w.roi_ws_bytes, w.roi_ws_bytes, stream);
      if (r2 != B2D_OK) {
        rc = r2;
        break;
      }
      fail(cudaEventRecord(ev_done, st));
      fail(cudaStreamWaitEvent(s_out, ev_done, 0));
      fail(cudaMemcpyAsync(rois_host + (size_t)f * mo * 5, w.rois + (size_t)f * mo * 5, sizeof(float) * mo * 5, cudaMemcpyDeviceToHost, s_out));
      fail(cudaMemcpyAsync(scores_host + (size_t)f * mo, w.scores + (size_t)f * mo, sizeof(float) * mo, cudaMemcpyDeviceToHost, s_out));
      fail(cudaMemcpyAsync(num_out_host + f, w.num_out + f, sizeof(int32_t), cudaMemcpyDeviceToHost, s_out));
      fail(cudaMemcpyAsync(pooled_host + f * n_pool, w.pooled + f * n_pool, sizeof(float) * n_pool, cudaMemcpyDeviceToHost, s_out));
    }
    // col0 = frame index, as b2d_proposal(batch_index_stride = 1) would have written it
    fail(cudaStreamSynchronize(s_out));
    fail(cudaStreamSynchronize(s_in));
    fail(cudaStreamSynchronize(st));
    if (rc == B2D_OK)
      for (int f = 1; f < F; ++f)
        for (int i = 0; i < num_out_host[f]; ++i) rois_host[((size_t)f * mo + i) * 5] = (float)f;
  }
  if (ev_start) cudaEventDestroy(ev_start);
  if (ev_in) cudaEventDestroy(ev_in);
  if (ev_done) cudaEventDestroy(ev_done);
  if (s_in) cudaStreamDestroy(s_in);
  if (s_out) cudaStreamDestroy(s_out);
  return rc;
}
