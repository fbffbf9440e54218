"""Item statistics of the rows kernel on a bench workload (A/B build with -DB2D_AB_COUNT=1):
whole-RoI items that took a pool tile / missed the lock, items on the bin-row path, bin-rows stored that way.
  B2D_LIB_PATH=_ab/count.so python profiles/rows_count.py [workload] [frames]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from faster_rcnn_pytorch_multimodal_b200 import ops, _lib

wl = sys.argv[1] if len(sys.argv) > 1 else "waymo_test"
F = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda", 0)
cfg = bench.WORKLOADS[wl]
anchors, a3d = bench.anchors_for(cfg, dev)
prob, deltas, feat, info = bench.synth_frames(cfg, F, dev, 0)
M = cfg["post_nms"]
rois, scores, _, _, num = ops.proposal_batched(prob, deltas, info, anchors, None, cfg["A"], cfg["pre_nms"], M, cfg["nms_thresh"],
                                               batch_index_stride=1)
pooled = torch.empty(F * M, cfg["C"], 7, 7, device=dev)
lib = _lib.lib()
buf = (ctypes.c_ulonglong * 8)()
lib.b2d_ab_counters(buf, 1)
ops._roi_align_forward(feat, rois.view(-1, 5), (7, 7), 1.0 / cfg["stride"], cfg["sampling_ratio"], False, seg_count=num, seg_stride=M,
                       out=pooled)
lib.b2d_ab_counters(buf, 1)
groups = cfg["C"] // 32
n = int(num.sum())
r = rois.view(-1, 5)[:, 1:]
h = ((r[:, 3] - r[:, 1]) / cfg["stride"])
w = ((r[:, 2] - r[:, 0]) / cfg["stride"])
valid = (torch.arange(M, device=dev)[None, :] < num[:, None]).reshape(-1)
h, w = h[valid], w[valid]
q = torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9], device=dev)
print(f"{wl}: {n} RoIs in {F} frames; per (RoI, channel group): tile {buf[0] / groups / n:.3f}, lock miss {buf[1] / groups / n:.3f}, "
      f"bin-row items {buf[2] / groups / n:.3f}, bin-rows on the slow path {buf[3] / groups / n:.3f} of 7")
print("height quantiles (feature rows):", [round(float(x), 1) for x in torch.quantile(h, q)])
print("width quantiles  (feature cols):", [round(float(x), 1) for x in torch.quantile(w, q)])
