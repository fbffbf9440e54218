"""One frame per call (what the reference API issues): proposal_layer + RoIAlign forward, a few iterations.
Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list, or plain for CUDA-event timings.
  python profiles/f1_latency.py [workload] [iters] [TRAIN]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from faster_rcnn_pytorch_multimodal_b200 import ops

wl = sys.argv[1] if len(sys.argv) > 1 else "waymo_test"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
train = len(sys.argv) > 3
cfg = dict(bench.WORKLOADS[wl])
if train:
    cfg.update(pre_nms=12000, post_nms=2000)
dev = torch.device("cuda", 0)
anchors, a3d = bench.anchors_for(cfg, dev)
prob, deltas, feat, info = bench.synth_frames(cfg, 1, dev, 0)
M, P = cfg["post_nms"], 7
pooled = torch.empty(M, cfg["C"], P, P, device=dev)


def one():
    r, _, _, _, n = ops.proposal_batched(prob, deltas, info, anchors, a3d, cfg["A"], cfg["pre_nms"], M, cfg["nms_thresh"])
    ops._roi_align_forward(feat, r.view(-1, 5), (P, P), 1.0 / cfg["stride"], cfg["sampling_ratio"], False, seg_count=n,
                           seg_stride=M, out=pooled)


for _ in range(3):
    one()
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    one()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
# the same calls back to back (what bench.py's latency_ms_f1 reports: the GPU never waits for the host)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    one()
e1.record()
torch.cuda.synchronize()
print(f"{wl}{' TRAIN' if train else ''}: one frame per call: synchronised median {ts[len(ts) // 2]:.4f} ms, min {ts[0]:.4f} ms; "
      f"back to back {e0.elapsed_time(e1) / iters:.4f} ms")
