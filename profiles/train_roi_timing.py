"""RoIAlign forward / backward device time on the train workload's shapes (CUDA events, median).
  python profiles/train_roi_timing.py [frames] [rois_per_frame]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from faster_rcnn_pytorch_multimodal_b200 import ops

F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
R = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
cfg = dict(bench.WORKLOADS["waymo_test"])
anchors, a3d = bench.anchors_for(cfg, dev)
prob, deltas, feat, info = bench.synth_frames(cfg, F, dev, 0)
rois, scores, _, _, num = ops.proposal_batched(prob, deltas, info, anchors, None, cfg["A"], 12000, 2000, 0.7, batch_index_stride=1)
sampled = torch.cat([rois[f, :R] for f in range(F)]).contiguous()
P = 7
grad_out = torch.randn(F * R, cfg["C"], P, P, device=dev)


def med(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


seg = torch.full((F,), R, dtype=torch.int32, device=dev) if len(sys.argv) > 3 else None      # third argument: padded per-frame layout
kw = dict(seg_count=seg, seg_stride=R) if seg is not None else {}
fwd = med(lambda: ops._roi_align_forward(feat, sampled, (P, P), 1.0 / 16, 2, False, **kw))
bwd = med(lambda: ops._roi_align_backward(grad_out, sampled, tuple(feat.shape), (P, P), 1.0 / 16, 2, False, **kw))
alg_f = F * cfg["C"] * cfg["Hf"] * cfg["Wf"] * 4 + F * R * cfg["C"] * 49 * 4
print(f"F={F} R={R}/frame{' (segmented)' if seg is not None else ''}: forward {fwd:.3f} ms ({alg_f / fwd / 1e6:.0f} GB/s), backward {bwd:.3f} ms ({alg_f / bwd / 1e6:.0f} GB/s)")
