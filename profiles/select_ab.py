"""Device time of the proposal stage alone on the bench workload (CUDA events): python profiles/select_ab.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from faster_rcnn_pytorch_multimodal_b200 import ops
F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
cfg = bench.WORKLOADS["waymo_test"]
anchors, _ = bench.anchors_for(cfg, dev)
prob, deltas, feat, info = bench.synth_frames(cfg, F, dev, 0)
fn = lambda: ops.proposal_batched(prob, deltas, info, anchors, None, cfg["A"], cfg["pre_nms"], cfg["post_nms"], cfg["nms_thresh"], batch_index_stride=1)
for _ in range(3): fn()
torch.cuda.synchronize()
ts = []
for _ in range(15):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"{os.path.basename(os.environ.get('B2D_LIB_PATH', 'default'))}: proposal stage, {F} frames: median {np.median(ts) * 1e3:.1f} us")
