"""One RoIAlign forward launch on the bench workload (16 frames) for ncu metric passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from faster_rcnn_pytorch_multimodal_b200 import ops
from faster_rcnn_pytorch_multimodal_b200.layer_utils.snippets import generate_anchors_pre
F = 16
dev = torch.device('cuda', 0)
cfg = bench.CFG
anchors, _ = generate_anchors_pre(cfg["Hf"], cfg["Wf"], cfg["stride"], bench.SCALES, bench.RATIOS, 1.0, device=dev)
prob, deltas, feat, info = bench.synth_frames(cfg, F, dev, 0)
ns = os.environ.pop("B2D_NOSTORE", None)
rois, scores, _, _, num = ops.proposal_batched(prob, deltas, info, anchors, None, cfg["A"], cfg["pre_nms"], cfg["post_nms"], cfg["nms_thresh"], batch_index_stride=1)
if ns: os.environ["B2D_NOSTORE"] = ns
pooled = torch.empty(F * 300, 1024, 7, 7, device=dev)
ops._roi_align_forward(feat, rois.view(-1, 5), (7, 7), 1.0 / 16, 2, False, seg_count=num, seg_stride=300, out=pooled)
torch.cuda.synchronize()
