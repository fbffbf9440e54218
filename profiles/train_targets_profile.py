import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, bench
from faster_rcnn_pytorch_multimodal_b200 import ops
from faster_rcnn_pytorch_multimodal_b200.layer_utils.batched_targets import train_targets_batched
from faster_rcnn_pytorch_multimodal_b200.model.config import cfg as pcfg
pcfg.NET_TYPE = "image"
cfg = bench.WORKLOADS["waymo_train"]
dev = torch.device("cuda", 0)
F = 8
A, K = cfg["A"], cfg["K"]
H, W = cfg["frame_hw"]
anchors, _ = bench.anchors_for(cfg, dev)
a3d = torch.zeros(anchors.shape[0], 7, device=dev)
prob, deltas, feat, info = bench.synth_frames(cfg, F, dev, 0)
gts = [bench.synth_gt(100 + f, cfg["G"], W, H, K).to(dev) for f in range(F)]
rois, scores, a3k, _, num = ops.proposal_batched(prob, deltas, info, anchors, a3d, A, cfg["pre_nms"], cfg["post_nms"], cfg["nms_thresh"], batch_index_stride=0)
fn = lambda: train_targets_batched(gts, info, anchors, A, cfg["Hf"], cfg["Wf"], rois, scores, a3k, num, None, K, 4)
for _ in range(3): fn()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): fn()
torch.cuda.synchronize()
print("targets per step (8 frames): %.3f ms wall" % ((time.perf_counter() - t0) / 10 * 1e3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
    for _ in range(5): fn()
    torch.cuda.synchronize()
print(p.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=50))
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=50))
