"""A/B timing of the RoIAlign forward launch on the bench workload (device time, CUDA events).
usage: python profiles/rows_ab.py [frames] [reps]; env knobs of csrc/roi_align_rows.cu are read per launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from faster_rcnn_pytorch_multimodal_b200 import ops
from faster_rcnn_pytorch_multimodal_b200.layer_utils.snippets import generate_anchors_pre

F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device('cuda', 0)
cfg = bench.CFG
anchors, _ = generate_anchors_pre(cfg["Hf"], cfg["Wf"], cfg["stride"], bench.SCALES, bench.RATIOS, 1.0, device=dev)
prob, deltas, feat, info = bench.synth_frames(cfg, F, dev, 0)
rois, scores, _, _, num = ops.proposal_batched(prob, deltas, info, anchors, None, cfg["A"], cfg["pre_nms"],
                                               cfg["post_nms"], cfg["nms_thresh"], batch_index_stride=1)
pooled = torch.empty(F * 300, 1024, 7, 7, device=dev)
r = rois.view(-1, 5)
h = (r[:, 4] - r[:, 2]) / 16
w = (r[:, 3] - r[:, 1]) / 16
qs = torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9], device=dev)
print("roi height (feature px) quantiles", torch.quantile(h, qs).tolist(), "width", torch.quantile(w, qs).tolist())


def run(tag, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    ts = []
    for it in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops._roi_align_forward(feat, r, (7, 7), 1.0 / 16, 2, False, seg_count=num, seg_stride=300, out=pooled)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    for k in (env or {}):
        del os.environ[k]
    ts.sort()
    print(f"{tag}: median {ts[len(ts)//2]*1e3/F:.2f} us/frame  min {ts[0]*1e3/F:.2f}")


run("default")
run("nostore", {"B2D_NOSTORE": "1"})
for spec in sys.argv[3:]:
    k, v = spec.split("=")
    run(spec, {k: v})
run("default again")
