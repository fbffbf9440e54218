"""Per-kernel comparison with the stock sm_100 kernels the reference calls on a GPU (SURVEY.md §8d "the existing
sm_100 path to beat"): K1 = torchvision nms_kernel_impl (proposal_layer.py:46), K2 / K3 = torchvision
roi_align_forward / backward_kernel_impl (torchpoolers.py:165-170).  Same inputs, same B200, CUDA events, median of
`reps` runs after 3 warm-ups.  The boxes are what the proposal stage of the Waymo workload decodes; the RoIs what it
emits.

    python profiles/stock_kernels.py > profiles/r02_stock_kernels.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torchvision

import bench
from faster_rcnn_pytorch_multimodal_b200 import ops

dev = torch.device("cuda", 0)


def med(fn, reps=15):
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


def main():
    cfg = bench.WORKLOADS["waymo_test"]
    anchors, _ = bench.anchors_for(cfg, dev)
    out = {"clocks": None, "nms": {}, "roi_align_forward": {}, "roi_align_backward": {}}
    sampler = bench.ClockSampler(0)
    sampler.start()
    for F in (1, 32):
        prob, deltas, feat, info = bench.synth_frames(cfg, F, dev, 0)
        for tag, pre, post in (("test_6000_to_300", 6000, 300), ("train_12000_to_2000", 12000, 2000)):
            ops.proposal_batched(prob, deltas, info, anchors, None, cfg["A"], pre, post, 0.7)
            sb, ss, _ = ops.proposal_sorted_debug(F, cfg["Hf"] * cfg["Wf"], cfg["A"], pre, post, dev)
            sb, ss = sb.clone(), ss.clone()

            def stock():
                for f in range(F):
                    torchvision.ops.nms(sb[f], ss[f], 0.7)[:post]

            ours = lambda: ops.nms_sorted(sb, 0.7, max_keep=post)
            ours_unsorted = lambda: [ops.nms(sb[f], ss[f], 0.7) for f in range(F)]
            a, b = med(stock), med(ours)
            out["nms"][f"{tag}_F{F}"] = {"stock_K1_ms": a, "b2d_nms_sorted_ms": b, "speedup": a / b,
                                         "b2d_ops_nms_drop_in_ms": med(ours_unsorted) if F == 1 else None,
                                         "note": "stock: one torchvision.ops.nms call per frame (sorts again, full n x n/64 bitmask, "
                                                 "D2H of the mask + host sweep); ours: one launch for all frames, stops at post_nms"}
        # RoIAlign forward / backward
        for R in (300, 256, 2000):
            rois, _, _, _, num = ops.proposal_batched(prob, deltas, info, anchors, None, cfg["A"], 12000 if R == 2000 else 6000,
                                                      R, 0.7, batch_index_stride=1)
            flat = rois.view(-1, 5)
            fwd_stock = lambda: torchvision.ops.roi_align(feat, flat, (7, 7), 1.0 / 16, 2)
            fwd_ours = lambda: ops._roi_align_forward(feat, flat, (7, 7), 1.0 / 16, 2, False, seg_count=num, seg_stride=R)
            b = med(fwd_ours, 7)
            if F * R * cfg["C"] * 49 >= 2 ** 31:
                # the stock kernel indexes outputs with a 32-bit int: it returns early without covering them
                out["roi_align_forward"][f"R{R}_F{F}"] = {"stock_K2_ms": None, "b2d_ms": b, "speedup": None,
                                                          "note": "output has >= 2^31 elements: outside the stock kernel's range"}
                continue
            a = med(fwd_stock, 7)
            out["roi_align_forward"][f"R{R}_F{F}"] = {"stock_K2_ms": a, "b2d_ms": b, "speedup": a / b}
            if F * R <= 2000 * 4:
                g = torch.randn(F * R, cfg["C"], 7, 7, device=dev)
                fr = feat.clone().requires_grad_(True)
                o = torchvision.ops.roi_align(fr, flat, (7, 7), 1.0 / 16, 2)

                def bwd_stock():
                    fr.grad = None
                    o.backward(g, retain_graph=True)

                bwd_ours = lambda: ops._roi_align_backward(g, flat, tuple(feat.shape), (7, 7), 1.0 / 16, 2, False,
                                                           seg_count=num, seg_stride=R)
                a, b = med(bwd_stock, 7), med(bwd_ours, 7)
                out["roi_align_backward"][f"R{R}_F{F}"] = {"stock_K3_ms": a, "b2d_ms": b, "speedup": a / b,
                                                           "note": "stock: one atomicAdd per tap (non-deterministic); ours: "
                                                                   "atomic-free, bit-stable"}
                del g, fr, o
        del prob, deltas, feat
        torch.cuda.empty_cache()
    out["clocks"] = sampler.stop()
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
