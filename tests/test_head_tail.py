"""fused tail of the detection head over the MC-dropout stack.
CPU: the oracle restatement replays tests/golden/head_tail.npz (outputs of the unmodified reference functions).
GPU: `b2d_head_tail_decode` (one launch for all frames) against the same golden and against the oracle chain
feeding `b2d_final_detections`."""
import numpy as np
import pytest
import torch

from oracle import glue_oracle as O

KEYS = (("boxes", "boxes"), ("e_bbox_var", "e_var"), ("a_bbox_var", "a_var"), ("probs", "probs"),
        ("e_entropy", "ent"), ("e_mutual_info", "mi"))


def _close(a, b, tol=1e-5):
    a, b = torch.as_tensor(np.asarray(a)), torch.as_tensor(np.asarray(b))
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, rtol=tol, atol=tol), float((a - b).abs().max())


@pytest.mark.parametrize("tag", ["lidar", "image"])
def test_oracle_replays_reference_head_tail(golden, tag):
    g = golden("head_tail")
    bbox, cls, rois, a3d = (torch.from_numpy(g[f"{tag}_{k}"]) for k in ("bbox", "cls", "rois", "a3d"))
    a_in, info = torch.from_numpy(g[f"{tag}_a_var_in"]), g[f"{tag}_info"]
    for f in range(bbox.shape[1]):
        out = O.head_tail_decode(bbox[:, f], cls[:, f], rois[f], a3d[f], info[f], tag, a_bbox_var=a_in[f])
        for ours, theirs in KEYS:
            assert torch.equal(out[ours], torch.from_numpy(g[f"{tag}_{theirs}"][f])), (ours, f)
        sc = O.head_tail_decode(bbox[:, f], cls[:, f], rois[f], a3d[f], info[f], tag, use_scale=True)
        assert torch.equal(sc["boxes"], torch.from_numpy(g[f"{tag}_boxes_scaled"][f]))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["lidar", "image"])
def test_head_tail_decode_golden(golden, tag):
    from faster_rcnn_pytorch_multimodal_b200 import ops
    g = golden("head_tail")
    d = torch.device("cuda", 0)
    T_ = lambda k: torch.from_numpy(g[f"{tag}_{k}"]).to(d)
    out = ops.head_tail_decode(T_("bbox"), T_("cls"), T_("rois"), T_("a3d"), T_("info"), tag, a_bbox_var=T_("a_var_in"),
                               means=g[f"{tag}_means"].tolist(), stds=g[f"{tag}_stds"].tolist())
    for ours, theirs in KEYS:
        tol = 2e-5 if ours in ("boxes", "e_bbox_var", "a_bbox_var") else 1e-5
        a, b = out[ours].cpu(), torch.from_numpy(g[f"{tag}_{theirs}"])
        assert a.shape == b.shape
        assert torch.allclose(a, b, rtol=tol, atol=tol * max(1.0, float(b.abs().max()))), (ours, float((a - b).abs().max()))
    sc = ops.head_tail_decode(T_("bbox"), T_("cls"), T_("rois"), T_("a3d"), T_("info"), tag, use_scale=True)
    b = torch.from_numpy(g[f"{tag}_boxes_scaled"])
    assert torch.allclose(sc["boxes"].cpu(), b, rtol=2e-5, atol=2e-5 * float(b.abs().max()))
    # a single head pass (T = 1): plain decode, zero epistemic variance
    one = ops.head_tail_decode(T_("bbox")[:1], T_("cls")[:1], T_("rois"), T_("a3d"), T_("info"), tag)
    assert float(one["e_bbox_var"].abs().max()) == 0.0
    want = O.head_tail_decode(T_("bbox")[:1, 0].cpu(), T_("cls")[:1, 0].cpu(), T_("rois")[0].cpu(), T_("a3d")[0].cpu(),
                              g[f"{tag}_info"][0], tag)
    assert torch.allclose(one["boxes"][0].cpu(), want["boxes"], rtol=2e-5, atol=2e-3)


@pytest.mark.gpu
def test_head_tail_feeds_final_detections_like_the_oracle_chain():
    """BASELINE config 5 shape (T = 20, 300 RoIs, lidar K = 2): decode -> per-class filter, vs the oracle chain."""
    import bench
    from faster_rcnn_pytorch_multimodal_b200 import ops
    cfg = bench.WORKLOADS["mc_uncertainty"]
    d = torch.device("cuda", 0)
    F, R, K, E = 3, cfg["R"], cfg["K"], cfg["E"]
    bs, cs, rois, a3d, info = bench.synth_mc(cfg, F, d, 0)
    out = ops.head_tail_decode(bs, cs, rois, a3d, info, "lidar")
    dets, det_roi, counts, o_ur, o_uc = ops.final_detections(
        out["probs"], out["boxes"], info, E, "lidar", cfg["score_thresh"], cfg["nms_thresh"], cfg["max_dets"],
        uc_row=torch.stack((out["e_entropy"], out["e_mutual_info"]), 2), uc_cls=out["e_bbox_var"].view(F, R, 1, K * E),
        max_out=cfg["max_dets"])
    for f in range(F):
        boxes, evd, probs, ent, mi, odets, _ = bench.mc_oracle_frame(cfg, bs[:, f].cpu(), cs[:, f].cpu(), rois[f].cpu(),
                                                                     a3d[f].cpu(), info[f].cpu())
        assert torch.allclose(out["boxes"][f].cpu(), boxes, rtol=1e-4, atol=1e-3)
        assert torch.allclose(out["e_bbox_var"][f].cpu(), evd, rtol=1e-4, atol=1e-5)
        assert torch.allclose(out["probs"][f].cpu(), probs, rtol=1e-5, atol=1e-6)
        assert torch.allclose(out["e_mutual_info"][f].cpu(), mi, rtol=1e-4, atol=1e-5)
        for j in range(1, K):
            n = int(counts[f, j])
            assert n == len(odets[j]["dets"]) and n > 0
            assert np.allclose(dets[f, j, :n].cpu().numpy(), odets[j]["dets"], rtol=1e-4, atol=1e-3)
            assert np.array_equal(det_roi[f, j, :n].cpu().numpy(), odets[j]["roi"])
            assert np.allclose(o_ur[f, j, :n].cpu().numpy(), odets[j]["uc_row"], rtol=1e-4, atol=1e-5)
