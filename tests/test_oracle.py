"""The oracle restatement vs. the golden vectors produced by the real reference
(oracle/gen_golden.py) and vs. the reference's two in-tree known-answer vectors."""
import numpy as np
import pytest
import torch

from oracle import glue_oracle as O

FORK_SCALES = [2, 4, 8, 16, 32]
FORK_RATIOS = [0.5, 0.75, 1, 1.25, 2]
T = torch.from_numpy


def test_kat_nine_anchors(golden):
    # generate_anchors.py:20-28 lists the 1-based MATLAB table; the function returns it minus 1.
    want = np.array([[-84, -40, 99, 55], [-176, -88, 191, 103], [-360, -184, 375, 199],
                     [-56, -56, 71, 71], [-120, -120, 135, 135], [-248, -248, 263, 263],
                     [-36, -80, 51, 95], [-80, -168, 95, 183], [-168, -344, 183, 359]], dtype=np.float64)
    assert np.array_equal(O.generate_anchors(), want)
    assert np.array_equal(golden("anchors")["kat9"], want)


def test_anchor_grids(golden):
    g = golden("anchors")
    assert np.array_equal(O.generate_anchors(ratios=FORK_RATIOS, scales=FORK_SCALES), g["base25"])
    a, n = O.generate_anchors_pre(6, 8, 16, FORK_SCALES, FORK_RATIOS, 1.0)
    assert a.dtype == np.float32 and n == g["grid_len"] and np.array_equal(a, g["grid_6x8"])
    a, _ = O.generate_anchors_pre(5, 7, 16, FORK_SCALES, FORK_RATIOS, 0.5)
    assert np.array_equal(a, g["grid_5x7_s05"])


def test_3d_anchors_and_aabb(golden):
    g = golden("anchors")
    n, a3 = O.generate_3d_anchors(5, 4, 16, np.array([1]), np.array([0, np.pi / 2]), 1.0)
    assert n == g["a3d_n"] and np.array_equal(a3, g["a3d"])
    assert np.array_equal(O.bbaa_graphics_gems(a3.copy(), 64, 80, clip=False), g["a3d_aabb"])
    assert np.array_equal(O.bbaa_graphics_gems(a3.copy(), 64, 80, clip=True), g["a3d_aabb_clip"])
    # tools/bbox_rot_debug.py:7 boxes (SURVEY §8c probe values)
    want = np.array([[323.0762, 381.1941, 376.9238, 418.8059], [373.0762, 281.1941, 426.9238, 318.8059],
                     [73.0762, 81.1941, 126.9238, 118.8059]])
    got = O.bbaa_graphics_gems(g["dbg"].copy(), 700, 800)
    assert np.allclose(got, want, atol=1e-4) and np.array_equal(got, g["dbg_aabb"])
    rb = T(g["rot_boxes"])
    assert torch.equal(O.bbaa_graphics_gems_torch(rb.clone(), 700, 800, True), T(g["rot_aabb_t"]))
    assert torch.equal(O.bbaa_graphics_gems_torch(rb.clone(), 700, 800, False), T(g["rot_aabb_t_nc"]))


def test_codecs(golden):
    g = golden("codecs")
    ex, gt = T(g["ex"]), T(g["gt"])
    assert torch.equal(O.bbox_transform(ex, gt), T(g["enc"]))
    assert torch.equal(O.bbox_transform_inv(ex, T(g["d1"])), T(g["dec1"]))
    assert torch.equal(O.bbox_transform_inv(ex, T(g["d3"])), T(g["dec3"]))
    assert torch.equal(O.bbox_transform_inv(ex, T(g["d3"]), scales=1.5), T(g["dec3_s"]))
    assert torch.equal(O.clip_boxes(T(g["dec3"]) * 1.7 - 200, g["info"]), T(g["clip3"]))
    assert torch.equal(O.bbox_overlaps(ex, T(g["qb"])), T(g["iou"]))
    assert np.array_equal(O.bbox_overlaps(g["ex"].astype(np.float64), g["qb"].astype(np.float64)), g["iou_np64"])
    assert torch.equal(O.lidar_3d_bbox_transform(ex, T(g["a3d"]), T(g["gt7"])), T(g["l_enc"]))
    assert torch.equal(O.lidar_3d_bbox_transform_inv(ex, T(g["a3d"]).clone(), T(g["d7"])), T(g["l_dec"]))
    assert torch.equal(O.lidar_3d_uncertainty_transform_inv(ex, T(g["a3d"]).clone(), T(g["d7"]), T(g["uc7"])),
                       T(g["l_uc"]))
    assert O.bbox_transform_inv(torch.zeros(0, 4), torch.zeros(0, 4)).shape == (0, 4)


@pytest.mark.parametrize("key", ["TEST", "TRAIN"])
def test_proposal_layer(golden, key):
    g = golden("proposal")
    Hf, Wf, A = int(g["Hf"]), int(g["Wf"]), int(g["A"])
    anchors, _ = O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)
    a3 = torch.arange(anchors.shape[0] * 7, dtype=torch.float32).view(-1, 7)
    cfg = O.GlueCfg()
    pre, post = int(g[f"{key}_pre"]), int(g[f"{key}_post"])
    if key == "TEST":
        cfg.test_pre_nms, cfg.test_post_nms = pre, post
    else:
        cfg.train_pre_nms, cfg.train_post_nms = pre, post
    for stable in (True, False):            # scores are tie-free, so both orders agree
        blob, sc, a3k = O.proposal_layer(T(g[f"{key}_prob"]), T(g[f"{key}_deltas"]), g["info"], key,
                                         T(anchors), a3, A, cfg=cfg, stable_sort=stable)
        assert torch.equal(blob, T(g[f"{key}_blob"]))
        assert torch.equal(sc, T(g[f"{key}_scores"]))
        assert torch.equal(a3k[:, 0], T(g[f"{key}_a3d_col0"]))
    assert blob.shape[0] <= post and blob.shape[1] == 5


def test_proposal_top_layer(golden):
    g = golden("proposal")
    Hf, Wf, A = int(g["Hf"]), int(g["Wf"]), int(g["A"])
    anchors, _ = O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)
    cfg = O.GlueCfg(test_rpn_top_n=int(g["top_n"]))
    blob, sc, anc = O.proposal_top_layer(T(g["TRAIN_prob"]), T(g["TRAIN_deltas"]), g["info"], T(anchors), A, cfg=cfg)
    assert torch.equal(blob, T(g["top_blob"])) and torch.equal(sc, T(g["top_scores"]))
    assert torch.equal(anc, T(g["top_anchors"]))


def test_anchor_target_layer(golden):
    g = golden("anchor_target")
    Hf, Wf, A = int(g["Hf"]), int(g["Wf"]), int(g["A"])
    anchors, _ = O.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)
    torch.manual_seed(int(g["seed"]))
    lab, tg, iw, ow = O.anchor_target_layer(T(g["gt"]), torch.zeros(0, 5), g["info"], T(anchors), A, Hf, Wf)
    assert torch.equal(lab, T(g["labels"]))
    assert torch.equal(tg, T(g["targets"]))
    assert torch.equal(iw, T(g["inside_w"])) and torch.equal(ow, T(g["outside_w"]))
    assert (lab == 1).sum() <= 128 and (lab >= 0).sum() <= 256


@pytest.mark.parametrize("nt,E", [("image", 4), ("lidar", 7)])
def test_proposal_target_layer(golden, nt, E):
    g = golden("proposal_target")
    cfg = O.GlueCfg(net_type=nt)
    torch.manual_seed(int(g["seed"]))
    out = O.proposal_target_layer(T(g["rois"]), T(g["scores"]), T(g["a3d"]), T(g["gt"]), T(g["gt8"]),
                                  torch.zeros(0, 5), int(g["K"]), E, cfg=cfg, bg_mode="strict")
    for name, v in zip(("labels", "rois", "a3d", "scores", "targets", "inside_w", "outside_w"), out):
        assert torch.equal(v, T(g[f"{nt}_{name}"])), name
    # intended mode produces background rows; strict (reference-as-run) never does
    torch.manual_seed(int(g["seed"]))
    out_i = O.proposal_target_layer(T(g["rois"]), T(g["scores"]), T(g["a3d"]), T(g["gt"]), T(g["gt8"]),
                                    torch.zeros(0, 5), int(g["K"]), E, cfg=cfg, bg_mode="intended")
    assert (out[0] > 0).all() and (out_i[0] == 0).any() and out_i[0].shape[0] == 256


def test_thirdparty_nms(golden):
    g = golden("thirdparty")
    b, s = g["nms_boxes"], g["nms_scores"]
    for t in (0.3, 0.5, 0.6, 0.7):
        want = g[f"keep_{int(t * 100)}"]
        assert np.array_equal(O.nms_greedy_np(b, s, t), want)
        assert np.array_equal(O.nms(T(b), T(s), t).numpy(), want)
    assert O.nms_greedy_np(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 0.5).shape == (0,)


@pytest.mark.parametrize("sr", [2, 0, 1])
def test_thirdparty_roi_align(golden, sr):
    g = golden("thirdparty")
    feat, rois = g["feat"], g["rois"]
    out = O.roi_align_np(feat, rois, (7, 7), 1.0 / 16, sr, False)
    np.testing.assert_allclose(out, g[f"out_s{sr}"], rtol=1e-5, atol=1e-6)
    assert torch.equal(O.roi_align(T(feat), T(rois), (7, 7), 1.0 / 16, sr, False), T(g[f"out_s{sr}"]))
    gin = O.roi_align_backward_np(g[f"gout_s{sr}"], rois, feat.shape, 1.0 / 16, sr, False)
    np.testing.assert_allclose(gin, g[f"gin_s{sr}"], rtol=1e-5, atol=1e-5 * np.abs(g[f"gin_s{sr}"]).max())


def test_thirdparty_roi_align_aligned(golden):
    g = golden("thirdparty")
    out = O.roi_align_np(g["feat"], g["rois"], (7, 7), 1.0 / 16, 2, True)
    np.testing.assert_allclose(out, g["out_s2_aligned"], rtol=1e-5, atol=1e-6)


def test_fpn_mapper_and_multiscale(golden):
    g = golden("thirdparty")
    fb = T(g["fpn_boxes"])
    feats = [T(g[f"fpn_feat{i}"]) for i in range(4)]
    assert torch.equal(O.fpn_level_map(fb, 2, 5), T(g["fpn_levels"]))
    assert len(set(g["fpn_levels"].tolist())) == 4
    out = O.multiscale_roi_align(feats, fb, (512, 768), (7, 7), 2)
    assert torch.equal(out, T(g["fpn_out"]))


def test_uncertainty(golden):
    g = golden("uncertainty")
    assert torch.equal(O.compute_bbox_var(T(g["samples"])), T(g["var"]))
    assert torch.equal(O.categorical_mutual_information(T(g["logits"])), T(g["mutual_info"]))
    assert torch.equal(O.categorical_entropy(torch.softmax(T(g["logits"])[0], dim=1)), T(g["entropy"]))
    cov = O.compute_bbox_cov(T(g["samples"]).double())
    var_b = T(g["samples"]).double().var(dim=0, unbiased=False)
    assert torch.allclose(cov, var_b, rtol=1e-6, atol=1e-9)
    v = g["var"]
    order = O.sort_by_uncertainty(v, descending=True)
    assert np.all(np.diff(v.mean(1)[order]) <= 0)


# ------------------------------------------------------------------------------------------
# final per-class detection filter (utils/filter_predictions.py), pinned by the reference's own output
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,db_type,E", [("img", "image", 4), ("lid", "lidar", 7)])
def test_filter_detections_matches_reference(golden, tag, db_type, E):
    g = golden("detections")
    uc_row = torch.from_numpy(np.concatenate([g[f"{tag}_uc_a_entropy"][:, None], g[f"{tag}_uc_a_mutual_info"][:, None],
                                              g[f"{tag}_uc_a_cls_var"]], axis=1))
    uc_cls = torch.from_numpy(np.stack([g[f"{tag}_uc_a_bbox_var"], g[f"{tag}_uc_e_bbox_var"]], axis=1))
    out = O.filter_detections(torch.from_numpy(g[f"{tag}_probs"]), torch.from_numpy(g[f"{tag}_boxes"]),
                              g[f"{tag}_info"], 2, E, db_type, thresh=0.3, nms_thresh=0.6, uc_row=uc_row, uc_cls=uc_cls)
    d = out[1]
    assert np.array_equal(d["dets"], g[f"{tag}_dets"])
    assert np.array_equal(d["uc_row"][:, 0:1], g[f"{tag}_out_a_entropy"])
    assert np.array_equal(d["uc_row"][:, 1:2], g[f"{tag}_out_a_mutual_info"])
    assert np.array_equal(d["uc_row"][:, 2:], g[f"{tag}_out_a_cls_var"])
    assert np.array_equal(d["uc_cls"][:, 0], g[f"{tag}_out_a_bbox_var"])
    assert np.array_equal(d["uc_cls"][:, 1], g[f"{tag}_out_e_bbox_var"])


def test_nms_hstack_matches_reference(golden):
    g = golden("detections")
    for c in range(1, 4):
        dets, inds, keep = O.nms_hstack(torch.from_numpy(g["k4_probs"]), torch.from_numpy(g["k4_boxes"]), 0.1, c, 4,
                                        "image", 0.6)
        assert np.array_equal(dets, g[f"k4_dets{c}"])
        assert np.array_equal(np.asarray(inds), g[f"k4_inds{c}"]) and np.array_equal(keep, g[f"k4_keep{c}"])
    # max-dets filter (model/test.py:213-221) on the same rows
    out = O.filter_detections(torch.from_numpy(g["k4_probs"]), torch.from_numpy(g["k4_boxes"]),
                              np.array([0, 1e9, 0, 1e9, 0, 0, 1], np.float32), 4, 4, "image",
                              thresh=0.1, nms_thresh=0.6, max_dets=5)
    for c in range(1, 4):
        assert len(out[c]["dets"]) == min(5, len(g[f"k4_dets{c}"]))
