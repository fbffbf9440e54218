"""Result files and detection matching against tests/golden/eval.npz, which holds the output of
the reference's own writers (datasets/db.py:305-367), `stack_uncertainties` (model/test.py:260-270),
`bbox_voxel_grid_to_pc` (utils/bbox.py:140-162) and its own `waymo_eval` loop (datasets/waymo_eval.py:44-250) run with the
missing utils/eval_utils.py bound to oracle/eval_oracle.py (oracle/gen_golden.py `gen_eval`)."""
import os
import pickle
import re

import numpy as np
import pytest
import torch

from oracle import eval_oracle as E

NETS = (("image", 4, "2d"), ("lidar", 7, "bev_aa"))


def load_case(g, nt):
    tokens = [str(t) for t in g["tokens"]]
    classes = [str(c) for c in g["classes"]]
    F, K = len(tokens), len(classes)
    all_boxes = [[np.empty(0) for _ in range(F)] for _ in range(K)]
    for c in range(1, K):
        for f in range(F):
            all_boxes[c][f] = g[f"{nt}_dets_{c}_{f}"]
    return tokens, classes, all_boxes


def class_recs(g, nt, tokens, c):
    """What load_recs (waymo_eval.py:266-310) builds for class c from the fixture's ground truth."""
    recs = []
    for f, tok in enumerate(tokens):
        key = f"{nt}_gt_{f}_boxes"
        if key not in g.files:
            recs.append({"ignore_frame": True, "filename": tok})
            continue
        boxes = g[key]
        if boxes.size == 0:
            recs.append({"ignore_frame": True, "filename": tok, "boxes": boxes})
            continue
        sel = np.where(g[f"{nt}_gt_{f}_gt_classes"] == c)[0]
        recs.append({"ignore_frame": False, "filename": tok, "boxes": boxes[sel], "boxes_dc": g[f"{nt}_gt_{f}_boxes_dc"],
                     "ignore": g[f"{nt}_gt_{f}_ignore"][sel], "difficulty": g[f"{nt}_gt_{f}_difficulty"][sel]})
    return recs


def results_of(golden_lines):
    """(confidence, fp flag, iou text) of every line waymo_eval.py handed to save_detection_results (write_det :325-395)."""
    out = []
    for l in golden_lines:
        m = re.search(r"confidence: (\S+) fp: (\d) .* iou: (\S+)$", str(l))
        out.append((float(m.group(1)), int(m.group(2)), m.group(3)))
    return out


# ------------------------------------------------------------------------------------------ CPU: oracle and host code
@pytest.mark.parametrize("nt,E_,eval_type", NETS)
def test_oracle_result_lines_match_the_reference_writer(golden, nt, E_, eval_type):
    g = golden("eval")
    tokens, classes, all_boxes = load_case(g, nt)
    for c in range(1, len(classes)):
        assert E.result_lines(all_boxes, c, tokens, lidar=nt == "lidar") == [str(x) for x in g[f"{nt}_lines_{c}"]]


@pytest.mark.parametrize("nt,E_,eval_type", NETS)
def test_result_writers_and_pickle_match_the_reference(golden, nt, E_, eval_type, tmp_path):
    from faster_rcnn_pytorch_multimodal_b200.datasets import results as R
    g = golden("eval")
    tokens, classes, all_boxes = load_case(g, nt)
    write = R.write_image_results_file if nt == "image" else R.write_lidar_results_file
    paths = write(all_boxes, classes, tokens, str(tmp_path), "test")
    assert [os.path.basename(p) for p in paths] == [f"det_test_{c}.txt" for c in classes[1:]]
    for c, p in zip(range(1, len(classes)), paths):
        assert open(p).readlines() == [str(x) for x in g[f"{nt}_lines_{c}"]]
    back = pickle.load(open(R.dump_detections(all_boxes, str(tmp_path)), "rb"))
    assert all(np.array_equal(a, b) for ra, rb in zip(back, all_boxes) for a, b in zip(ra, rb))


def test_stack_uncertainties_and_voxel_grid_oracle(golden):
    from faster_rcnn_pytorch_multimodal_b200.datasets import results as R
    g = golden("eval")
    ucs = {k: g["stack_uc_" + k] for k in ("a_bbox_var", "e_bbox_var", "a_entropy", "e_mutual_info")}
    assert np.array_equal(R.stack_uncertainties(g["stack_boxes"], ucs, 16), g["stack_hstack"])
    ext = [float(x) for x in g["vg_extents"]]
    assert np.array_equal(E.bbox_voxel_grid_to_pc(g["vg_in"].copy(), ext, g["vg_info"]), g["vg_out"])
    assert np.array_equal(E.bbox_voxel_grid_to_pc(g["vg_in"].copy(), ext, g["vg_info2"]), g["vg_out2"])
    assert np.array_equal(E.bbox_voxel_grid_to_pc(g["vg_in"][:, :4].copy(), ext, g["vg_info"], aabb=True), g["vg_out_aabb"])


@pytest.mark.parametrize("nt,E_,eval_type", NETS)
@pytest.mark.parametrize("ign_dc", [False, True])
def test_oracle_matching_reproduces_the_reference_loop(golden, nt, E_, eval_type, ign_dc):
    g = golden("eval")
    tokens, classes, _ = load_case(g, nt)
    for c in range(1, len(classes)):
        toks, conf, bb, _ = E.parse_result_lines([str(x) for x in g[f"{nt}_lines_{c}"]], E_)
        out = E.match_and_score(toks, conf, bb, class_recs(g, nt, tokens, c), 0.5, eval_type, 2, ignore_dc=ign_dc)
        tag = f"{nt}_eval_{c}_{int(ign_dc)}"
        assert np.array_equal(out["map"], g[tag + "_map"]) and np.array_equal(out["mrec"], g[tag + "_mrec"])
        want = results_of(g[tag + "_results"])
        got = [(conf[d], int(code != 1), "{:.3f}".format(ov)) for d, code, ov in zip(out["order"], out["code"], out["ovmax"])
               if code > 0]
        assert got == want
        assert (out["code"] == 1).sum() > 0 and (out["code"] == 3).sum() > 0 and (out["code"] == -1).sum() > 0


# ------------------------------------------------------------------------------------------ GPU: the device path
@pytest.mark.gpu
@pytest.mark.parametrize("nt,E_,eval_type", NETS)
@pytest.mark.parametrize("ign_dc", [False, True])
def test_device_matching_vs_reference_golden_and_oracle(golden, nt, E_, eval_type, ign_dc, tmp_path):
    from faster_rcnn_pytorch_multimodal_b200.datasets import results as R, waymo_eval as W
    from faster_rcnn_pytorch_multimodal_b200.model.config import cfg
    g = golden("eval")
    tokens, classes, all_boxes = load_case(g, nt)
    write = R.write_image_results_file if nt == "image" else R.write_lidar_results_file
    write(all_boxes, classes, tokens, str(tmp_path), "test")
    detpath = os.path.join(str(tmp_path), "results", "det_test_{:s}.txt")
    saved = cfg.NET_TYPE
    cfg.NET_TYPE = nt
    try:
        for c in range(1, len(classes)):
            recs = class_recs(g, nt, tokens, c)
            tag = f"{nt}_eval_{c}_{int(ign_dc)}"
            mrec, mprec, mp = W.waymo_eval(detpath, recs, classes[c], 0.5, eval_type, 2, ignore_dc=ign_dc)
            assert np.array_equal(mp, g[tag + "_map"]) and np.array_equal(mrec, g[tag + "_mrec"])       # the reference's values
            toks, conf, bb = W.parse_result_file(open(detpath.format(classes[c])).readlines(), E_)
            got = W.evaluate(toks, conf, bb, recs, 0.5, eval_type, 2, ignore_dc=ign_dc)
            want = E.match_and_score(toks, conf, bb, recs, 0.5, eval_type, 2, ignore_dc=ign_dc)
            assert np.array_equal(got["order"], want["order"]) and np.array_equal(got["code"], want["code"])
            ev = want["code"] >= 0
            assert np.array_equal(got["ovmax"][ev], want["ovmax"][ev]) and np.array_equal(got["jmax"][ev], want["jmax"][ev])
            m = got["tp"].shape[0]             # the reference sizes tp / fp by all detections and fills the evaluated ones
            assert m == int(ev.sum()) and not want["tp"][m:].any() and not want["fp"][m:].any()
            assert np.array_equal(got["tp"], want["tp"][:m]) and np.array_equal(got["fp"], want["fp"][:m])
    finally:
        cfg.NET_TYPE = saved


@pytest.mark.gpu
@pytest.mark.parametrize("eval_type,E_", [("2d", 4), ("bev_aa", 7)])
def test_device_matching_large_random_vs_oracle(eval_type, E_):
    """400 frames, up to 70 ground-truth boxes (more than one warp pass) and 90 detections per frame, duplicated
    detections, tied confidences, ignored boxes, frames without ground truth, unknown frames."""
    from faster_rcnn_pytorch_multimodal_b200.datasets import waymo_eval as W
    rng = np.random.RandomState(5)
    recs, toks, conf, bbs = [], [], [], []
    for f in range(400):
        tok = "f%05d" % f
        if f % 37 == 0:
            recs.append({"ignore_frame": True, "filename": tok})
        else:
            G = int(rng.randint(0, 71))
            if eval_type == "2d":
                xy = rng.rand(G, 2) * 1000
                boxes = np.concatenate((xy, xy + rng.rand(G, 2) * 120 + 10), 1).astype(np.float32)
            else:
                boxes = np.concatenate((rng.rand(G, 2) * 80, rng.rand(G, 1), rng.rand(G, 3) * 4 + 1, rng.rand(G, 1)), 1).astype(np.float32)
            recs.append({"ignore_frame": G == 0, "filename": tok, "boxes": boxes,
                         "boxes_dc": boxes[:G // 5] + np.float32(3.0), "ignore": rng.rand(G) < 0.15,
                         "difficulty": rng.randint(0, 4, G)})
            for j in range(G):
                for rep in range(int(rng.randint(0, 3))):
                    toks.append(tok)
                    bbs.append(boxes[j].astype(np.float64) + rng.randn(E_) * (3.0 if eval_type == "2d" else 0.1))
                    conf.append(round(float(rng.rand()), 2))                 # two decimals: many ties
    for _ in range(50):
        toks.append("unknown_frame")
        bbs.append(rng.rand(E_) * 50)
        conf.append(0.5)
    conf, bbs = np.array(conf), np.stack(bbs)
    for ign in (False, True):
        got = W.evaluate(toks, conf, bbs, recs, 0.5, eval_type, 2, ignore_dc=ign)
        want = E.match_and_score(toks, conf, bbs, recs, 0.5, eval_type, 2, ignore_dc=ign)
        assert np.array_equal(got["code"], want["code"])
        ev = want["code"] >= 0
        assert np.array_equal(got["ovmax"][ev], want["ovmax"][ev]) and np.array_equal(got["jmax"][ev], want["jmax"][ev])
        assert np.array_equal(got["ap"], want["map"])
        assert (want["code"] == 2).sum() > 10 and (want["code"] == 1).sum() > 100


@pytest.mark.gpu
def test_device_voxel_grid_to_pc_bit_exact(golden):
    from faster_rcnn_pytorch_multimodal_b200.datasets import results as R
    g = golden("eval")
    dev = torch.device("cuda", 0)
    ext = [float(x) for x in g["vg_extents"]]
    for info, want in ((g["vg_info"], g["vg_out"]), (g["vg_info2"], g["vg_out2"])):
        got = R.bbox_voxel_grid_to_pc(torch.from_numpy(g["vg_in"].copy()).to(dev), ext, info)
        assert np.array_equal(got.cpu().numpy(), want)
    got = R.bbox_voxel_grid_to_pc(torch.from_numpy(g["vg_in"][:, :4].copy()).to(dev), ext, g["vg_info"], aabb=True)
    assert np.array_equal(got.cpu().numpy(), g["vg_out_aabb"])
    # padded records [F, K, D, width]: every row in one launch
    rec = torch.from_numpy(np.tile(g["vg_in"], (2, 3, 1, 1)).copy()).to(dev)
    assert np.array_equal(R.bbox_voxel_grid_to_pc(rec, ext, g["vg_info"])[1, 2].cpu().numpy(), g["vg_out"])
