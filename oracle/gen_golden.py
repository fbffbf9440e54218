"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions.

Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden

Every fixture stores the inputs and the reference's outputs, so the tests can
replay them through (a) the oracle restatement on CPU and (b) the CUDA path on
the GPU box, where /root/reference does not exist.  Seeds follow the reference's
``cfg.RNG_SEED = 3`` (model/config.py:346).
"""
import os
import sys

import numpy as np
import torch

from . import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 3
FORK_SCALES = [2, 4, 8, 16, 32]             # config.py:373
FORK_RATIOS = [0.5, 0.75, 1, 1.25, 2]       # config.py:378


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _detie(scores_flat):
    """Make positive fg scores tie-free: bump duplicates by whole ulps."""
    bits = scores_flat.clone().view(torch.int32)
    for _ in range(64):
        order = torch.argsort(bits, stable=True)
        sb = bits[order]
        dup = torch.zeros_like(sb, dtype=torch.bool)
        dup[1:] = sb[1:] == sb[:-1]
        if not dup.any():
            break
        bits[order[dup]] += 1
    out = bits.view(torch.float32)
    assert torch.unique(out).numel() == out.numel()
    return out


def synth_rpn(g, Hf, Wf, A):
    """Synthetic RPN outputs per SURVEY §8d: softmax pair scores, small deltas."""
    logits = torch.randn(1, Hf, Wf, 2 * A, generator=g)
    pair = torch.stack((logits[..., :A], logits[..., A:]), dim=-1).softmax(-1)
    prob = torch.cat((pair[..., 0], pair[..., 1]), dim=-1).contiguous()
    fg = _detie(prob[..., A:].contiguous().view(-1))
    prob[..., A:] = fg.view(1, Hf, Wf, A)
    d = torch.randn(1, Hf, Wf, A, 4, generator=g)
    d[..., :2] *= 0.1
    d[..., 2:] *= 0.2
    return prob, d.reshape(1, Hf, Wf, 4 * A).contiguous()


def synth_gt_image(g, G, W, H, K):
    wh = torch.exp(torch.rand(G, 2, generator=g) * (np.log(400.0) - np.log(16.0)) + np.log(16.0))
    wh[:, 0].clamp_(max=W - 2)
    wh[:, 1].clamp_(max=H - 2)
    x1 = torch.rand(G, generator=g) * (W - 1 - wh[:, 0])
    y1 = torch.rand(G, generator=g) * (H - 1 - wh[:, 1])
    cls = torch.randint(1, K, (G,), generator=g).float()
    return torch.stack((x1, y1, x1 + wh[:, 0], y1 + wh[:, 1], cls), dim=1)


def main():
    ref = ref_import.load()
    cfg = ref.cfg
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)

    # ---- anchors ------------------------------------------------------------------
    kat9 = ref.ga.generate_anchors()
    base25 = ref.ga.generate_anchors(ratios=np.array(FORK_RATIOS), scales=np.array(FORK_SCALES))
    a_grid, a_len = ref.sn.generate_anchors_pre(6, 8, 16, FORK_SCALES, FORK_RATIOS, 1.0)
    a_grid_s, _ = ref.sn.generate_anchors_pre(5, 7, 16, FORK_SCALES, FORK_RATIOS, 0.5)
    cfg.NET_TYPE = "lidar"
    n3, a3 = ref.g3.GridAnchor3dGenerator()._generate(5, 4, 16, cfg.LIDAR.ANCHOR_SCALES[0],
                                                       cfg.LIDAR.ANCHOR_ANGLES, 1.0)
    a3_aabb = ref.ub.bbaa_graphics_gems(a3.copy(), 4 * 16, 5 * 16, clip=False)
    a3_aabb_clip = ref.ub.bbaa_graphics_gems(a3.copy(), 4 * 16, 5 * 16, clip=True)
    dbg = np.array([[350, 400, 2, 50, 20, 3, np.pi / 8],            # tools/bbox_rot_debug.py:7
                    [400, 300, 2, 50, 20, 3, np.pi / 8],
                    [100, 100, 2, 50, 20, 3, np.pi / 8]], dtype=np.float64)
    dbg_aabb = ref.ub.bbaa_graphics_gems(dbg.copy(), 700, 800)
    g = torch.Generator().manual_seed(SEED)
    rot_boxes = torch.rand(40, 7, generator=g) * torch.tensor([700, 800, 4, 60, 30, 3, np.pi]) \
        - torch.tensor([0, 0, 0, -5, -5, -1, np.pi / 2])
    rot_aabb_t = ref.ub.bbaa_graphics_gems_torch(rot_boxes.clone(), 700, 800, clip=True)
    rot_aabb_t_nc = ref.ub.bbaa_graphics_gems_torch(rot_boxes.clone(), 700, 800, clip=False)
    np.savez_compressed(os.path.join(OUT, "anchors.npz"), kat9=kat9, base25=base25, grid_6x8=a_grid,
                        grid_len=a_len, grid_5x7_s05=a_grid_s, a3d_n=n3, a3d=a3, a3d_aabb=a3_aabb,
                        a3d_aabb_clip=a3_aabb_clip, dbg=dbg, dbg_aabb=dbg_aabb, rot_boxes=_np(rot_boxes),
                        rot_aabb_t=_np(rot_aabb_t), rot_aabb_t_nc=_np(rot_aabb_t_nc))

    # ---- codecs + IoU ---------------------------------------------------------------
    g = torch.Generator().manual_seed(SEED + 1)
    ex = synth_gt_image(g, 64, 1920, 1280, 4)[:, :4]
    gt = synth_gt_image(g, 64, 1920, 1280, 4)[:, :4]
    enc = ref.bt.bbox_transform(ex, gt)
    d1 = torch.randn(64, 4, generator=g) * torch.tensor([0.1, 0.1, 0.2, 0.2])
    d3 = torch.randn(64, 12, generator=g) * 0.2
    dec1 = ref.bt.bbox_transform_inv(ex, d1)
    dec3 = ref.bt.bbox_transform_inv(ex, d3)
    dec3_s = ref.bt.bbox_transform_inv(ex, d3, scales=1.5)
    info_img = np.array([0, 1920, 0, 1280, 0, 0, 1.0], dtype=np.float32)
    clip3 = ref.bt.clip_boxes(dec3 * 1.7 - 200, info_img)
    qb = synth_gt_image(g, 9, 1920, 1280, 4)[:, :4]
    iou = ref.ub.bbox_overlaps(ex, qb)
    iou_np = ref.ub.bbox_overlaps(_np(ex).astype(np.float64), _np(qb).astype(np.float64))
    a3d_e = torch.rand(64, 7, generator=g) * torch.tensor([700, 800, 4, 60, 30, 3, 3.0]) + \
        torch.tensor([0, 0, 0.5, 5, 5, 1, -1.5])
    gt7 = torch.rand(64, 7, generator=g) * torch.tensor([700, 800, 4, 60, 30, 3, 3.0]) + \
        torch.tensor([0, 0, 0.5, 5, 5, 1, -1.5])
    l_enc = ref.bt.lidar_3d_bbox_transform(ex, a3d_e, gt7)
    d7 = torch.randn(64, 14, generator=g) * 0.2
    l_dec = ref.bt.lidar_3d_bbox_transform_inv(ex, a3d_e.clone(), d7)
    uc7 = torch.rand(64, 14, generator=g) * 0.3
    l_uc = ref.bt.lidar_3d_uncertainty_transform_inv(ex, a3d_e.clone(), d7, uc7)
    np.savez_compressed(os.path.join(OUT, "codecs.npz"), ex=_np(ex), gt=_np(gt), enc=_np(enc), d1=_np(d1),
                        d3=_np(d3), dec1=_np(dec1), dec3=_np(dec3), dec3_s=_np(dec3_s), info=info_img,
                        clip3=_np(clip3), qb=_np(qb), iou=_np(iou), iou_np64=iou_np, a3d=_np(a3d_e),
                        gt7=_np(gt7), l_enc=_np(l_enc), d7=_np(d7), l_dec=_np(l_dec), uc7=_np(uc7),
                        l_uc=_np(l_uc))

    # ---- proposal_layer -------------------------------------------------------------
    Hf, Wf, A = 20, 30, 25                     # 320x480 frame, N = 15000
    info = np.array([0, Wf * 16, 0, Hf * 16, 0, 0, 1.0], dtype=np.float32)
    anchors, _ = ref.sn.generate_anchors_pre(Hf, Wf, 16, FORK_SCALES, FORK_RATIOS, 1.0)
    anchors_t = torch.from_numpy(anchors)
    a3_dummy = torch.arange(anchors.shape[0] * 7, dtype=torch.float32).view(-1, 7)
    cfg.NET_TYPE = "image"
    prop = {}
    for key, pre, post in (("TEST", 3000, 150), ("TRAIN", 6000, 800)):
        g = torch.Generator().manual_seed(SEED + 10 + len(prop))
        prob, deltas = synth_rpn(g, Hf, Wf, A)
        cfg[key].RPN_PRE_NMS_TOP_N = pre
        cfg[key].RPN_POST_NMS_TOP_N = post
        blob, sc, a3k = ref.pl.proposal_layer(prob, deltas, info, key, anchors_t, a3_dummy, A)
        prop.update({f"{key}_prob": _np(prob), f"{key}_deltas": _np(deltas), f"{key}_pre": pre,
                     f"{key}_post": post, f"{key}_blob": _np(blob), f"{key}_scores": _np(sc),
                     f"{key}_a3d_col0": _np(a3k[:, 0])})
    cfg.TEST.RPN_TOP_N = 500
    tblob, tsc, tanc = ref.ptl.proposal_top_layer(prob, deltas, info, anchors_t, A)
    prop.update(top_blob=_np(tblob), top_scores=_np(tsc), top_anchors=_np(tanc), top_n=500)
    cfg.TEST.RPN_TOP_N = 5000
    cfg.TRAIN.RPN_PRE_NMS_TOP_N, cfg.TRAIN.RPN_POST_NMS_TOP_N = 12000, 2000
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N = 6000, 300
    np.savez_compressed(os.path.join(OUT, "proposal.npz"), Hf=Hf, Wf=Wf, A=A, info=info, **prop)

    # ---- anchor_target_layer_torch ---------------------------------------------------
    g = torch.Generator().manual_seed(SEED + 20)
    gtb = synth_gt_image(g, 12, Wf * 16, Hf * 16, 4)
    gtb[:, 2:4] = torch.minimum(gtb[:, 2:4], torch.tensor([Wf * 16 - 1.0, Hf * 16 - 1.0]))
    torch.manual_seed(SEED)
    lab, tg, iw, ow = ref.atl.anchor_target_layer_torch(gtb, torch.zeros(0, 5), info, anchors_t, A, Hf, Wf,
                                                         torch.device("cpu"))
    np.savez_compressed(os.path.join(OUT, "anchor_target.npz"), Hf=Hf, Wf=Wf, A=A, info=info, gt=_np(gtb),
                        seed=SEED, labels=_np(lab), targets=_np(tg), inside_w=_np(iw), outside_w=_np(ow))

    # ---- proposal_target_layer (image strict; lidar strict) -----------------------------
    g = torch.Generator().manual_seed(SEED + 30)
    R = 400
    jitter = torch.randn(R, 4, generator=g) * 6.0
    rois = gtb[torch.randint(0, gtb.shape[0], (R,), generator=g), :4] + jitter
    rois = torch.cat((rois[:200], synth_gt_image(g, 200, Wf * 16, Hf * 16, 4)[:, :4]), 0)
    rois = torch.cat((torch.zeros(R, 1), rois), 1)
    rsc = torch.rand(R, 1, generator=g)
    a3r = torch.rand(R, 7, generator=g) * torch.tensor([480, 320, 4, 60, 30, 3, 3.0]) + \
        torch.tensor([0, 0, 0.5, 5, 5, 1, -1.5])
    tgt8 = torch.cat((torch.rand(gtb.shape[0], 7, generator=g) * torch.tensor([480, 320, 4, 60, 30, 3, 3.0])
                      + torch.tensor([0, 0, 0.5, 5, 5, 1, -1.5]), gtb[:, 4:5]), 1)
    pt = {}
    for nt, E in (("image", 4), ("lidar", 7)):
        cfg.NET_TYPE = nt
        torch.manual_seed(SEED)
        o = ref.prt.proposal_target_layer(rois, rsc, a3r, gtb, tgt8, torch.zeros(0, 5), 4, E)
        for name, v in zip(("labels", "rois", "a3d", "scores", "targets", "inside_w", "outside_w"), o):
            pt[f"{nt}_{name}"] = _np(v)
    cfg.NET_TYPE = "image"
    np.savez_compressed(os.path.join(OUT, "proposal_target.npz"), rois=_np(rois), scores=_np(rsc), a3d=_np(a3r),
                        gt=_np(gtb), gt8=_np(tgt8), K=4, seed=SEED, **pt)

    # ---- third-party kernels: NMS, RoIAlign fwd/bwd, FPN mapper ---------------------------
    from torchvision.ops import nms as tv_nms, roi_align as tv_roi_align
    g = torch.Generator().manual_seed(SEED + 40)
    centers = torch.rand(60, 2, generator=g) * torch.tensor([400.0, 300.0])
    nb = centers[torch.randint(0, 60, (1500,), generator=g)] + torch.randn(1500, 2, generator=g) * 6
    nwh = torch.rand(1500, 2, generator=g) * 60 + 10
    nboxes = torch.cat((nb - nwh / 2, nb + nwh / 2), 1)
    nboxes[::97, 2:] = nboxes[::97, :2]                    # zero-area boxes (NaN IoU) survive
    nscores = torch.rand(1500, generator=g).sort(descending=True)[0]
    nms_out = {f"keep_{int(t * 100)}": _np(tv_nms(nboxes, nscores, t)) for t in (0.3, 0.5, 0.6, 0.7)}
    C, H, W = 6, 24, 30
    feat = torch.randn(1, C, H, W, generator=g)
    rr = torch.rand(40, 4, generator=g) * torch.tensor([W * 16.0, H * 16.0, W * 8.0, H * 8.0])
    rr = torch.stack((rr[:, 0], rr[:, 1], rr[:, 0] + rr[:, 2] + 1, rr[:, 1] + rr[:, 3] + 1), 1)
    rr[0] = torch.tensor([-40.0, -30.0, 20.0, 25.0])         # partly outside
    rr[1] = torch.tensor([W * 16 - 10.0, H * 16 - 10.0, W * 16 + 50.0, H * 16 + 40.0])
    rr[2] = torch.tensor([100.0, 100.0, 100.0, 100.0])        # degenerate
    rr[3] = torch.tensor([0.0, 0.0, W * 16.0 - 1, H * 16.0 - 1])  # whole frame
    rrois = torch.cat((torch.zeros(40, 1), rr), 1)
    ra = {}
    for sr in (2, 0, 1):
        f_ = feat.clone().requires_grad_(True)
        o = tv_roi_align(f_, rrois, (7, 7), 1.0 / 16, sr, False)
        go = torch.randn(o.shape, generator=torch.Generator().manual_seed(SEED + 41))
        o.backward(go)
        ra[f"out_s{sr}"] = _np(o)
        ra[f"gin_s{sr}"] = _np(f_.grad)
        ra[f"gout_s{sr}"] = _np(go)
    ra["out_s2_aligned"] = _np(tv_roi_align(feat, rrois, (7, 7), 1.0 / 16, 2, True))
    # FPN
    feats = [torch.randn(1, 2, 128 // s, 192 // s, generator=g) for s in (1, 2, 4, 8)]   # strides 4..32 of 512x768
    fsz = torch.exp(torch.rand(50, 2, generator=g) * (np.log(700.0) - np.log(20.0)) + np.log(20.0))
    fsz[:, 1].clamp_(max=500.0)
    fxy = torch.rand(50, 2, generator=g) * (torch.tensor([767.0, 511.0]) - fsz)
    fb = torch.cat((fxy, fxy + fsz), 1)
    from collections import OrderedDict
    msra = ref.tp.MultiScaleRoIAlign(["p2", "p3", "p4", "p5"], 7, 2)
    od = OrderedDict((n, ft) for n, ft in zip(["p2", "p3", "p4", "p5"], feats))
    fout = msra(od, [fb], [(512, 768)])
    flv = msra.map_levels([fb])
    np.savez_compressed(os.path.join(OUT, "thirdparty.npz"), nms_boxes=_np(nboxes), nms_scores=_np(nscores),
                        feat=_np(feat), rois=_np(rrois), fpn_boxes=_np(fb), fpn_out=_np(fout),
                        fpn_levels=_np(flv), fpn_scales=np.array(msra.scales),
                        **{f"fpn_feat{i}": _np(ft) for i, ft in enumerate(feats)}, **nms_out, **ra)

    # ---- MC-dropout reductions -----------------------------------------------------------
    g = torch.Generator().manual_seed(SEED + 50)
    mu = torch.randn(1, 60, 14, generator=g) * 30
    samp = mu + torch.randn(20, 60, 14, generator=g) * 0.05
    var = ref.lu.compute_bbox_var(samp)
    logits = torch.randn(20, 60, 4, generator=g)
    import torch.nn.functional as F
    ref.lu.F = F
    mi = ref.lu.categorical_mutual_information(logits)
    ent = ref.lu.categorical_entropy(torch.softmax(logits[0], dim=1))
    np.savez_compressed(os.path.join(OUT, "uncertainty.npz"), samples=_np(samp), var=_np(var),
                        logits=_np(logits), mutual_info=_np(mi), entropy=_np(ent))
    # ---- final per-class detection filter (utils/filter_predictions.py) -----------------------
    # The reference's own filter_and_draw_prep, K = 2 (Waymo vehicles; its uncertainty bookkeeping is
    # only well-defined for one foreground class), image and lidar, with every UC flag on.
    det = {}
    g = torch.Generator().manual_seed(SEED + 60)
    for tag, db_type, E in (("img", "image", 4), ("lid", "lidar", 7)):
        cfg.NET_TYPE = db_type
        R, K = 120, 2
        for flag in ("EN_CLS_ALEATORIC", "EN_CLS_EPISTEMIC", "EN_BBOX_ALEATORIC", "EN_BBOX_EPISTEMIC"):
            cfg.UC[flag] = True
        probs = torch.softmax(torch.randn(R, K, generator=g) * 2.0, dim=1)
        if db_type == "image":
            ctr = torch.rand(R, 2, generator=g) * torch.tensor([1920.0, 1280.0])
            ctr[R // 2:] = ctr[:R // 2] + torch.randn(R - R // 2, 2, generator=g) * 12      # overlapping clusters
            wh = torch.exp(torch.rand(R, 2, generator=g) * 2.0 + 3.5)
            one = torch.cat((ctr - wh / 2, ctr + wh / 2), dim=1) + torch.tensor([-30.0, -30.0, 40.0, 40.0]) * 0
            one[:8] += torch.tensor([-200.0, -200.0, 300.0, 300.0])                        # some outside the frame
            info = np.array([0, 1920, 0, 1280, 0, 0, 1.0], dtype=np.float32)
        else:
            ctr = torch.rand(R, 2, generator=g) * torch.tensor([700.0, 800.0])
            ctr[R // 2:] = ctr[:R // 2] + torch.randn(R - R // 2, 2, generator=g) * 6
            size = torch.tensor([47.3, 20.8, 1.77]) * (0.8 + 0.4 * torch.rand(R, 3, generator=g))
            one = torch.cat((ctr, torch.rand(R, 1, generator=g) * 4, size, torch.rand(R, 1, generator=g) * 3 - 1.5), 1)
            info = np.array([0, 700, 0, 800, 0, 12, 1.0], dtype=np.float32)
        boxes = torch.cat((torch.zeros(R, E), one), dim=1)                                 # class 0 slot unused
        uc = {"a_entropy": torch.rand(R, generator=g), "a_mutual_info": torch.rand(R, generator=g),
              "a_cls_var": torch.rand(R, K, generator=g), "e_entropy": torch.rand(R, generator=g),
              "e_mutual_info": torch.rand(R, generator=g), "e_cls_var": torch.rand(R, K, generator=g),
              "a_bbox_var": torch.rand(R, K * E, generator=g), "e_bbox_var": torch.rand(R, K * E, generator=g)}
        rois = torch.cat((torch.zeros(R, 1), torch.rand(R, 4, generator=g) * 500), dim=1)
        r_np, all_boxes, all_uc = ref.fp.filter_and_draw_prep(rois, probs.clone(), boxes.clone(),
                                                              {k: v.clone() for k, v in uc.items()}, info, K, 0.3,
                                                              db_type)
        det.update({f"{tag}_probs": _np(probs), f"{tag}_boxes": _np(boxes), f"{tag}_info": info,
                    f"{tag}_rois": _np(rois), f"{tag}_out_rois": r_np, f"{tag}_dets": all_boxes[1]})
        for k, v in uc.items():
            det[f"{tag}_uc_{k}"] = _np(v)
            det[f"{tag}_out_{k}"] = np.asarray(all_uc[1][k])
        for flag in ("EN_CLS_ALEATORIC", "EN_CLS_EPISTEMIC", "EN_BBOX_ALEATORIC", "EN_BBOX_EPISTEMIC"):
            cfg.UC[flag] = False
    # the per-class call on its own, K = 4, thresholds as in test.py (0.1) -> dets / inds / keep per class
    cfg.NET_TYPE = "image"
    probs4 = torch.softmax(torch.randn(200, 4, generator=g) * 1.5, dim=1)
    ctr = torch.rand(200, 2, generator=g) * torch.tensor([1242.0, 375.0])
    ctr[100:] = ctr[:100] + torch.randn(100, 2, generator=g) * 8
    wh = torch.exp(torch.rand(200, 2, generator=g) * 1.5 + 3.0)
    b4 = torch.cat([torch.cat((ctr - wh / 2, ctr + wh / 2), dim=1) + torch.randn(200, 4, generator=g) * 3
                    for _ in range(4)], dim=1)
    det.update(k4_probs=_np(probs4), k4_boxes=_np(b4))
    for c in range(1, 4):
        d, inds, keep = ref.fp.nms_hstack_torch(probs4, b4, 0.1, c, 4, "image")
        det[f"k4_dets{c}"], det[f"k4_inds{c}"], det[f"k4_keep{c}"] = d, _np(inds), np.asarray(keep)
    np.savez_compressed(os.path.join(OUT, "detections.npz"), **det)

    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


def gen_bev():
    """tests/golden/bev.npz: the reference's own `_get_lidar_blob` (roi_data_layer/minibatch.py:237-512, test mode,
    no augmentation) on synthetic sweeps.  `spconv` is not installable here: the voxeliser it calls is the
    restatement in oracle/bev_oracle.py (see oracle/ref_import.load_minibatch), everything after it is the
    reference's code.  Small ranges keep the fixture small; the caps are set low so that both bind."""
    import tempfile
    from . import bev_oracle as B
    mb = ref_import.load_minibatch()
    from model.config import cfg
    out = {}
    cases = {
        # name: (x_range, y_range, max_pts, max_voxels, n_points, nfeat, db_name)
        "caps": ((0, 8), (-4, 4), 4, 300, 6000, 5, "waymo"),
        "roomy": ((0, 12), (-6, 6), 32, 25000, 20000, 5, "waymo"),
        "kitti4": ((0, 8), (-4, 4), 32, 25000, 5000, 4, "kitti"),
    }
    saved = (cfg.LIDAR.X_RANGE, cfg.LIDAR.Y_RANGE, cfg.LIDAR.MAX_PTS_PER_VOXEL, cfg.LIDAR.MAX_NUM_VOXEL, cfg.DB_NAME,
             cfg.LIDAR.NUM_META_CHANNEL, cfg.LIDAR.NUM_CHANNEL)
    tmp = tempfile.mkdtemp()
    for i, (name, (xr, yr, mp, mv, n, nf, db)) in enumerate(cases.items()):
        cfg.LIDAR.X_RANGE, cfg.LIDAR.Y_RANGE = list(xr), list(yr)
        cfg.LIDAR.MAX_PTS_PER_VOXEL, cfg.LIDAR.MAX_NUM_VOXEL = mp, mv
        # the fork's KITTI sweeps carry no elongation: it runs them with the '.npy' loader and 2 meta channels
        cfg.DB_NAME = "waymo" if db == "waymo" else "nuscenes"
        cfg.LIDAR.NUM_META_CHANNEL = 3 if nf == 5 else 2
        cfg.LIDAR.NUM_CHANNEL = cfg.LIDAR.NUM_SLICES + cfg.LIDAR.NUM_META_CHANNEL
        lc = B.LidarCfg(x_range=xr, y_range=yr, max_pts_per_voxel=mp, max_num_voxel=mv, db_name=cfg.DB_NAME,
                        num_meta_channel=cfg.LIDAR.NUM_META_CHANNEL)
        pts = B.synth_point_cloud(SEED + i, n, lc, nfeat=nf)
        f = os.path.join(tmp, f"{name}.npy")
        np.save(f, pts)
        ext = [xr[0], yr[0], cfg.LIDAR.Z_RANGE[0], xr[1], yr[1], cfg.LIDAR.Z_RANGE[1]]
        infos, blob, _ = mb._get_lidar_blob([f], ext, 1.0, augment_en=False, mode="test")
        out[f"{name}_points"] = pts
        out[f"{name}_cfg"] = np.array([xr[0], xr[1], yr[0], yr[1], mp, mv, cfg.LIDAR.NUM_META_CHANNEL, db == "waymo"], dtype=np.float64)
        out[f"{name}_info"] = np.asarray(infos[0], dtype=np.float64)
        # sparse form of the [num_y, num_x, C] map: flat indices + values (the map is > 95 % zeros)
        nz = np.flatnonzero(blob[0])
        out[f"{name}_shape"] = np.asarray(blob[0].shape)
        out[f"{name}_nz_idx"] = nz.astype(np.int64)
        out[f"{name}_nz_val"] = blob[0].ravel()[nz]
    (cfg.LIDAR.X_RANGE, cfg.LIDAR.Y_RANGE, cfg.LIDAR.MAX_PTS_PER_VOXEL, cfg.LIDAR.MAX_NUM_VOXEL, cfg.DB_NAME,
     cfg.LIDAR.NUM_META_CHANNEL, cfg.LIDAR.NUM_CHANNEL) = saved
    np.savez_compressed(os.path.join(OUT, "bev.npz"), **out)
    print("bev.npz", os.path.getsize(os.path.join(OUT, "bev.npz")))


def gen_head_tail():
    """tests/golden/head_tail.npz: the tail of the detection head over T MC-dropout passes,
    composed from the UNMODIFIED reference functions - de-normalisation with cfg.TRAIN.*.BBOX_NORMALIZE_* (config.py:
    219-223), torch.mean, compute_bbox_var, lidar_3d_bbox_transform_inv / bbox_transform_inv + clip_boxes,
    lidar_3d_uncertainty_transform_inv, softmax mean, categorical_entropy, categorical_mutual_information.
    (The method that strings them together, Network.test_frame, is in the missing lib/nets/network.py.)"""
    R = ref_import.load()
    g = torch.Generator().manual_seed(SEED + 40)
    out = {}
    T, F, n, K = 6, 2, 37, 3
    for tag, E in (("lidar", 7), ("image", 4)):
        tr = R.cfg.TRAIN.LIDAR if tag == "lidar" else R.cfg.TRAIN.IMAGE
        stds = torch.tensor(tr.BBOX_NORMALIZE_STDS, dtype=torch.float32).repeat(K)
        means = torch.tensor(tr.BBOX_NORMALIZE_MEANS, dtype=torch.float32).repeat(K)
        mu = torch.randn(F, n, K * E, generator=g)
        bbox = mu.unsqueeze(0) + 0.1 * torch.randn(T, F, n, K * E, generator=g)
        cls = 2.0 * torch.randn(F, n, K, generator=g).unsqueeze(0) + 0.4 * torch.randn(T, F, n, K, generator=g)
        xy = torch.rand(F, n, 2, generator=g) * torch.tensor([600.0, 700.0])
        wh = torch.rand(F, n, 2, generator=g) * 70 + 6
        rois = torch.cat((torch.arange(F).view(F, 1, 1).expand(F, n, 1).float(), xy, xy + wh), 2)
        a3d = torch.tensor([0, 0, 0.885, 47.3, 20.8, 1.77, 0.0]).repeat(F, n, 1) + 0.05 * torch.randn(F, n, 7, generator=g)
        a_var = 0.02 * torch.rand(F, n, K * E, generator=g)
        info = torch.tensor([[0, 700.0, 0, 800.0, 0, 12.0, 1.0], [0, 700.0, 0, 800.0, 0, 12.0, 1.25]])
        res = {k: [] for k in ("boxes", "boxes_scaled", "e_var", "a_var", "probs", "ent", "mi")}
        for f in range(F):
            x = bbox[:, f] * stds + means                                   # config.py:219-223
            mean_pred = torch.mean(x, dim=0)
            var = R.lu.compute_bbox_var(x)
            if tag == "lidar":
                res["boxes"].append(R.bt.lidar_3d_bbox_transform_inv(rois[f, :, 1:5], a3d[f].clone(), mean_pred))
                res["boxes_scaled"].append(R.bt.lidar_3d_bbox_transform_inv(rois[f, :, 1:5], a3d[f].clone(), mean_pred,
                                                                            scales=float(info[f, 6])))
                res["e_var"].append(R.bt.lidar_3d_uncertainty_transform_inv(rois[f, :, 1:5], a3d[f].clone(), mean_pred, var))
                res["a_var"].append(R.bt.lidar_3d_uncertainty_transform_inv(rois[f, :, 1:5], a3d[f].clone(), mean_pred,
                                                                            a_var[f]))
            else:
                res["boxes"].append(R.bt.clip_boxes(R.bt.bbox_transform_inv(rois[f, :, 1:5], mean_pred), info[f].numpy()))
                res["boxes_scaled"].append(R.bt.clip_boxes(
                    R.bt.bbox_transform_inv(rois[f, :, 1:5], mean_pred, scales=float(info[f, 6])), info[f].numpy()))
                res["e_var"].append(var)
                res["a_var"].append(a_var[f])
            probs = torch.mean(torch.softmax(cls[:, f], dim=2), dim=0)
            res["probs"].append(probs)
            res["ent"].append(R.lu.categorical_entropy(probs))
            res["mi"].append(R.lu.categorical_mutual_information(cls[:, f]))
        out.update({f"{tag}_bbox": _np(bbox), f"{tag}_cls": _np(cls), f"{tag}_rois": _np(rois), f"{tag}_a3d": _np(a3d),
                    f"{tag}_a_var_in": _np(a_var), f"{tag}_info": _np(info), f"{tag}_stds": _np(stds[:E]),
                    f"{tag}_means": _np(means[:E])})
        out.update({f"{tag}_{k}": _np(torch.stack(v)) for k, v in res.items()})
    np.savez_compressed(os.path.join(OUT, "head_tail.npz"), **out)
    print("head_tail.npz", os.path.getsize(os.path.join(OUT, "head_tail.npz")))


def gen_eval():
    """tests/golden/eval.npz: result files + detection matching from the reference's own
    writers and its own waymo_eval loop, run with the missing utils/eval_utils.py bound to oracle/eval_oracle.py."""
    import json
    import tempfile
    import types
    from . import eval_oracle as E
    ns = ref_import.load_eval()
    cfg = ns.cfg
    rng = np.random.RandomState(SEED)
    out = {}
    tmp = tempfile.mkdtemp(prefix="b2d_eval_")
    cfg.ROOT_DIR = tmp                                   # get_output_dir() builds its path under cfg.ROOT_DIR
    classes = ('__background__', 'vehicle', 'pedestrian')
    tokens = ['%07d.png' % (1000 * s + i) for s, i in ((0, 3), (0, 4), (1, 0), (1, 7), (2, 5), (2, 6))]
    F, K, U = len(tokens), len(classes), 3

    for nt, E_, eval_type in (("image", 4, "2d"), ("lidar", 7, "bev_aa")):
        cfg.NET_TYPE = nt
        # ---- ground truth: frame 3 is absent from the label file, frame 4 has no boxes at all
        labels, recs_in = [], {}
        for f, tok in enumerate(tokens):
            if f == 3:
                continue
            G = 0 if f == 4 else 5 + f
            if nt == "image":
                xy = rng.rand(G, 2) * np.array([1500.0, 900.0])
                wh = rng.rand(G, 2) * 200 + 30
                boxes = np.concatenate((xy, xy + wh), 1)
                dcxy = rng.rand(2, 2) * np.array([1500.0, 900.0])
                boxes_dc = np.concatenate((dcxy, dcxy + 150), 1)
            else:
                boxes = np.concatenate((rng.rand(G, 2) * 60 - 30, rng.rand(G, 1), rng.rand(G, 3) * np.array([4.0, 2.0, 1.0]) + 1.5,
                                        rng.rand(G, 1) * 3 - 1.5), 1)
                boxes_dc = np.concatenate((rng.rand(2, 2) * 60 - 30, np.zeros((2, 1)), np.full((2, 3), 6.0), np.zeros((2, 1))), 1)
            rec = dict(boxes=boxes.astype(np.float32), boxes_dc=boxes_dc.astype(np.float32),
                       gt_classes=rng.randint(1, K, G), difficulty=rng.randint(0, 4, G), ignore=rng.rand(G) < 0.2,
                       pts=rng.randint(5, 500, G), ids=['trk%d_%d' % (f, j) for j in range(G)], scene_idx=f // 2,
                       scene_desc='scene%d' % (f // 2))
            recs_in[tok] = rec
            labels.append({'assoc_frame': str(int(''.join(c for c in tok if c.isdigit()))).zfill(7), 'token': tok})

        class FakeDb(object):
            name = 'waymo_fake'
            _devkit_path = tmp
            _class_to_ind = dict(zip(classes, range(K)))

            def __init__(self):
                self.classes = classes

            def _get_index_for_mode(self, mode):
                return tokens
            _get_results_file_template = ns.db.db._get_results_file_template

            def _load_waymo_annotation(self, frame, label, remove_without_gt=False, tod_filter_list=None, en_aux_features=False):
                r = recs_in[label['token']]
                G = len(r['gt_classes'])
                return dict(boxes=r['boxes'].copy(), boxes_dc=r['boxes_dc'].copy(), gt_classes=r['gt_classes'].copy(),
                            gt_overlaps=np.zeros((G, K)), det=np.zeros(G, dtype=bool), ignore=r['ignore'].copy(),
                            hit=np.zeros(G, dtype=bool), ids=list(r['ids']), pts=r['pts'].copy(),
                            difficulty=r['difficulty'].copy(), scene_idx=r['scene_idx'], scene_desc=r['scene_desc'])
        db = FakeDb()
        os.makedirs(os.path.join(tmp, 'test', 'labels'), exist_ok=True)
        with open(os.path.join(tmp, 'test', 'labels', E.get_labels_filename(db, eval_type)), 'w') as fh:
            json.dump(labels, fh)
        # ---- detections: jittered GT (some twice), clutter, detections in frames that are not evaluated
        all_boxes = [[np.empty(0) for _ in range(F)] for _ in range(K)]
        for f, tok in enumerate(tokens):
            for c in range(1, K):
                rows = []
                r = recs_in.get(tok)
                if r is not None:
                    for j in np.where(r['gt_classes'] == c)[0]:
                        for rep in range(1 + (j % 3 == 0)):
                            b = r['boxes'][j].astype(np.float64) + rng.randn(E_) * (4.0 if nt == "image" else 0.15) * (1 + rep)
                            rows.append(np.concatenate((b, [rng.rand() * 0.9 + 0.1], rng.rand(U))))
                    for j in range(2):                                   # on top of the don't-care boxes
                        b = r['boxes_dc'][j].astype(np.float64) + rng.randn(E_) * 0.5
                        rows.append(np.concatenate((b, [rng.rand()], rng.rand(U))))
                for _ in range(3):                                       # clutter
                    if nt == "image":
                        xy = rng.rand(2) * np.array([1500.0, 900.0])
                        b = np.concatenate((xy, xy + rng.rand(2) * 100 + 20))
                    else:
                        b = np.concatenate((rng.rand(2) * 60 - 30, [0.5], rng.rand(3) * 3 + 1, [0.0]))
                    rows.append(np.concatenate((b, [rng.rand() * 0.5], rng.rand(U))))
                if f == 5 and c == 2:
                    rows = []                                            # a class without detections in a frame
                if rows:
                    all_boxes[c][f] = np.stack(rows).astype(np.float32)
        # confidences unique at the three decimals the result file keeps: the reference orders equal confidences
        # with numpy's unstable argsort (waymo_eval.py:129), which no fixture should depend on
        for c in range(1, K):
            total = sum(a.shape[0] for a in all_boxes[c] if a.size)
            uniq = (rng.permutation(998)[:total] + 1) / 1000.0
            p0 = 0
            for f in range(F):
                if all_boxes[c][f].size:
                    m = all_boxes[c][f].shape[0]
                    all_boxes[c][f][:, E_] = uniq[p0:p0 + m]
                    p0 += m
        out_dir = os.path.join(tmp, nt)
        os.makedirs(out_dir, exist_ok=True)
        writer = ns.db.db._write_image_results_file if nt == "image" else ns.db.db._write_lidar_results_file
        writer(db, all_boxes, out_dir, 'test')
        for c in range(1, K):
            path = db._get_results_file_template('test', classes[c], out_dir)
            with open(path) as fh:
                out[f"{nt}_lines_{c}"] = np.array(fh.readlines())
        for c in range(1, K):
            for f in range(F):
                out[f"{nt}_dets_{c}_{f}"] = all_boxes[c][f]
        for tok, r in recs_in.items():
            f = tokens.index(tok)
            for k in ('boxes', 'boxes_dc', 'gt_classes', 'difficulty', 'ignore', 'pts'):
                out[f"{nt}_gt_{f}_{k}"] = np.asarray(r[k])
        # ---- the reference's evaluation loop
        detpath = os.path.join(out_dir, 'results', 'det_test_{:s}.txt')
        for ign_dc in (False, True):
            cfg.TEST.IGNORE_DC = ign_dc
            for c in range(1, K):
                E.SAVED.clear()
                mrec, mprec, mp = ns.waymo_eval.waymo_eval(detpath, db, tokens, classes[c], None, 'test', ovthresh=0.5,
                                                           eval_type=eval_type, d_levels=2)
                tag = f"{nt}_eval_{c}_{int(ign_dc)}"
                out[tag + "_map"] = np.asarray(mp, dtype=np.float64)
                out[tag + "_mrec"] = np.asarray(mrec, dtype=np.float64)
                out[tag + "_mprec"] = np.asarray(mprec, dtype=np.float64)
                out[tag + "_results"] = np.array(E.SAVED['{}_detection_results.txt'.format(classes[c])])
        cfg.TEST.IGNORE_DC = False
    out["tokens"] = np.array(tokens)
    out["classes"] = np.array(classes)
    # ---- stack_uncertainties (model/test.py:260-270) and bbox_voxel_grid_to_pc (utils/bbox.py:140-162)
    cls_boxes = (rng.rand(9, 8) * 50).astype(np.float32)
    ucs = {'a_bbox_var': rng.rand(9, 7).astype(np.float32), 'e_bbox_var': rng.rand(9, 7).astype(np.float32),
           'a_entropy': rng.rand(9, 1).astype(np.float32), 'e_mutual_info': rng.rand(9, 1).astype(np.float32)}
    out["stack_boxes"], out["stack_hstack"] = cls_boxes, ns.test.stack_uncertainties(cls_boxes, ucs, 16)
    for k, v in ucs.items():
        out["stack_uc_" + k] = v
    extents = [0.0, -40.0, 0.0, 70.0, 40.0, 3.0]
    info = np.array([0, 700, 0, 800, 0, 12, 1.0], dtype=np.float32)
    info2 = np.array([0, 1400, 0, 1600, 0, 24, 2.0], dtype=np.float32)
    vg = (rng.rand(11, 8) * 300).astype(np.float32)
    out["vg_in"], out["vg_extents"], out["vg_info"], out["vg_info2"] = vg, np.array(extents), info, info2
    out["vg_out"] = ns.ub.bbox_voxel_grid_to_pc(vg.copy(), extents, info)
    out["vg_out2"] = ns.ub.bbox_voxel_grid_to_pc(vg.copy(), extents, info2)
    out["vg_out_aabb"] = ns.ub.bbox_voxel_grid_to_pc(vg[:, :4].copy(), extents, info, aabb=True)
    np.savez_compressed(os.path.join(OUT, "eval.npz"), **out)
    print("eval.npz", os.path.getsize(os.path.join(OUT, "eval.npz")))


if __name__ == "__main__":
    if "--only-eval" in sys.argv:
        sys.exit(gen_eval())
    if "--only-bev" in sys.argv:
        sys.exit(gen_bev())
    if "--only-head-tail" in sys.argv:
        sys.exit(gen_head_tail())
    rc = main()
    gen_bev()
    gen_head_tail()
    gen_eval()
    sys.exit(rc)
