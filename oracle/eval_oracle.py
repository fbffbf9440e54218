"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's result files and
detection-vs-ground-truth matching.

What is pinned and what is not
  * PINNED by the reference's own code run unmodified (oracle/gen_golden.py `gen_eval`, fixtures in
    tests/golden/eval.npz): the result-file writers `db._write_image_results_file` / `_write_lidar_results_file`
    (datasets/db.py:305-367), `stack_uncertainties` (model/test.py:260-270), `bbox_voxel_grid_to_pc`
    (utils/bbox.py:140-162) and the whole body of `waymo_eval` (datasets/waymo_eval.py:44-250: parsing the result
    file, the confidence-ordered greedy matching with don't-care and difficulty handling, the cumulative
    precision / recall assembly).
  * PARITY UNPINNED: the reference imports `utils.eval_utils`, a module that is MISSING from the snapshot
    (SURVEY.md F2).  `waymo_eval` calls eight of its functions; they are restated below from their call sites and
    from the PASCAL-VOC code this evaluation descends from (`iou` = VOC overlap with the +1 pixel convention the
    fork's own `bbox_overlaps` uses, utils/bbox.py:5-33; `ap` = VOC all-point interpolated average precision), and
    the golden fixture is produced by the reference's loop running ON these restatements.  A different `iou` / `ap`
    in the lost file would change matches and AP values but not the loop.
"""
import os
import re
from typing import Dict, List

import numpy as np


# ------------------------------------------------------------------------------------------
# restatement of the missing utils/eval_utils.py (call sites: datasets/waymo_eval.py:90-247)
# ------------------------------------------------------------------------------------------
def get_labels_filename(db, eval_type):
    """waymo_eval.py:90.  The fork's label files are `labels/<name>.json` under the devkit path."""
    return "image_labels.json" if eval_type == "2d" else "lidar_labels.json"


def extract_uncertainties(bbox_elem, splitlines):
    """waymo_eval.py:109.  Columns after `idx token score box[bbox_elem]` of a result line are the stacked
    uncertainties in the order stack_uncertainties wrote them (model/test.py:260-270): returns (per-scene
    accumulators, {name: [n_det, width]}).  Without the lost file the column names are unknown: one block `uc`."""
    n_extra = (len(splitlines[0]) - 3 - bbox_elem) if splitlines else 0
    if n_extra <= 0:
        return {}, {}
    vals = np.array([[float(z) for z in x[3 + bbox_elem:]] for x in splitlines])
    from_cfg_scenes = 1000
    return {"uc": np.zeros((from_cfg_scenes, n_extra))}, {"uc": vals}


def find_rec(class_recs, token):
    """waymo_eval.py:142: the record of the frame a detection belongs to (None: frame not evaluated)."""
    for rec in class_recs:
        if rec["filename"] == token:
            return None if rec.get("ignore_frame", False) else rec
    return None


def iou(bbgt, bb, eval_type):
    """waymo_eval.py:164,168: overlaps of ONE detection with every GT box of its frame, float64.
    '2d': [x1,y1,x2,y2] with the +1 pixel convention; 'bev_aa': axis-aligned footprint of [xc,yc,zc,l,w,h,ry]."""
    bbgt = np.asarray(bbgt, dtype=np.float64)
    bb = np.asarray(bb, dtype=np.float64)
    if eval_type == "2d":
        g, d, one = bbgt[:, :4], bb[:4], 1.0
    elif eval_type == "bev_aa":
        g = np.stack((bbgt[:, 0] - bbgt[:, 3] / 2, bbgt[:, 1] - bbgt[:, 4] / 2, bbgt[:, 0] + bbgt[:, 3] / 2,
                      bbgt[:, 1] + bbgt[:, 4] / 2), axis=1)
        d = np.array([bb[0] - bb[3] / 2, bb[1] - bb[4] / 2, bb[0] + bb[3] / 2, bb[1] + bb[4] / 2])
        one = 0.0
    else:
        raise NotImplementedError("eval_type %r: the rotated overlaps lived in the missing eval_utils.py" % eval_type)
    ixmin = np.maximum(g[:, 0], d[0])
    iymin = np.maximum(g[:, 1], d[1])
    ixmax = np.minimum(g[:, 2], d[2])
    iymax = np.minimum(g[:, 3], d[3])
    iw = np.maximum(ixmax - ixmin + one, 0.0)
    ih = np.maximum(iymax - iymin + one, 0.0)
    inters = iw * ih
    uni = (d[2] - d[0] + one) * (d[3] - d[1] + one) + (g[:, 2] - g[:, 0] + one) * (g[:, 3] - g[:, 1] + one) - inters
    return inters / uni


def ap(rec, prec):
    """waymo_eval.py:247: VOC all-point average precision (area under the monotone precision envelope)."""
    mrec = np.concatenate(([0.0], np.asarray(rec, dtype=np.float64), [1.0]))
    mpre = np.concatenate(([0.0], np.asarray(prec, dtype=np.float64), [0.0]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1]))


def write_scene_uncertainty(uc_avg, scene_dets, scene_idx):
    """waymo_eval.py:221: per-scene text summary; empty when the scene has no detections."""
    if scene_dets == 0 or not uc_avg:
        return ""
    return "scene: {} num_dets: {} ".format(scene_idx, int(scene_dets)) + " ".join(
        "{}: {}".format(k, " ".join("{:.10f}".format(x) for x in v[scene_idx] / scene_dets)) for k, v in uc_avg.items())


def display_frame_counts(tp_frame, fp_frame, npos_frame):
    """waymo_eval.py:228 (debug print only)."""


SAVED: Dict[str, List[str]] = {}


def save_detection_results(results, out_dir, out_file):
    """waymo_eval.py:231,234: one line per evaluated detection."""
    SAVED[out_file] = list(results)
    if out_dir and os.path.isdir(out_dir):
        with open(os.path.join(out_dir, out_file), "w") as f:
            for line in results:
                f.write(line + "\n")


# ------------------------------------------------------------------------------------------
# result files (datasets/db.py:305-367, model/test.py:252-270, utils/bbox.py:140-162)
# ------------------------------------------------------------------------------------------
def bbox_voxel_grid_to_pc(bboxes, bev_extants, info, aabb=False):
    """utils/bbox.py:140-162 (in place, as the reference)."""
    scale = info[6]
    s_info = np.asarray(info[0:6]) * 1 / scale
    fx = (bev_extants[3] - bev_extants[0]) / (s_info[1] - s_info[0])
    fy = (bev_extants[4] - bev_extants[1]) / (s_info[3] - s_info[2])
    if aabb:
        bboxes[:, 0] = bboxes[:, 0] * fx + bev_extants[0]
        bboxes[:, 1] = bboxes[:, 1] * fy + bev_extants[1]
        bboxes[:, 2] = bboxes[:, 2] * fx + bev_extants[0]
        bboxes[:, 3] = bboxes[:, 3] * fy + bev_extants[1]
    else:
        bboxes[:, 0] = bboxes[:, 0] * fx + bev_extants[0]
        bboxes[:, 1] = bboxes[:, 1] * fy + bev_extants[1]
        bboxes[:, 3] = bboxes[:, 3] * fx
        bboxes[:, 4] = bboxes[:, 4] * fy
    return bboxes


def result_lines(all_boxes, cls_ind, frame_tokens, lidar, num_bbox_elem=7):
    """The lines db._write_image_results_file (:305-332) / _write_lidar_results_file (:334-367) write for a class."""
    lines = []
    for ind, token in enumerate(frame_tokens):
        dets = all_boxes[cls_ind][ind]
        if dets.size == 0:
            continue
        for k in range(dets.shape[0]):
            if lidar:
                s = '{:d} {:s} {:.3f} {:.3f} {:.3f} {:.3f} {:.3f} {:.3f} {:.3f} {:.5f}'.format(
                    ind, token, dets[k, 7], dets[k, 0], dets[k, 1], dets[k, 2], dets[k, 3], dets[k, 4], dets[k, 5], dets[k, 6])
                first_uc = 8 if dets.shape[1] > num_bbox_elem + 1 else dets.shape[1]
            else:
                s = '{:d} {:s} {:.3f} {:.1f} {:.1f} {:.1f} {:.1f}'.format(ind, token, dets[k, 4], dets[k, 0], dets[k, 1],
                                                                          dets[k, 2], dets[k, 3])
                first_uc = 5
            for l in range(first_uc, dets.shape[1]):
                s += ' {:.10f}'.format(dets[k, l])
            lines.append(s + '\n')
    return lines


# ------------------------------------------------------------------------------------------
# the matching loop and the AP assembly (datasets/waymo_eval.py:96-250), restated on in-memory inputs
# ------------------------------------------------------------------------------------------
def parse_result_lines(lines, bbox_elem):
    """waymo_eval.py:97-108."""
    split = [x.strip().split(' ') for x in lines]
    tokens = [x[1] for x in split]
    conf = np.array([float(x[2]) for x in split])
    bb = np.array([[float(z) for z in x[3:3 + bbox_elem]] for x in split])
    return tokens, conf, bb, split


def count_npos(class_recs, d_levels):
    """waymo_eval.py:252-262."""
    npos = np.zeros((len(class_recs), d_levels))
    for i, rec in enumerate(class_recs):
        if rec['ignore_frame'] is False:
            for j, ign in enumerate(rec['ignore']):
                if not ign:
                    if rec['difficulty'][j] <= 2:
                        npos[i, 1] += 1
                    if rec['difficulty'][j] <= 1:
                        npos[i, 0] += 1
    return npos


def match_and_score(tokens, conf, bb_all, class_recs, ovthresh=0.5, eval_type='2d', d_levels=2, ignore_dc=True,
                    ovthresh_dc=0.5):
    """-> dict(mrec, mprec, map, tp [n,d_levels], fp [n,d_levels], code [n_det] in confidence order
    (-1 frame not evaluated, 0 nothing recorded, 1 tp, 2 duplicate fp, 3 low-overlap fp), ovmax, jmax)."""
    recs = [dict(r) for r in class_recs]
    for r in recs:
        if not r.get('ignore_frame', False):
            r['hit'] = np.zeros(len(r['ignore']), dtype=bool)
    n = len(tokens)
    tp, fp = np.zeros((n, d_levels)), np.zeros((n, d_levels))
    npos = count_npos(recs, d_levels)
    order = np.argsort(-conf, kind='stable') if n else np.zeros(0, dtype=np.int64)
    code = np.full(n, -1, dtype=np.int64)
    ovm = np.full(n, -np.inf)
    jm = np.zeros(n, dtype=np.int64)
    idx = 0
    for pos, d in enumerate(order):
        R = find_rec(recs, tokens[d])
        if R is None:
            continue
        bb = bb_all[d].astype(float)
        ovmax, jmax, ovmax_dc = -np.inf, 0, 0
        BBGT, BBGT_dc = R['boxes'].astype(float), R['boxes_dc'].astype(float)
        if BBGT_dc.size > 0 and ignore_dc:
            ovmax_dc = np.max(iou(BBGT_dc, bb, eval_type))
        if BBGT.size > 0:
            ov = iou(BBGT, bb, eval_type)
            ovmax, jmax = np.max(ov), int(np.argmax(ov))
        code[pos] = 0
        if ovmax > ovthresh and ovmax_dc < ovthresh_dc:
            if not R['ignore'][jmax]:
                if not R['hit'][jmax]:
                    if R['difficulty'][jmax] <= 2:
                        tp[idx, 1] += 1
                    if R['difficulty'][jmax] <= 1:
                        tp[idx, 0] += 1
                    R['hit'][jmax] = True
                    code[pos] = 1
                else:
                    if R['difficulty'][jmax] <= 2:
                        fp[idx, 1] += 1
                    if R['difficulty'][jmax] <= 1:
                        fp[idx, 0] += 1
                    code[pos] = 2
        elif BBGT.size > 0 and ovmax_dc < ovthresh_dc:
            fp[idx, 0] += 1
            fp[idx, 1] += 1
            code[pos] = 3
        ovm[pos], jm[pos] = ovmax, jmax
        idx += 1
    out = assemble_ap(tp, fp, npos, d_levels)
    out.update(tp=tp, fp=fp, code=code, ovmax=ovm, jmax=jm, order=order)
    return out


def assemble_ap(tp, fp, npos, d_levels):
    """waymo_eval.py:224-250.  NOTE `map = mrec = mprec = np.zeros(...)` makes the three names ONE array in the
    reference: each level's slot is written three times and keeps the last write (the AP)."""
    shared = np.zeros((d_levels,))
    fp_sum, tp_sum = np.cumsum(fp, axis=0), np.cumsum(tp, axis=0)
    npos_sum = np.sum(npos, axis=0)
    for i in range(d_levels):
        npos_d = npos_sum[i]
        if npos_d == 0:
            npos_d = np.sum([1])
        rec = tp_sum[:, i] / npos_d.astype(float)
        prec = tp_sum[:, i] / np.maximum(tp_sum[:, i] + fp_sum[:, i], np.finfo(np.float64).eps)
        rec, prec = zip(*sorted(zip(rec, prec)))
        shared[i] = np.average(prec)
        shared[i] = np.average(rec)
        shared[i] = ap(rec, prec)
    return {"mrec": shared, "mprec": shared, "map": shared}
