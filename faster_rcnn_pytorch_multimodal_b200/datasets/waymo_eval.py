"""Detection evaluation: ``datasets/waymo_eval.py:44-250`` (the same loop serves
``kitti_eval.py`` and ``cadc_eval.py``).

The reference reads the per-class result file, orders the detections by confidence and walks them one by one against
the ground truth of their frame in Python.  Here the result file is parsed the same way, the ordered detections and
the per-frame ground truth go to the device once, ``b2d_eval_match`` does the matching loop (one warp per frame) and
the cumulative precision / recall / AP assembly follows the reference line by line on the few kilobytes that come
back.

Differences a maintainer should know:
  * ``class_recs`` (what the reference's ``load_recs`` builds from the dataset's label files, waymo_eval.py:266-310)
    is an argument: dataset loading is outside this repository's path.  Per frame a dict with ``filename``,
    ``ignore_frame`` and, unless ignored, ``boxes`` [G,E], ``boxes_dc`` [D,E], ``ignore`` [G], ``difficulty`` [G].
  * ``iou`` and ``ap`` lived in the reference's ``utils/eval_utils.py``, which is missing from the snapshot
    (SURVEY.md F2): they follow oracle/eval_oracle.py (PASCAL-VOC overlap with the fork's +1 convention for
    '2d', axis-aligned footprint for 'bev_aa'; all-point interpolated AP) and are parity-unpinned.  The rotated
    'bev' / '3d' overlaps are not provided.
  * Equal confidences are ordered by position in the file (stable); the reference's ``np.argsort`` leaves them open.
"""
from typing import Sequence

import numpy as np
import torch

from .._lib import check, lib, ptr, stream_ptr
from ..model.config import cfg

_MODES = {'2d': 0, 'bev_aa': 1}


def voc_ap(rec, prec) -> float:
    """All-point interpolated average precision (eval_utils.ap, waymo_eval.py:247)."""
    mrec = np.concatenate(([0.0], np.asarray(rec, dtype=np.float64), [1.0]))
    mpre = np.concatenate(([0.0], np.asarray(prec, dtype=np.float64), [0.0]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1]))


def parse_result_file(lines: Sequence[str], bbox_elem: int):
    """waymo_eval.py:97-108."""
    split = [x.strip().split(' ') for x in lines]
    tokens = [x[1] for x in split]
    conf = np.array([float(x[2]) for x in split])
    bb = np.array([[float(z) for z in x[3:3 + bbox_elem]] for x in split]).reshape(len(split), bbox_elem)
    return tokens, conf, bb


def match_detections(tokens, confidence, boxes, class_recs, ovthresh=0.5, eval_type='2d', ignore_dc=None,
                     ovthresh_dc=0.5, device=None):
    """The matching loop (waymo_eval.py:120-215) on the device.
    -> order (detection indices by descending confidence), code / ovmax / jmax / difficulty per position of that order
    (code -1: the detection's frame is not evaluated)."""
    if eval_type not in _MODES:
        raise NotImplementedError(f"eval_type {eval_type!r}: only '2d' and 'bev_aa' overlaps are provided")
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    ignore_dc = bool(cfg.TEST.IGNORE_DC) if ignore_dc is None else bool(ignore_dc)
    n = len(tokens)
    E = boxes.shape[1] if n else (4 if eval_type == '2d' else 7)
    order = np.argsort(-np.asarray(confidence, dtype=np.float64), kind='stable') if n else np.zeros(0, dtype=np.int64)
    rec_of = {r['filename']: i for i, r in enumerate(class_recs) if not r.get('ignore_frame', False)}
    det_rec = np.array([rec_of.get(tokens[d], -1) for d in order], dtype=np.int64)
    # detections grouped by frame, ascending position within a frame
    listed = np.nonzero(det_rec >= 0)[0]
    grouped = listed[np.argsort(det_rec[listed], kind='stable')]
    n_rec = len(class_recs)
    det_off = np.zeros(n_rec + 1, dtype=np.int32)
    np.add.at(det_off, det_rec[listed] + 1, 1)
    det_off = np.cumsum(det_off).astype(np.int32)
    gt_off, dc_off = np.zeros(n_rec + 1, dtype=np.int32), np.zeros(n_rec + 1, dtype=np.int32)
    gts, dcs, flags = [], [], []
    for i, r in enumerate(class_recs):
        if r.get('ignore_frame', False):
            gt_off[i + 1], dc_off[i + 1] = gt_off[i], dc_off[i]
            continue
        b = np.asarray(r['boxes'], dtype=np.float64).reshape(-1, E)
        d = np.asarray(r['boxes_dc'], dtype=np.float64).reshape(-1, E)
        gts.append(b)
        dcs.append(d)
        flags.append(np.asarray(r['ignore']).astype(np.int32) | (np.asarray(r['difficulty']).astype(np.int32) << 8))
        gt_off[i + 1], dc_off[i + 1] = gt_off[i] + b.shape[0], dc_off[i] + d.shape[0]
    cat = lambda xs, w, dt: (np.concatenate(xs) if xs else np.zeros((0, w) if w else 0)).astype(dt)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    d_boxes = t(np.asarray(boxes, dtype=np.float64)[order].reshape(n, E))
    d_gt, d_dc, d_fl = t(cat(gts, E, np.float64)), t(cat(dcs, E, np.float64)), t(cat(flags, 0, np.int32))
    code = torch.full((max(n, 1),), -1, dtype=torch.int32, device=device)
    ovmax = torch.full((max(n, 1),), float('-inf'), dtype=torch.float64, device=device)
    jmax = torch.zeros(max(n, 1), dtype=torch.int32, device=device)
    diff = torch.full((max(n, 1),), -1, dtype=torch.int32, device=device)
    hit = torch.zeros(max(int(gt_off[-1]), 1), dtype=torch.uint8, device=device)
    # (every device buffer is held in a local until the results are read back: a temporary would return to the
    # caching allocator before the kernel runs)
    d_det_off, d_grouped, d_gt_off, d_dc_off = t(det_off), t(grouped.astype(np.int32)), t(gt_off), t(dc_off)
    with torch.cuda.device(device):
        check(lib(device).b2d_eval_match(n, n_rec, E, _MODES[eval_type], ptr(d_boxes), ptr(d_det_off), ptr(d_grouped),
                                         ptr(d_gt_off), ptr(d_gt), ptr(d_fl), ptr(d_dc_off), ptr(d_dc), float(ovthresh),
                                         float(ovthresh_dc), int(ignore_dc), ptr(code), ptr(ovmax), ptr(jmax), ptr(diff),
                                         ptr(hit), stream_ptr(device)), "b2d_eval_match")
    return order, code[:n].cpu().numpy(), ovmax[:n].cpu().numpy(), jmax[:n].cpu().numpy(), diff[:n].cpu().numpy()


def count_npos(class_recs, d_levels):
    """waymo_eval.py:252-262."""
    npos = np.zeros((len(class_recs), d_levels))
    for i, rec in enumerate(class_recs):
        if rec.get('ignore_frame', False) is False:
            ok = ~np.asarray(rec['ignore']).astype(bool)
            diff = np.asarray(rec['difficulty'])
            npos[i, 1] += int(np.sum(ok & (diff <= 2)))
            npos[i, 0] += int(np.sum(ok & (diff <= 1)))
    return npos


def waymo_eval(detpath, class_recs, classname, ovthresh=0.5, eval_type='2d', d_levels=2, ignore_dc=None, device=None):
    """``mrec, mprec, map = waymo_eval(...)`` (waymo_eval.py:44-250).  ``detpath.format(classname)`` is the result file
    written by ``datasets.results.write_*_results_file``.  As in the reference the three returned names are ONE
    array (`map = mrec = mprec = np.zeros(...)`, :224): each holds the AP per difficulty level."""
    with open(detpath.format(classname), 'r') as f:
        lines = f.readlines()
    bbox_elem = int(cfg[cfg.NET_TYPE.upper()].NUM_BBOX_ELEM)
    tokens, conf, bb = parse_result_file(lines, bbox_elem)
    return evaluate(tokens, conf, bb, class_recs, ovthresh, eval_type, d_levels, ignore_dc, device)["map_triplet"]


def evaluate(tokens, confidence, boxes, class_recs, ovthresh=0.5, eval_type='2d', d_levels=2, ignore_dc=None, device=None):
    """Matching + the precision / recall / AP assembly (waymo_eval.py:224-250) -> dict with tp, fp [n_evaluated,
    d_levels], code / ovmax / jmax per ordered detection, and `map_triplet` = what the reference returns."""
    order, code, ovmax, jmax, diff = match_detections(tokens, confidence, boxes, class_recs, ovthresh, eval_type, ignore_dc,
                                                      device=device)
    ev = code >= 0                                     # detections whose frame is evaluated advance `idx` (:216)
    c, df = code[ev], diff[ev]
    tp, fp = np.zeros((c.size, d_levels)), np.zeros((c.size, d_levels))
    for lvl, lim in ((1, 2), (0, 1)):
        tp[:, lvl] = (c == 1) & (df <= lim)
        fp[:, lvl] = ((c == 2) & (df <= lim)) | (c == 3)
    npos = count_npos(class_recs, d_levels)
    shared = np.zeros((d_levels,))
    fp_sum, tp_sum = np.cumsum(fp, axis=0), np.cumsum(tp, axis=0)
    npos_sum = np.sum(npos, axis=0)
    aps = np.zeros(d_levels)
    for i in range(d_levels):
        npos_d = npos_sum[i]
        if npos_d == 0:
            npos_d = np.sum([1])
        rec = tp_sum[:, i] / npos_d.astype(float)
        prec = tp_sum[:, i] / np.maximum(tp_sum[:, i] + fp_sum[:, i], np.finfo(np.float64).eps)
        if rec.size:
            rec, prec = zip(*sorted(zip(rec, prec)))
        aps[i] = voc_ap(rec, prec)
        shared[i] = aps[i]
    return {"map_triplet": (shared, shared, shared), "ap": aps, "tp": tp, "fp": fp, "order": order, "code": code,
            "ovmax": ovmax, "jmax": jmax, "difficulty": diff, "npos": npos}
