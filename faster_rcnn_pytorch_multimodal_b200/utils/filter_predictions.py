"""Final per-class detection filter with the reference's signatures (lib/utils/filter_predictions.py).

The reference loops over classes and, per class, does a boolean-mask compaction, a torchvision NMS and
several device->host copies (:45-72, :101-125).  Here ONE kernel launch (`ops.final_detections`) does all
classes on the device and the results come back in a single padded record, which is also what the
multi-GPU end-of-stream gather ships (`stream.gather_detections`).
"""
import numpy as np
import torch

from .. import ops
from ..model.config import cfg

_ROW_KEYS = (("EN_CLS_ALEATORIC", ("a_entropy", "a_mutual_info", "a_cls_var")),
             ("EN_CLS_EPISTEMIC", ("e_entropy", "e_mutual_info", "e_cls_var")))
_BOX_KEYS = (("EN_BBOX_ALEATORIC", "a_bbox_var"), ("EN_BBOX_EPISTEMIC", "e_bbox_var"))


def _elem(db_type):
    if db_type == "image":
        return cfg.IMAGE.NUM_BBOX_ELEM
    if db_type == "lidar":
        return cfg.LIDAR.NUM_BBOX_ELEM
    return None


def nms_hstack_torch(scores, mean_boxes, thresh, c, bbox_elem, db_type):
    """filter_predictions.py:45-72 -> (cls_dets [m, E+1] float32 ndarray, inds tensor, keep ndarray).

    `keep` indexes `inds` in descending score order, as torchvision's nms returns it."""
    inds = torch.where(scores[:, c] > thresh)[0]
    if inds.shape[0] == 0:
        return np.empty(0), [], []
    # the reference clamps image boxes in filter_and_draw_prep BEFORE this call (:82-91), not here
    info = torch.zeros(1, 7, device=scores.device)
    dets, det_roi, counts, _, _ = ops.final_detections(scores.unsqueeze(0), mean_boxes.unsqueeze(0), info, bbox_elem,
                                                       "lidar" if db_type == "lidar" else "image_noclamp", thresh,
                                                       cfg.TEST.NMS_THRESH)
    m = int(counts[0, c].item())
    keep = torch.searchsorted(inds, det_roi[0, c, :m].long())
    return dets[0, c, :m].cpu().numpy(), inds, keep.cpu().numpy()


def filter_and_draw_prep(rois, cls_score, pred_boxes, uncertainties, info, num_classes, thresh=0.1, db_type='none'):
    """filter_predictions.py:75-130 -> (rois [R,4] ndarray, all_boxes, all_uncertainty).

    all_boxes[j] is the [m, E+1] float32 array of class j (rows by descending score); all_uncertainty[j]
    holds the gathered uncertainty arrays of the kept detections for every enabled cfg.UC flag (gathered
    from the original tensors for every class; the reference reuses the dict it has just overwritten,
    which is only well defined for one foreground class).  Unlike the reference, `pred_boxes` is not
    clamped in place."""
    bbox_elem = _elem(db_type)
    if bbox_elem is None:
        return None
    dev = cls_score.device
    info_t = torch.as_tensor(np.asarray(info, dtype=np.float32)).view(1, 7).to(dev)
    R = cls_score.shape[0]
    row_cols, row_layout = [], []
    for flag, keys in _ROW_KEYS:
        if cfg.UC[flag]:
            for k in keys:
                v = uncertainties[k].reshape(R, -1).float()
                row_layout.append((k, v.shape[1], k.endswith("_cls_var")))
                row_cols.append(v)
    box_keys = [k for flag, k in _BOX_KEYS if cfg.UC[flag]]
    uc_row = torch.cat(row_cols, dim=1).unsqueeze(0) if row_cols else None
    uc_cls = torch.stack([uncertainties[k].float() for k in box_keys], dim=1).unsqueeze(0) if box_keys else None
    dets, det_roi, counts, o_row, o_cls = ops.final_detections(cls_score.unsqueeze(0), pred_boxes.unsqueeze(0), info_t,
                                                               bbox_elem, db_type, thresh, cfg.TEST.NMS_THRESH,
                                                               uc_row=uc_row, uc_cls=uc_cls)
    # one device->host transfer per output
    counts_h = counts[0].cpu().numpy()
    dets_h = dets[0].cpu().numpy()
    o_row_h = o_row[0].cpu().numpy() if o_row is not None else None
    o_cls_h = o_cls[0].cpu().numpy() if o_cls is not None else None
    all_boxes = [[] for _ in range(num_classes)]
    all_uncertainty = [{} for _ in range(num_classes)]
    for j in range(1, num_classes):
        m = int(counts_h[j])
        if m == 0:
            all_boxes[j] = np.empty(0)
            all_uncertainty[j] = uncertainties
            continue
        all_boxes[j] = dets_h[j, :m].astype(np.float32, copy=False)
        uc = dict(uncertainties)
        col = 0
        for k, width, is_var in row_layout:
            uc[k] = o_row_h[j, :m, col:col + width]           # 'cls' -> [m,1], 'cls_var' -> [m,K]   (:24-36)
            col += width
        for u, k in enumerate(box_keys):
            uc[k] = o_cls_h[j, :m, u * bbox_elem:(u + 1) * bbox_elem]                              # (:37-42)
        all_uncertainty[j] = uc
    return rois[:, 1:5].detach().cpu().numpy(), all_boxes, all_uncertainty


def max_dets_filter(cls_boxes, cls_uncertainties, max_dets):
    """model/test.py:213-221: keep the rows whose score is >= the max_dets-th best."""
    if max_dets > 0 and len(cls_boxes) > max_dets:
        filter_thresh = np.sort(cls_boxes[:, -1])[-max_dets]
        keep = np.where(cls_boxes[:, -1] >= filter_thresh)[0]
        cls_boxes = cls_boxes[keep, :]
        cls_uncertainties = {k: (v[keep, :] if isinstance(v, np.ndarray) else v) for k, v in cls_uncertainties.items()}
    return cls_boxes, cls_uncertainties
