"""Result files of a test run: what ``model/test.py:213-254`` and ``datasets/db.py:305-367`` do
once the detections of every frame are known.

    records  = stream.gather_detections(stream.pack_detection_records(...))        # §8e wire format
    all_boxes = stream.unpack_detection_records(records, num_classes, num_frames)   # all_boxes[cls][frame]
    dump_detections(all_boxes, output_dir)                                          # detections.pkl
    write_image_results_file(all_boxes, classes, frame_tokens, output_dir, 'test')  # results/det_test_<cls>.txt

The per-frame arithmetic stays on the device (``bbox_voxel_grid_to_pc`` below, the stacking of box | score |
uncertainties in ``ops.final_detections``' padded records); formatting text is host work in the reference and here.
"""
import os
import pickle
from typing import Sequence

import numpy as np
import torch

from .._lib import check, lib, ptr, require_cuda, stream_ptr
from ..model.config import cfg


def bbox_voxel_grid_to_pc(bboxes: torch.Tensor, bev_extants, info, aabb: bool = False) -> torch.Tensor:
    """utils/bbox.py:140-162 on a CUDA tensor [n, >=5] (or [.., D, width] padded records), IN PLACE as the reference:
    voxel-grid x, y (and l, w - or x2, y2 when ``aabb``) back to point-cloud metres.  The scale factors are computed
    on the host in fp32 exactly as numpy computes them."""
    require_cuda(bboxes)
    if not bboxes.is_contiguous() or bboxes.dtype != torch.float32:
        raise ValueError("bbox_voxel_grid_to_pc works in place on a contiguous fp32 tensor")
    info = np.asarray(info.detach().cpu().numpy() if isinstance(info, torch.Tensor) else info, dtype=np.float32)
    bev_extants = [float(x) for x in bev_extants]          # python floats: weak scalars, the arithmetic stays in fp32
    scale = info[6]
    s_info = np.asarray(info[0:6]) * 1 / scale
    fx = (bev_extants[3] - bev_extants[0]) / (s_info[1] - s_info[0])
    fy = (bev_extants[4] - bev_extants[1]) / (s_info[3] - s_info[2])
    width = bboxes.shape[-1]
    n = bboxes.numel() // width
    check(lib(bboxes.device).b2d_bbox_voxel_grid_to_pc(n, width, float(np.float32(fx)), float(np.float32(fy)),
                                                       float(np.float32(bev_extants[0])), float(np.float32(bev_extants[1])),
                                                       int(bool(aabb)), ptr(bboxes), stream_ptr(bboxes.device)),
          "b2d_bbox_voxel_grid_to_pc")
    return bboxes


def stack_uncertainties(cls_bbox: np.ndarray, cls_uncertainties: dict, num_uc_pos: int) -> np.ndarray:
    """model/test.py:260-270 for callers that hold the reference's per-class dicts (the device path delivers the
    stacked rows directly: ops.final_detections gathers the uncertainty columns next to box and score)."""
    out = np.zeros((cls_bbox.shape[0], cls_bbox.shape[1] + num_uc_pos))
    out[:, 0:cls_bbox.shape[1]] = cls_bbox
    p = cls_bbox.shape[1]
    for _, val in cls_uncertainties.items():
        out[:, p:p + val.shape[1]] = val[:, :]
        p += val.shape[1]
    return out


def dump_detections(all_boxes, output_dir: str) -> str:
    """model/test.py:252-254."""
    det_file = os.path.join(output_dir, 'detections.pkl')
    with open(det_file, 'wb') as f:
        pickle.dump(all_boxes, f, pickle.HIGHEST_PROTOCOL)
    return det_file


def get_results_file_template(mode: str, class_name: str, output_dir: str) -> str:
    """datasets/db.py:130-137."""
    result_dir = os.path.join(output_dir, 'results')
    if not os.path.isdir(result_dir):
        os.mkdir(result_dir)
    return os.path.join(result_dir, 'det_' + mode + '_{:s}.txt'.format(class_name))


def _write(all_boxes, classes: Sequence[str], frame_tokens: Sequence[str], output_dir: str, mode: str, line):
    paths = []
    for cls_ind, cls in enumerate(classes):
        if cls == 'dontcare' or cls == '__background__':
            continue
        filename = get_results_file_template(mode, cls, output_dir)
        with open(filename, 'wt') as f:
            for ind, token in enumerate(frame_tokens):
                dets = all_boxes[cls_ind][ind]
                if dets.size == 0:
                    continue
                for k in range(dets.shape[0]):
                    f.write(line(ind, token, dets, k))
                    f.write('\n')
        paths.append(filename)
    return paths


def write_image_results_file(all_boxes, classes, frame_tokens, output_dir, mode):
    """datasets/db.py:305-332: `<frame idx> <token> <score> <x1> <y1> <x2> <y2> [uncertainties ...]`."""
    def line(ind, token, dets, k):
        s = '{:d} {:s} {:.3f} {:.1f} {:.1f} {:.1f} {:.1f}'.format(ind, token, dets[k, 4], dets[k, 0], dets[k, 1], dets[k, 2],
                                                                  dets[k, 3])
        return s + ''.join(' {:.10f}'.format(dets[k, l]) for l in range(5, dets.shape[1]))
    return _write(all_boxes, classes, frame_tokens, output_dir, mode, line)


def write_lidar_results_file(all_boxes, classes, frame_tokens, output_dir, mode):
    """datasets/db.py:334-367: `<frame idx> <token> <score> <xc> <yc> <zc> <l> <w> <h> <ry> [uncertainties ...]`."""
    nbe = int(cfg.LIDAR.NUM_BBOX_ELEM)

    def line(ind, token, dets, k):
        s = '{:d} {:s} {:.3f} {:.3f} {:.3f} {:.3f} {:.3f} {:.3f} {:.3f} {:.5f}'.format(
            ind, token, dets[k, 7], dets[k, 0], dets[k, 1], dets[k, 2], dets[k, 3], dets[k, 4], dets[k, 5], dets[k, 6])
        if dets.shape[1] > nbe + 1:
            s += ''.join(' {:.10f}'.format(dets[k, l]) for l in range(8, dets.shape[1]))
        return s
    return _write(all_boxes, classes, frame_tokens, output_dir, mode, line)
