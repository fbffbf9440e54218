"""Base anchor windows (lib/layer_utils/generate_anchors.py:41-105).

A x 4 numbers computed once on the host in fp64 exactly as the reference does (np.round is
half-to-even); the per-frame tiling over the feature grid is the CUDA kernel in snippets.py.
"""
import numpy as np


def _whctrs(box):
    w = box[2] - box[0] + 1.0
    h = box[3] - box[1] + 1.0
    return w, h, box[0] + 0.5 * (w - 1.0), box[1] + 0.5 * (h - 1.0)


def _mk(ws, hs, cx, cy):
    ws = np.reshape(np.asarray(ws, np.float64), (-1, 1))
    hs = np.reshape(np.asarray(hs, np.float64), (-1, 1))
    return np.hstack((cx - 0.5 * (ws - 1.0), cy - 0.5 * (hs - 1.0), cx + 0.5 * (ws - 1.0), cy + 0.5 * (hs - 1.0)))


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=2 ** np.arange(3, 6)):
    ratios = np.asarray(ratios, np.float64)
    scales = np.asarray(scales, np.float64)
    w, h, cx, cy = _whctrs(np.array([0.0, 0.0, base_size - 1.0, base_size - 1.0]))
    ws = np.round(np.sqrt(w * h / ratios))            # _ratio_enum :82-93
    hs = np.round(ws * ratios)
    rows = []
    for rb in _mk(ws, hs, cx, cy):                    # _scale_enum :96-105
        rw, rh, rcx, rcy = _whctrs(rb)
        rows.append(_mk(rw * scales, rh * scales, rcx, rcy))
    return np.vstack(rows)
