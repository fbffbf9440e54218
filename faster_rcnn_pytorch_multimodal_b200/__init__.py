"""B200-native two-stage detection glue (proposal / NMS / targets / RoIAlign / MC-dropout
reductions) behind the Python layer API of mathild7/faster_rcnn_pytorch_multimodal.

Sub-packages mirror the reference's ``lib/`` tree (``layer_utils``, ``model``, ``utils``,
``nets``) so that ``from layer_utils.proposal_layer import proposal_layer`` becomes
``from faster_rcnn_pytorch_multimodal_b200.layer_utils.proposal_layer import proposal_layer``.
All arithmetic runs in hand-written sm_100a CUDA kernels reached through the C ABI in
``include/b2d_glue.h`` (``libb2dglue.so``); there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
