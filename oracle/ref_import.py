"""Import the UNMODIFIED reference (``/root/reference/lib``) in the build container.

Used only by ``oracle/gen_golden.py`` and by the optional cross-check tests that
skip when ``/root/reference`` is absent (it never exists on the GPU box).  Two
modules the reference imports are not installed here; they are stubbed:
``easydict`` (model/config.py:9), ``matplotlib.pyplot`` (utils/loss_utils.py:7) and ``cv2``
(utils/filter_predictions.py:10, imported there but unused on the path).
"""
import os
import sys
import types

REF_LIB = "/root/reference/lib"


def available() -> bool:
    return os.path.isdir(REF_LIB)


def _install_stubs():
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")

        class EasyDict(dict):
            def __init__(self, d=None, **kw):
                super().__init__()
                for k, v in dict(d or {}, **kw).items():
                    self[k] = v

            def __setitem__(self, k, v):
                if isinstance(v, dict) and not isinstance(v, EasyDict):
                    v = EasyDict(v)
                super().__setitem__(k, v)

            def __setattr__(self, k, v):
                self[k] = v

            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError as e:
                    raise AttributeError(k) from e

        m.EasyDict = EasyDict
        sys.modules["easydict"] = m
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except ImportError:
            sys.modules["cv2"] = types.ModuleType("cv2")


def load():
    """Return a namespace of the reference's hot-path modules."""
    if not available():
        raise RuntimeError("/root/reference is not present (expected on the GPU box)")
    _install_stubs()
    if REF_LIB not in sys.path:
        sys.path.insert(0, REF_LIB)
    ns = types.SimpleNamespace()
    from model.config import cfg
    import model.bbox_transform as bt
    import utils.bbox as ub
    import layer_utils.generate_anchors as ga
    import layer_utils.snippets as sn
    import layer_utils.generate_3d_anchors as g3
    import layer_utils.proposal_layer as pl
    import layer_utils.proposal_top_layer as ptl
    import layer_utils.anchor_target_layer as atl
    import layer_utils.proposal_target_layer as prt
    import utils.loss_utils as lu
    import utils.torchpoolers as tp
    import utils.filter_predictions as fp
    ns.fp = fp
    ns.cfg, ns.bt, ns.ub, ns.ga, ns.sn, ns.g3 = cfg, bt, ub, ga, sn, g3
    ns.pl, ns.ptl, ns.atl, ns.prt, ns.lu, ns.tp = pl, ptl, atl, prt, lu, tp
    return ns
