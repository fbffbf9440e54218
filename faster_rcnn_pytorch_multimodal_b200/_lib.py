"""ctypes binding of libb2dglue.so (the C ABI declared in include/b2d_glue.h).

There is no CPU fallback: if the library is missing or fails to load, importing the
product path raises.  ``PROTOTYPES`` is the single source of truth for the Python side
and is checked against the header by tests/test_abi.py.
"""
import ctypes as C
import os
import threading

import torch  # noqa: F401  (loads libcudart before our library resolves it)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2D_LIB_PATH") or os.path.join(_HERE, "libb2dglue.so")   # override: A/B builds only

_f = C.c_void_p      # device / host float*
_i = C.c_void_p      # int32_t* / int64_t*
_v = C.c_void_p
I, F32, F64, SZ, U64 = C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_uint64

# name -> (restype, argtypes)
PROTOTYPES = {
    "b2d_abi_version": (I, []),
    "b2d_status_string": (C.c_char_p, [I]),
    "b2d_last_cuda_error": (I, []),
    "b2d_launch_count": (U64, []),
    "b2d_max_pre_nms": (I, []),
    "b2d_proposal_workspace_bytes": (SZ, [I, I, I, I, I]),
    "b2d_proposal": (I, [I, I, I, _f, _f, _f, _f, _f, I, I, F64, I, _f, _f, _f, _i, _i, _v, SZ, _v]),
    "b2d_proposal_top": (I, [I, I, I, _f, _f, _f, _f, I, I, _f, _f, _f, _v, SZ, _v]),
    "b2d_proposal_debug_sorted": (I, [I, I, I, I, I, _v, _f, _f, _i, _v]),
    "b2d_nms_sorted": (I, [I, I, _f, _i, F64, I, _i, _i, _v]),
    "b2d_argsort_desc": (I, [I, I, _f, _i, _v]),
    "b2d_roi_align_workspace_bytes": (SZ, [I, I, I, I, I, I]),
    "b2d_roi_align_forward": (I, [I, I, I, I, _f, _f, I, _i, I, _i, I, I, I, F32, I, I, _f, _v, SZ, _v]),
    "b2d_roi_align_backward": (I, [I, I, I, I, _f, _f, I, _i, I, _i, I, I, I, F32, I, I, I, _f, _v, SZ, _v]),
    "b2d_roi_align_forward_levels": (I, [I, I, I, _v, _i, _i, _v, _f, _i, I, I, I, I, I, _f, _v]),
    "b2d_fpn_level_map": (I, [I, _f, I, I, F32, I, F32, _i, _v]),
    "b2d_bbox_overlaps": (I, [I, I, _f, I, _f, I, _f, _v]),
    "b2d_bbox_transform": (I, [I, _f, I, _f, I, _f, _v]),
    "b2d_bbox_transform_inv": (I, [I, I, _f, I, _f, I, F32, I, _f, _f, _v]),
    "b2d_clip_boxes": (I, [I, I, _f, _f, _f, _v]),
    "b2d_lidar_bbox_transform": (I, [I, _f, I, _f, _f, I, _f, _v]),
    "b2d_lidar_bbox_transform_inv": (I, [I, I, _f, I, _f, _f, I, _f, _v]),
    "b2d_bbaa": (I, [I, _f, I, F32, F32, _f, _v]),
    "b2d_generate_anchors": (I, [I, I, I, I, C.POINTER(C.c_double), _f, _v]),
    "b2d_anchor_target_workspace_bytes": (SZ, [I, I, I]),
    "b2d_anchor_target_phase1": (I, [I, I, I, _f, _f, _i, _f, F32, F32, I, _i, _v, SZ, _v]),
    "b2d_anchor_target_phase2": (I, [I, I, I, I, I, I, _f, _f, _i, _i, _i, _i, _i, I, _f, F32, _f, _f, _f, _f,
                                     _v, SZ, _v]),
    "b2d_proposal_target_phase1": (I, [I, I, _f, _f, F32, F32, F32, I, _f, _i, _i, _i, _i, _v]),
    "b2d_proposal_target_phase2": (I, [I, I, _i, _f, _f, _f, _f, _f, _i, I, I, I, _f, _f, _f, _f, _f, _f, _f, _f,
                                       _f, _v]),
    "b2d_mc_variance": (I, [I, I, _f, I, _f, _v]),
    "b2d_mc_class_uncertainty": (I, [I, I, I, _f, _f, _f, _v]),
    "b2d_var_sort": (I, [I, I, _f, I, _f, _i, _v]),
    "b2d_bev_workspace_bytes": (SZ, [I, I, I, I]),
    "b2d_bev_rasterize": (I, [I, I, _f, F32, F32, F32, F32, F32, F32, F32, F32, I, I, I, I, I, I, I, _f, _i, _v, SZ, _v]),
    "b2d_final_detections": (I, [I, I, I, I, _f, _f, _i, _f, I, F32, F64, I, I, _f, I, _f, I, _f, _i, _f, _f, _i, _v]),
    "b2d_pipeline_device_bytes": (SZ, [I, I, I, I, I, I, I, I, I]),
    "b2d_proposal_crop_host": (I, [I, I, I, I, I, I, _f, _f, _f, _f, _f, I, I, F64, I, F32, I, _f, _f, _i, _f,
                                   _v, SZ, _v]),
}

_lib = None
_lock = threading.Lock()


class B2DError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library is absent."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise B2DError(
                        f"{LIB_PATH} not found: build it with `python -m faster_rcnn_pytorch_multimodal_b200.build` "
                        "(there is no CPU fallback)")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in PROTOTYPES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        L = lib()
        msg = L.b2d_status_string(rc).decode()
        if rc == -3:
            msg += f" (cudaError {L.b2d_last_cuda_error()})"
        raise B2DError(f"{what or 'b2d call'} failed: {msg}")


def ptr(t):
    """Device (or pinned host) pointer of a tensor, None -> NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise B2DError("b2d glue runs on CUDA tensors only (no CPU fallback); got a CPU tensor")


def f32c(t):
    """contiguous fp32 view/copy on the same device."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class _Workspaces:
    """Per-(device, tag) cached byte buffers (the caller owns all memory the library uses)."""

    def __init__(self):
        self._bufs = {}

    def get(self, device, tag, nbytes):
        key = (str(device), tag)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf


workspaces = _Workspaces()
