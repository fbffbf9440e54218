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
    "b2d_nms_workspace_bytes": (SZ, [I, I]),
    "b2d_nms_sorted": (I, [I, I, _f, _i, F64, I, _i, _i, _v, SZ, _v]),
    "b2d_argsort_workspace_bytes": (SZ, [I, I]),
    "b2d_argsort_desc": (I, [I, I, _f, _i, _v, SZ, _v]),
    "b2d_roi_align_workspace_bytes": (SZ, [I, I, I, I, I, I]),
    "b2d_roi_align_forward": (I, [I, I, I, I, _f, _f, I, _i, I, _i, I, I, I, F32, I, I, _f, _v, SZ, _v]),
    "b2d_roi_align_forward_route": (I, [I, I, I, I, _f, _f, I, _i, I, _i, I, I, I, F32, I, I, I, _f, _v, SZ, _v]),
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
    "b2d_bbox_voxel_grid_to_pc": (I, [I, I, F32, F32, F32, F32, I, _f, _v]),
    "b2d_eval_match": (I, [I, I, I, I, _f, _i, _i, _i, _f, _i, _i, _f, F64, F64, I, _i, _f, _i, _i, _v, _v]),
    "b2d_head_tail_decode": (I, [I, I, I, I, I, _f, _f, _f, _f, _f, _f, C.POINTER(C.c_float), C.POINTER(C.c_float), I, I, I,
                                 _f, _f, _f, _f, _f, _f, _v]),
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


def _load():
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


class _OnDevice:
    """The library handle with every call made under ``torch.cuda.device(device)``.

    The C ABI launches on the calling thread's CURRENT device (kernel launches, cudaFuncSetAttribute,
    cudaMemsetAsync), while the stream and the pointers it is given belong to the tensors' device; torchvision's
    ops switch devices with a guard, and so does every wrapper in this package."""

    __slots__ = ("_h", "_dev")

    def __init__(self, handle, device):
        self._h, self._dev = handle, device

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        dev = self._dev

        def guarded(*args):
            with torch.cuda.device(dev):
                return fn(*args)

        return guarded


def lib(device=None):
    """Load (once) and return the ctypes handle; raises if the CUDA library is absent.  With ``device`` (a CUDA
    torch.device) the returned handle makes that device current around every call."""
    handle = _load()
    if device is None:
        return handle
    device = torch.device(device)
    if device.type != "cuda":
        raise B2DError("b2d glue runs on CUDA devices only (no CPU fallback)")
    if device.index is None or device.index == torch.cuda.current_device():
        return handle
    return _OnDevice(handle, device)


def check(rc: int, what: str = ""):
    if rc != 0:
        L = _load()
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
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise B2DError("b2d glue runs on CUDA tensors only (no CPU fallback); got a CPU tensor")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise B2DError(f"b2d glue: tensors on different devices ({dev} and {t.device})")


def f32c(t):
    """contiguous fp32 view/copy on the same device."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def i32c(t, what="index tensor"):
    """contiguous int32 view/copy of an integer tensor (the kernels read int32); None passes through."""
    if t is None:
        return None
    if t.dtype == torch.int32:
        return t if t.is_contiguous() else t.contiguous()
    if t.dtype in (torch.int64, torch.int16, torch.int8, torch.uint8):
        return t.to(torch.int32).contiguous()
    raise B2DError(f"{what} must be an integer tensor, got {t.dtype}")


class _Workspaces:
    """Cached scratch buffers, one per (device, stream, tag): the caller owns all memory the library uses.

    Keyed by the CURRENT stream so that calls overlapped on two streams (camera and lidar branches) never
    share scratch; a buffer is only ever used on the stream it was allocated on, so the caching allocator's
    stream-ordered reuse rules hold without record_stream()."""

    def __init__(self):
        self._bufs = {}
        self._lock = threading.Lock()

    def get(self, device, tag, nbytes):
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, torch.cuda.current_stream(device).cuda_stream, tag)
        with self._lock:
            buf = self._bufs.get(key)
            if buf is None or buf.numel() < nbytes:
                with torch.cuda.device(idx):
                    buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
                self._bufs[key] = buf
        return buf


workspaces = _Workspaces()
