"""ctypes view of oracle/liboracle.so (oracle/hotpath_ref.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

CPU_TRUNC = 0  # roi_pooling_op.cc:167-170 (parity target)
GPU_CEIL = 1   # roi_pooling_op_gpu.cu.cc:51-58 (differential twin)


def build(force=False):
    src = os.path.join(_HERE, "hotpath_ref.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"] + (["-B"] if force else []),
                       check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int)
        dp = ctypes.POINTER(ctypes.c_double)
        lp = ctypes.POINTER(ctypes.c_int64)
        I, F, D = ctypes.c_int, ctypes.c_float, ctypes.c_double
        L.roi_pool_fwd_ref.argtypes = [fp, fp, I, I, I, I, I, I, I, F, I, fp, ip, I]
        L.roi_pool_fwd_ref.restype = None
        for name in ("roi_pool_bwd_ref", "roi_pool_bwd_ref_fast"):
            f = getattr(L, name)
            f.argtypes = [fp, ip, fp, I, I, I, I, I, I, I, F, fp, I]
            f.restype = None
        L.nms_ref.argtypes = [fp, I, I, lp, D, I, lp]
        L.nms_ref.restype = I
        for name in ("bbox_overlaps_ref", "bbox_overlaps_ui_ref"):
            f = getattr(L, name)
            f.argtypes = [dp, I, dp, I, dp]
            f.restype = None
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def default_threads():
    return max(1, len(os.sched_getaffinity(0)))


def roi_pool_fwd(bottom, rois, pooled_h, pooled_w, spatial_scale, bin_mode=CPU_TRUNC,
                 threads=None):
    """bottom [B,H,W,C] f32, rois [R,5] f32 -> (top [R,PH,PW,C] f32, argmax i32)."""
    bottom = np.ascontiguousarray(bottom, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 5)
    B, H, W, C = bottom.shape
    R = rois.shape[0]
    top = np.empty((R, pooled_h, pooled_w, C), np.float32)
    arg = np.empty((R, pooled_h, pooled_w, C), np.int32)
    lib().roi_pool_fwd_ref(_p(bottom, ctypes.c_float), _p(rois, ctypes.c_float), B, H, W, C,
                           R, pooled_h, pooled_w, np.float32(spatial_scale), bin_mode,
                           _p(top, ctypes.c_float), _p(arg, ctypes.c_int),
                           threads or default_threads())
    return top, arg


def roi_pool_bwd(top_diff, argmax, rois, bottom_shape, spatial_scale, literal=False,
                 threads=None):
    """RoiPoolGrad: -> bottom_diff [B,H,W,C] f32."""
    top_diff = np.ascontiguousarray(top_diff, dtype=np.float32)
    argmax = np.ascontiguousarray(argmax, dtype=np.int32)
    rois = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 5)
    B, H, W, C = bottom_shape
    R, PH, PW, C2 = top_diff.shape
    assert C2 == C and argmax.shape == top_diff.shape and rois.shape[0] == R
    out = np.empty((B, H, W, C), np.float32)
    f = lib().roi_pool_bwd_ref if literal else lib().roi_pool_bwd_ref_fast
    f(_p(top_diff, ctypes.c_float), _p(argmax, ctypes.c_int), _p(rois, ctypes.c_float),
      B, H, W, C, R, PH, PW, np.float32(spatial_scale), _p(out, ctypes.c_float),
      threads or default_threads())
    return out


def nms(dets, thresh, variant=0, order=None):
    """C restatement of cpu_nms (variant 0) / nms_new (variant 1).  Returns list[int].
    Raises ZeroDivisionError like the reference when a visited pair has zero union."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    n = dets.shape[0]
    if order is None:
        order = dets[:, 4].argsort()[::-1]  # cpu_nms.pyx:25
    order = np.ascontiguousarray(order, dtype=np.int64)
    keep = np.empty(max(n, 1), np.int64)
    k = lib().nms_ref(_p(dets, ctypes.c_float), n, dets.shape[1], _p(order, ctypes.c_int64),
                      float(thresh), variant, _p(keep, ctypes.c_int64))
    if k < 0:
        raise ZeroDivisionError("float division")
    return keep[:k].tolist()


def bbox_overlaps(boxes, query, ui=False):
    boxes = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 4)
    query = np.ascontiguousarray(query, dtype=np.float64).reshape(-1, 4)
    out = np.empty((boxes.shape[0], query.shape[0]), np.float64)
    f = lib().bbox_overlaps_ui_ref if ui else lib().bbox_overlaps_ref
    f(_p(boxes, ctypes.c_double), boxes.shape[0], _p(query, ctypes.c_double), query.shape[0],
      _p(out, ctypes.c_double))
    return out
