"""Loader for the reference's own Cython modules compiled into oracle/_ref.

TEST INFRASTRUCTURE ONLY.  The binaries are produced by oracle/build_ref.py from the
sources under /root/reference (which exists only in the authoring container); the
built .so files travel to the GPU box with the repo snapshot, the sources do not.
"""
import importlib
import os
import sys

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mods = {}


def available():
    from . import build_ref
    return build_ref.built()


def _load(name):
    if name not in _mods:
        if not available():
            from . import build_ref
            if not build_ref.build():
                raise ImportError("oracle/_ref is not built and /root/reference is absent")
        # the reference evaluates np.float / np.int at run time (cpu_nms.pyx:17,29; bbox.pyx:12)
        if not hasattr(np, "float"):
            np.float = float
        if not hasattr(np, "int"):
            np.int = int
        if _DIR not in sys.path:
            sys.path.insert(0, _DIR)
        _mods[name] = importlib.import_module(name)
    return _mods[name]


def cpu_nms(dets, thresh):
    """nms/cpu_nms.pyx:17 -- the live NMS path (cfg.USE_GPU_NMS is False)."""
    return _load("ref_cpu_nms").cpu_nms(np.ascontiguousarray(dets, np.float32), float(thresh))


def cython_nms(dets, thresh):
    """utils/nms.pyx:17 -- used by the per-class test loops (test_bus.py:366)."""
    return _load("ref_cython_nms").nms(np.ascontiguousarray(dets, np.float32), float(thresh))


def nms_new(dets, thresh):
    """utils/nms.pyx:70."""
    return _load("ref_cython_nms").nms_new(np.ascontiguousarray(dets, np.float32), float(thresh))


def bbox_overlaps(boxes, query):
    """utils/bbox.pyx:15 (float64 only)."""
    return _load("ref_cython_bbox").bbox_overlaps(
        np.ascontiguousarray(boxes, np.float64), np.ascontiguousarray(query, np.float64))


def bbox_overlaps_ui(boxes, query):
    """utils/bbox_ui.pyx:12 (float64 only)."""
    return _load("ref_cython_bbox_ui").bbox_overlaps_ui(
        np.ascontiguousarray(boxes, np.float64), np.ascontiguousarray(query, np.float64))


# ---------------------------------------------------------------- RoiPool / RoiPoolGrad
_roi_lib = None


def roi_pool_available():
    """True when the reference's own roi_pooling_op.cc has been compiled into
    oracle/_ref/ref_roi_pool.so (oracle/build_ref.py: against oracle/tf_stub)."""
    from . import build_ref
    return build_ref.roi_pool_built() or build_ref.build_roi_pool()


def _roi():
    global _roi_lib
    if _roi_lib is None:
        import ctypes
        from . import build_ref
        if not roi_pool_available():
            raise ImportError("oracle/_ref/ref_roi_pool.so is not built and /root/reference is absent")
        L = ctypes.CDLL(build_ref.ROI_POOL_SO)
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.ref_roi_pool_fwd.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, ci, cf, ci, vp, vp, ctypes.c_char_p]
        L.ref_roi_pool_bwd.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, cf, ci, vp, ctypes.c_char_p]
        _roi_lib = L
    return _roi_lib


def roi_pool_fwd(bottom, rois, pooled_h, pooled_w, spatial_scale, threads=1):
    """RoiPoolOp<CPUDevice, float>::Compute (roi_pooling_op.cc:88-204), the reference's own
    object code.  bottom [B,H,W,C] f32, rois [R,5] f32 -> (top f32, argmax i32) [R,PH,PW,C].
    Batch indices must be inside [0,B): the reference reads out of bounds otherwise."""
    import ctypes
    bottom = np.ascontiguousarray(bottom, np.float32)
    rois = np.ascontiguousarray(rois, np.float32)
    B, H, W, C = bottom.shape
    R = rois.shape[0]
    shape = (R, max(pooled_h, 0), max(pooled_w, 0), C)   # negative sizes: the op itself rejects
    top = np.empty(shape, np.float32)
    arg = np.empty(shape, np.int32)
    err = ctypes.create_string_buffer(256)
    rc = _roi().ref_roi_pool_fwd(bottom.ctypes.data, rois.ctypes.data, B, H, W, C, R, pooled_h,
                                 pooled_w, spatial_scale, int(threads), top.ctypes.data,
                                 arg.ctypes.data, err)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return top, arg


def roi_pool_bwd(grad, argmax, rois, bottom_shape, spatial_scale, threads=1):
    """RoiPoolGradOp<CPUDevice, float>::Compute (roi_pooling_op.cc:333-466) -> [B,H,W,C] f32."""
    import ctypes
    grad = np.ascontiguousarray(grad, np.float32)
    argmax = np.ascontiguousarray(argmax, np.int32)
    rois = np.ascontiguousarray(rois, np.float32)
    B, H, W, C = [int(v) for v in bottom_shape]
    R, PH, PW, _ = grad.shape
    bottom = np.zeros((B, H, W, C), np.float32)          # only its shape is read
    out = np.empty((B, H, W, C), np.float32)
    err = ctypes.create_string_buffer(256)
    rc = _roi().ref_roi_pool_bwd(bottom.ctypes.data, rois.ctypes.data, argmax.ctypes.data,
                                 grad.ctypes.data, B, H, W, C, R, PH, PW, spatial_scale,
                                 int(threads), out.ctypes.data, err)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return out


_twin_lib = None


def cuda_twin_available():
    """True when the reference's CUDA RoI-pool kernels have been built for the host
    (oracle/build_ref.py:build_cuda_twin)."""
    from . import build_ref
    return build_ref.cuda_twin_built() or build_ref.build_cuda_twin()


def roi_pool_fwd_cuda_twin(bottom, rois, pooled_h, pooled_w, spatial_scale):
    """ROIPoolForwardLaucher / ROIPoolForward (roi_pooling_op_gpu.cu.cc:19-110), the kernel body
    compiled for the host and called once per emulated CUDA thread: the GPU_CEIL bins."""
    global _twin_lib
    import ctypes
    from . import build_ref
    if _twin_lib is None:
        if not cuda_twin_available():
            raise ImportError("oracle/_ref/ref_roi_pool_cudatwin.so is not built and /root/reference is absent")
        L = ctypes.CDLL(build_ref.CUDA_TWIN_SO)
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.ref_roi_pool_fwd_cudatwin.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, cf, vp, vp]
        _twin_lib = L
    bottom = np.ascontiguousarray(bottom, np.float32)
    rois = np.ascontiguousarray(rois, np.float32)
    B, H, W, C = bottom.shape
    R = rois.shape[0]
    top = np.empty((R, pooled_h, pooled_w, C), np.float32)
    arg = np.empty((R, pooled_h, pooled_w, C), np.int32)
    rc = _twin_lib.ref_roi_pool_fwd_cudatwin(bottom.ctypes.data, rois.ctypes.data, H, W, C, R,
                                             pooled_h, pooled_w, spatial_scale, top.ctypes.data,
                                             arg.ctypes.data)
    if rc != 0:
        raise RuntimeError("ROIPoolForwardLaucher failed")
    return top, arg


_nms_lib = None


def gpu_nms_available():
    """True when nms/nms_kernel.cu has been built for the host (build_ref.build_cuda_nms)."""
    from . import build_ref
    return build_ref.cuda_nms_built() or build_ref.build_cuda_nms()


def gpu_nms(dets, thresh):
    """nms/gpu_nms.pyx:16-31 around the reference's own _nms (nms_kernel.cu:91-144, host build):
    sort by score (:25-28), _nms on the sorted boxes, map the kept positions back (:31).
    The '>' predicate in float32: `devIoU(...) > nms_overlap_thresh` (nms_kernel.cu:71)."""
    global _nms_lib
    import ctypes
    from . import build_ref
    if _nms_lib is None:
        if not gpu_nms_available():
            raise ImportError("oracle/_ref/ref_gpu_nms_hostbuild.so is not built and /root/reference is absent")
        L = ctypes.CDLL(build_ref.CUDA_NMS_SO)
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.ref_gpu_nms_sorted.argtypes = [vp, vp, vp, ci, ci, cf]
        L.ref_gpu_nms_sorted.restype = None
        _nms_lib = L
    dets = np.ascontiguousarray(dets, np.float32)
    boxes_num, boxes_dim = dets.shape
    if boxes_num == 0:
        return []
    keep = np.zeros(boxes_num, dtype=np.int32)
    num_out = ctypes.c_int(0)
    order = dets[:, 4].argsort()[::-1]
    sorted_dets = np.ascontiguousarray(dets[order, :])
    _nms_lib.ref_gpu_nms_sorted(keep.ctypes.data, ctypes.byref(num_out), sorted_dets.ctypes.data,
                                boxes_num, boxes_dim, float(thresh))
    return list(order[keep[:num_out.value]])
