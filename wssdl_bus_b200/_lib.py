"""ctypes binding of libwssdl_b200.so (the C ABI declared in include/wssdl_b200.h).

This is the only place the package touches native code.  There is no fallback: if the
shared library is missing the import fails with the build command, and every wrapper
raises when CUDA is unavailable.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwssdl_b200.so")
CSRC = os.path.join(_HERE, "csrc")

OK, EINVAL, EWORKSPACE, EALIGN, ELIMIT, EZERODIV = 0, -1, -2, -3, -4, -5
BIN_CPU_TRUNC, BIN_GPU_CEIL = 0, 1
BWD_ATOMIC, BWD_GATHER = 0, 1
NMS_GE_F64, NMS_GT_F32, NMS_CONTAIN = 0, 1, 4
IOU, IOU_UI = 0, 1
SAMPLE_RANKS, SAMPLE_PHILOX = 0, 1

_vp, _i, _f, _d, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/wssdl_b200.h one to one
SIGNATURES = {
    "wssdl_version": (_i, []),
    "wssdl_error_string": (ctypes.c_char_p, [_i]),
    "wssdl_set_tuning": (_i, [_i, _i]),
    "wssdl_get_tuning": (_i, [_i]),
    "wssdl_roi_pool_fwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "wssdl_roi_pool_fwd_plan": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "wssdl_roi_pool_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp, _sz,
                                _vp]),
    "wssdl_roi_pool_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp, _vp]),
    "wssdl_nms_workspace_bytes": (_sz, [_i]),
    "wssdl_nms": (_i, [_vp, _i, _i, _d, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "wssdl_gpu_nms_host": (_i, [_vp, _vp, _vp, _i, _i, _f, _i]),
    "wssdl_nms_host": (_i, [_vp, _vp, _vp, _i, _i, _d, _i, _i, _i]),
    "wssdl_bbox_overlaps_f64": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp]),
    "wssdl_bbox_overlaps_f32": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp]),
    "wssdl_bbox_transform_inv": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "wssdl_clip_boxes": (_i, [_vp, _i, _i, _f, _f, _vp]),
    "wssdl_bbox_transform": (_i, [_vp, _vp, _i, _vp, _vp]),
    "wssdl_proposals_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "wssdl_proposals": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _d, _i, _f,
                             _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "wssdl_detect_postprocess": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _f, _d, _i, _i, _vp, _vp,
                                      _vp, _vp, _vp]),
    "wssdl_eval_match": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _d, _f, _vp, _vp, _vp, _vp, _vp,
                              _vp]),
    "wssdl_anchor_labels_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "wssdl_anchor_label_counts": (_i, [_vp, _i, _i, _vp, _vp]),
    "wssdl_anchor_targets": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp,
                                  ctypes.c_ulonglong, _vp, _d, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wssdl_hot_path_fwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "wssdl_hot_path_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _d, _i, _f,
                                _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "wssdl_hot_path_proposals": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _d, _i, _f,
                                      _vp, _vp, _vp, _vp]),
    "wssdl_roi_pool_fwd_grouped": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp,
                                        _sz, _vp]),
    "wssdl_roi_targets_workspace_bytes": (_sz, [_i, _i, _i]),
    "wssdl_roi_match": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _d, _d, _d, _vp, _sz, _vp, _vp]),
    "wssdl_roi_targets": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i,
                               ctypes.c_ulonglong, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wssdl_anchor_labels": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _d, _d, _i,
                                 _vp, _vp, _vp, _vp, _sz, _vp]),
}


class WssdlError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib().wssdl_error_string(code)
        super().__init__("%s failed: %s (code %d)" % (where, msg.decode() if msg else "?", code))


def build(verbose=False):
    """Compile every CUDA source for sm_100a into libwssdl_b200.so (nvcc, in-tree)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libwssdl_b200.so failed")
    global _lib
    _lib = None
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                "%s is missing: the CUDA extension is not built. Run `make -C %s` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "There is no CPU fallback." % (LIB_PATH, CSRC))
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)       # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


# wssdl_set_tuning keys / values (include/wssdl_b200.h)
TUNE_KEYS = {"roi_fwd_kernel": 0, "roi_fwd_slices": 1, "roi_fwd_chunks": 2, "nms_sweep_cluster": 3,
             "proposals_cluster": 4, "roi_fwd_threads": 5, "pdl": 6,
             "roi_fwd_balanced": 7}
ROI_FWD_KERNELS = {"auto": 0, "direct": 1, "tiled": 2, "band": 3, "sorted": 4}


def set_tuning(key, value):
    """Process-wide tuning switch (tests, experiments); returns the previous value.
    key: a name of TUNE_KEYS; value: int, or a kernel name for 'roi_fwd_kernel'."""
    k = TUNE_KEYS[key]
    if isinstance(value, str):
        value = ROI_FWD_KERNELS[value]
    prev = lib().wssdl_get_tuning(k)
    check(lib().wssdl_set_tuning(k, int(value)), "wssdl_set_tuning")
    return prev


def check(code, where):
    if code != OK:
        raise WssdlError(code, where)
