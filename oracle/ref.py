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
