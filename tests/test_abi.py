"""The C-ABI shared library loads and exports every symbol include/*.h declares (no GPU,
no compute calls), and the Python mirror of the reference's module tree imports."""
import ctypes
import glob
import importlib
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(wssdl_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


@pytest.fixture(scope="module")
def built_lib():
    from wssdl_bus_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return _lib


def test_header_declares_expected_entry_points():
    names = _declared_symbols()
    for must in ("wssdl_roi_pool_fwd", "wssdl_roi_pool_bwd", "wssdl_nms", "wssdl_nms_host",
                 "wssdl_gpu_nms_host", "wssdl_bbox_overlaps_f64", "wssdl_proposals",
                 "wssdl_anchor_labels", "wssdl_bbox_transform_inv", "wssdl_clip_boxes"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    L = ctypes.CDLL(built_lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(L, name), "libwssdl_b200.so lacks %s" % name


def test_binding_table_matches_header(built_lib):
    assert sorted(built_lib.SIGNATURES) == _declared_symbols()
    lib = built_lib.lib()
    assert lib.wssdl_version() >= 100
    assert lib.wssdl_error_string(-4) == b"size beyond kernel limits"
    # pure host-side queries are safe without a GPU
    assert lib.wssdl_nms_workspace_bytes(6000) >= 6000 * 94 * 8
    assert lib.wssdl_nms_workspace_bytes(0) > 0


def test_tuning_keys_match_the_header(built_lib):
    """The tuning ABI: the Python names map onto the header's enum values, every key can be
    read and written without a GPU, unknown keys are rejected."""
    text = open(os.path.join(ROOT, "include", "wssdl_b200.h")).read()
    enum = dict((k.lower(), int(v)) for k, v in re.findall(r"WSSDL_TUNE_([A-Z_]+)\s*=\s*(\d+)", text))
    count = enum.pop("count")
    assert enum == built_lib.TUNE_KEYS and count == len(enum)
    lib = built_lib.lib()
    for name, key in enum.items():
        prev = lib.wssdl_get_tuning(key)
        assert lib.wssdl_set_tuning(key, prev) == built_lib.OK, name
    assert lib.wssdl_set_tuning(count, 0) == built_lib.EINVAL
    assert lib.wssdl_set_tuning(-1, 0) == built_lib.EINVAL
    # the hot-path entries validate their sizes before any CUDA call
    assert lib.wssdl_hot_path_fwd_workspace_bytes(256, 300, 7, 7) >= lib.wssdl_roi_pool_fwd_workspace_bytes(
        256, 256 * 300, 7, 7)
    args = [None] * 4 + [3, 1, 38, 50, 512, 9, None, 16, 6000]
    assert lib.wssdl_hot_path_fwd(*args, 0, 0.7, 0, 16.0, 7, 7, 0.0625, 0, None, None, None, None, None,
                                  None, 0, None, None) == built_lib.EINVAL      # post_nms_topN <= 0
    assert lib.wssdl_roi_pool_fwd_grouped(None, None, -1, 1, 2, 2, 4, 7, 7, 0.0625, 0, None, None, None, 0,
                                          None) == built_lib.EINVAL


def test_argument_validation_without_gpu(built_lib):
    lib = built_lib.lib()
    # negative sizes / bad enums are rejected before any CUDA call
    assert lib.wssdl_roi_pool_fwd(None, None, 1, 2, 2, 4, -1, 7, 7, 0.0625, 0, None, None, None, 0, None) == built_lib.EINVAL
    assert lib.wssdl_roi_pool_fwd(None, None, 1, 2, 2, 4, 1, 7, 7, 0.0625, 9, None, None, None, 0, None) == built_lib.EINVAL
    assert lib.wssdl_roi_pool_fwd(None, None, 1, 2, 2, 4, 0, 7, 7, 0.0625, 0, None, None, None, 0, None) == built_lib.OK
    assert lib.wssdl_bbox_overlaps_f64(None, 0, None, 5, 0, None, None) == built_lib.OK
    assert lib.wssdl_bbox_overlaps_f64(None, 3, None, 5, 7, None, None) == built_lib.EINVAL


def test_missing_library_fails_loudly(monkeypatch, built_lib):
    monkeypatch.setattr(built_lib, "LIB_PATH", "/nonexistent/libwssdl_b200.so")
    monkeypatch.setattr(built_lib, "_lib", None)
    with pytest.raises(ImportError, match="no CPU fallback"):
        built_lib.lib()


def test_no_cuda_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    import wssdl_bus_b200 as w
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        w.nms(np.zeros((3, 5), np.float32), 0.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        w.roi_pool(np.zeros((1, 2, 2, 4), np.float32), np.zeros((1, 5), np.float32), 2, 2, 1.0)


@pytest.mark.parametrize("mod,attrs", [
    ("roi_pooling_layer.roi_pooling_op", ["roi_pool", "roi_pool_grad"]),
    ("roi_pooling_layer.roi_pooling_op_grad", []),
    ("nms.cpu_nms", ["cpu_nms"]), ("nms.gpu_nms", ["gpu_nms"]), ("nms.py_cpu_nms", ["py_cpu_nms"]),
    ("utils.cython_bbox", ["bbox_overlaps"]), ("utils.cython_bbox_ui", ["bbox_overlaps_ui"]),
    ("utils.cython_nms", ["nms", "nms_new"]),
    ("fast_rcnn.nms_wrapper", ["nms"]),
    ("fast_rcnn.bbox_transform", ["bbox_transform", "bbox_transform_inv", "clip_boxes"]),
    ("fast_rcnn.config", ["cfg"]),
    ("rpn_msr.generate_anchors", ["generate_anchors"]),
    ("rpn_msr.proposal_layer_tf_bus", ["proposal_layer"]),
    ("rpn_msr.anchor_target_layer_tf_bus", ["anchor_target_layer", "anchor_target_layer_ws",
                                            "anchor_target_layer_joint"]),
    ("rpn_msr.proposal_target_layer_tf_bus", ["proposal_target_layer", "proposal_target_layer_joint"]),
])
def test_reference_module_tree_is_mirrored(mod, attrs):
    m = importlib.import_module("wssdl_bus_b200." + mod)
    for a in attrs:
        assert hasattr(m, a)


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "wssdl_bus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f
