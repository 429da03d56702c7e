"""Generate the golden fixtures in tests/golden/ FROM THE REFERENCE ITSELF.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py
Sources of truth:
  * oracle/_ref  -- the reference's Cython modules (cpu_nms.pyx, utils/nms.pyx,
                    utils/bbox.pyx, utils/bbox_ui.pyx) compiled by oracle/build_ref.py;
  * /root/reference/code/lib/fast_rcnn/bbox_transform.py loaded by file path (unmodified);
  * the anchor table in the header comment of rpn_msr/generate_anchors.py:17-25 (1-based
    MATLAB boxes; the function returns table - 1), typed in below.
The RoiPool CPU op is pinned differently: oracle/_ref/ref_roi_pool.so is the reference's
roi_pooling_op.cc compiled against oracle/tf_stub and is compared live in tests/test_oracle.py
(it travels to the GPU box with the snapshot); the hand-derived bin tables of SURVEY.md
Appendix A.1 are a second, independent check.
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from wssdl_bus_b200 import synthetic as syn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = "/root/reference/code/lib"


def load_by_path(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_LIB, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    assert ref.available() or True
    out = {}
    # ---- NMS: uniform and clustered boxes, three thresholds, both reference twins + nms_new
    for tag, kw in (("uni", {}), ("clu", {"clustered": True})):
        for seed, n in ((0, 257), (1, 1000)):
            d = syn.dets(seed, n, **kw)
            key = "nms_%s_%d" % (tag, n)
            out[key + "_dets"] = d
            for t in (0.3, 0.5, 0.7):
                k = ref.cpu_nms(d, t)
                assert k == ref.cython_nms(d, t)
                out[key + "_keep_%02d" % int(t * 10)] = np.asarray(k, np.int32)
            out[key + "_keepnew_05"] = np.asarray(ref.nms_new(d, 0.5), np.int32)
    # the fp32-vs-double threshold fork: pairs whose IoU is exactly
    # the float nearest to 0.7 / 0.3
    d = np.array([[0, 0, 9, 9, 0.9], [0, 0, 9, 6, 0.8],      # iou = 70/100 -> 0.7f
                  [100, 100, 109, 109, 0.7], [100, 100, 109, 102, 0.6]], np.float32)  # 30/100
    out["nms_fork_dets"] = d
    for t in (0.3, 0.5, 0.7):
        out["nms_fork_keep_%02d" % int(t * 10)] = np.asarray(ref.cpu_nms(d, t), np.int32)
    # ---- IoU matrices
    b = syn.random_boxes(3, 300).astype(np.float64)
    q = syn.random_boxes(4, 20, lo=40, hi=300).astype(np.float64)
    out["iou_boxes"], out["iou_query"] = b, q
    out["iou_out"] = ref.bbox_overlaps(b, q)
    out["iou_ui_out"] = ref.bbox_overlaps_ui(b, q)
    # ---- bbox_transform.py (unmodified reference module)
    bt = load_by_path("fast_rcnn/bbox_transform.py", "ref_bbox_transform")
    rng = np.random.default_rng(5)
    boxes = syn.random_boxes(6, 400).astype(np.float64)
    deltas = (rng.standard_normal((400, 12)) * 0.5).astype(np.float32)
    inv = bt.bbox_transform_inv(boxes, deltas)
    out["bt_boxes"], out["bt_deltas"], out["bt_inv"] = boxes, deltas, inv
    out["bt_clip"] = bt.clip_boxes(inv.copy(), np.array([600, 800], np.float32))
    ex = syn.random_boxes(7, 200)
    gt = syn.random_boxes(8, 200)
    out["bt_ex"], out["bt_gt"] = ex, gt
    out["bt_targets"] = bt.bbox_transform(ex, gt).astype(np.float32)
    # ---- anchors: generate_anchors.py:17-25 table (1-based) -> function output = table - 1
    table = np.array([[-83, -39, 100, 56], [-175, -87, 192, 104], [-359, -183, 376, 200],
                      [-55, -55, 72, 72], [-119, -119, 136, 136], [-247, -247, 264, 264],
                      [-35, -79, 52, 96], [-79, -167, 96, 184], [-167, -343, 184, 360]],
                     np.float64)
    out["anchors_table_minus_1"] = table - 1
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
