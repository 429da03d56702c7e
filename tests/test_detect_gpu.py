"""Per-class detection post-processing (fast_rcnn/test_bus.py:207-223, :360-401) on the GPU
against the numpy restatement composed from the pinned oracles (oracle.layers)."""
import numpy as np
import pytest
import torch

from wssdl_bus_b200 import ops, synthetic as syn

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # regressed boxes: exp() is 1 ulp away from numpy's


def _inputs(seed, B, S, K, counts=None, im_hw=(437, 583), im_scale=600.0 / 437):
    """RoIs like the proposal layer emits them (scaled frame), fixed stride S per image."""
    rois = np.zeros((B * S, 5), np.float32)
    for b in range(B):
        r = syn.rois_for_pool(seed + b, S, 1, im_w=int(im_hw[1] * im_scale), im_h=int(im_hw[0] * im_scale))
        r[:, 0] = b
        rois[b * S:(b + 1) * S] = r
    scores, deltas = syn.rcnn_head_outputs(seed, B * S, K)
    meta = np.tile(np.array([[im_hw[0], im_hw[1], im_scale]], np.float32), (B, 1))
    return rois, scores, deltas, meta


def _oracle(oracle_mod, rois, scores, deltas, meta, S, counts, pred_boxes, **kw):
    """Reference flow per image.  The discrete steps are fed the DEVICE-regressed boxes so
    they can be compared bit-exactly (exp ulp differences would otherwise flip NMS ties)."""
    B = meta.shape[0]
    outs = []
    for b in range(B):
        n = S if counts is None else int(counts[b])
        sl = slice(b * S, b * S + n)
        outs.append(oracle_mod.layers.detections_postprocess(scores[sl], pred_boxes[sl], **kw))
    return outs


@pytest.mark.parametrize("B,S,K,agn,cap", [(1, 300, 3, False, 300), (4, 300, 3, False, 40),
                                           (3, 300, 3, True, 300), (2, 128, 5, True, 25),
                                           (2, 1000, 2, False, 100), (5, 64, 21, False, 0)])
def test_postprocess_matches_reference_flow(oracle_mod, B, S, K, agn, cap):
    rois, scores, deltas, meta = _inputs(700 + S + K, B, S, K)
    counts = None
    if B > 1:
        counts = np.full((B,), S, np.int32)
        counts[1] = S // 3
        counts[-1] = 0 if B > 2 else S - 1
    out = ops.detect_postprocess(rois, scores, deltas, meta, roi_counts=counts, roi_stride=S,
                                 score_thresh=0.05, nms_thresh=0.3, max_per_image=cap,
                                 cls_agnostic=agn, want_pred_boxes=True)
    pred = out["pred_boxes"].cpu().numpy()
    # 1. regressed + clipped boxes vs the reference arithmetic (tolerance: exp)
    for b in range(B):
        n = S if counts is None else int(counts[b])
        sl = slice(b * S, b * S + n)
        want = oracle_mod.layers.im_detect_boxes(rois[sl], deltas[sl], (meta[b, 0], meta[b, 1]),
                                                 float(meta[b, 2]))
        np.testing.assert_allclose(pred[sl], want, rtol=RTOL, atol=1e-3)
    # 2. threshold / NMS / agnostic NMS / cap: bit-exact given the same boxes
    want = _oracle(oracle_mod, rois, scores, deltas, meta, S, counts, pred, thresh=np.float32(0.05),
                   nms_thresh=0.3, max_per_image=cap, cls_agnostic_nms=agn)
    dets = out["dets"].cpu().numpy()
    cnt = out["counts"].cpu().numpy()
    assert int(out["status"].item()) == 0
    for b in range(B):
        assert cnt[b, 0] == 0
        for j in range(1, K):
            got = dets[b, j, :cnt[b, j]]
            assert got.shape == want[b][j].shape, (b, j, got.shape, want[b][j].shape)
            assert np.array_equal(got, want[b][j])
            assert not dets[b, j, cnt[b, j]:].any()          # padding is zero-filled


def test_reference_named_wrappers(oracle_mod):
    from wssdl_bus_b200.fast_rcnn import test_bus
    rois, scores, deltas, meta = _inputs(900, 2, 300, 3)
    all_boxes, out = test_bus.test_net_batch(rois, scores, deltas, meta[:, :2], meta[:, 2], roi_stride=300)
    assert len(all_boxes) == 3 and len(all_boxes[1]) == 2 and all_boxes[0][0].shape == (0, 5)
    pb = test_bus.detect_boxes(rois[:300], deltas[:300], (437, 583, 3), float(meta[0, 2]))
    want = oracle_mod.layers.detections_postprocess(scores[:300], pb, thresh=np.float32(0.05))
    got = test_bus.postprocess_detections(scores[:300], pb)
    for j in range(3):
        assert np.array_equal(got[j], want[j]) and np.array_equal(all_boxes[j][0], want[j])


def test_limits_and_empty():
    from wssdl_bus_b200 import _lib
    rois, scores, deltas, meta = _inputs(901, 1, 8, 3)
    out = ops.detect_postprocess(rois, scores, deltas, meta, score_thresh=2.0)   # nothing passes
    assert int(out["counts"].sum()) == 0 and not out["dets"].any()
    with pytest.raises(_lib.WssdlError):
        ops.detect_postprocess(np.zeros((2000, 5), np.float32), np.zeros((2000, 3), np.float32),
                               np.zeros((2000, 12), np.float32), meta)


def test_postprocess_matches_reference_generated_golden(monkeypatch):
    """The drop-in wrappers and the fused kernel against the reference's own code blocks
    (fast_rcnn/test_bus.py:207-223, :359-401 executed as fragments by
    tests/golden/make_layers_golden.py): regressed boxes within 1e-5, per-class NMS lists /
    class-agnostic pass / max_per_image cap bit for bit on the reference's boxes."""
    import os
    from wssdl_bus_b200.fast_rcnn import test_bus
    from wssdl_bus_b200.fast_rcnn.config import cfg
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_layers_golden.npz"))
    if str(g["numpy_version"]) != np.__version__:
        pytest.skip("fixture made with numpy %s" % g["numpy_version"])
    im_h, im_w, im_scale = g["det_meta"]
    pb = test_bus.detect_boxes(g["det_rois"], g["det_deltas"], (int(im_h), int(im_w), 3), float(im_scale))
    np.testing.assert_allclose(pb, g["det_pred_boxes"], rtol=RTOL, atol=1e-3)
    # discrete steps on the REFERENCE's boxes: exact
    got = test_bus.postprocess_detections(g["det_scores"], g["det_pred_boxes"])
    for j in (1, 2):
        assert np.array_equal(got[j], g["det_plain_cls%d" % j])
    monkeypatch.setattr(cfg.TEST, "CLS_AGNOSTIC_NMS", True)
    got = test_bus.postprocess_detections(g["det_scores"], g["det_pred_boxes"], max_per_image=40)
    for j in (1, 2):
        assert np.array_equal(got[j], g["det_agnostic_cap_cls%d" % j])
    monkeypatch.setattr(cfg.TEST, "CLS_AGNOSTIC_NMS", False)
    # the fused kernel end to end (its own regression): same lists within the box tolerance
    meta = np.array([[im_h, im_w, im_scale]], np.float32)
    out = ops.detect_postprocess(g["det_rois"], g["det_scores"], g["det_deltas"], meta,
                                 roi_stride=300, score_thresh=0.05, nms_thresh=0.3,
                                 max_per_image=300, cls_agnostic=False)
    dets, cnt = out["dets"].cpu().numpy(), out["counts"].cpu().numpy()
    for j in (1, 2):
        want = g["det_plain_cls%d" % j]
        assert cnt[0, j] == len(want)
        np.testing.assert_allclose(dets[0, j, :cnt[0, j]], want, rtol=RTOL, atol=1e-3)


def test_voc_sized_head_21_classes_fits(oracle_mod):
    """21 classes x 300 RoIs with max_per_image = 100 (a VOC-sized head at the reference's
    default settings): the per-class NMS needs mask rows for 300 boxes only; just the sort keys
    of the cap pass cover all (K-1)*300 detections."""
    B, S, K = 2, 300, 21
    rois = np.concatenate([syn.rois_for_pool(70 + b, S) for b in range(B)])
    scores, deltas = syn.rcnn_head_outputs(71, B * S, K)
    scores = np.ascontiguousarray(scores)
    scores[:, 1:] *= 6.0                               # enough rows above the 0.05 threshold
    meta = np.tile(np.array([[600, 800, 1.0]], np.float32), (B, 1))
    o = ops.detect_postprocess(rois, scores, deltas, meta, roi_stride=S, max_per_image=100,
                               want_pred_boxes=True)
    pb = o["pred_boxes"].cpu().numpy()
    dets, cnt = o["dets"].cpu().numpy(), o["counts"].cpu().numpy()
    for i in range(B):
        want = oracle_mod.layers.detections_postprocess(scores[i * S:(i + 1) * S], pb[i * S:(i + 1) * S],
                                                        thresh=np.float32(0.05), max_per_image=100)
        assert sum(len(w) for w in want[1:]) <= 100 + 20
        for j in range(1, K):
            assert cnt[i, j] == len(want[j])
            assert np.array_equal(dets[i, j, :cnt[i, j]], want[j])
