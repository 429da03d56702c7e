"""Pin the CPU oracle (no GPU): C restatement vs the reference's own compiled Cython
(oracle/_ref), vs fixtures generated from the reference (tests/golden), vs the reference's
bbox_transform.py loaded by path, and vs the hand-derived RoI-pool bin tables of SURVEY.md
Appendix A.1 (the reference has no RoI-pool test vectors at all)."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import REFERENCE_PRESENT
from wssdl_bus_b200 import synthetic as syn


def test_anchor_table_golden(oracle_mod, golden):
    # the only golden vector the reference ships: generate_anchors.py:17-25 (1-based)
    a = oracle_mod.layers.generate_anchors()
    assert np.array_equal(a, golden["anchors_table_minus_1"])


@pytest.mark.parametrize("key", ["nms_uni_257", "nms_uni_1000", "nms_clu_257", "nms_clu_1000",
                                 "nms_fork"])
def test_c_nms_matches_reference_golden(oracle_mod, golden, key):
    d = golden[key + "_dets"]
    for t in (0.3, 0.5, 0.7):
        want = golden[key + "_keep_%02d" % int(t * 10)].tolist()
        assert oracle_mod.clib.nms(d, t) == want
    if key + "_keepnew_05" in golden:
        assert oracle_mod.clib.nms(d, 0.5, variant=1) == golden[key + "_keepnew_05"].tolist()


def test_threshold_fork_is_double_compare(oracle_mod, golden):
    # iou == 0.7f (=0.69999998) must NOT be suppressed at thresh 0.7; iou == 0.3f must be
    d = golden["nms_fork_dets"]
    assert oracle_mod.clib.nms(d, 0.7) == [0, 1, 2, 3]
    assert oracle_mod.clib.nms(d, 0.3) == [0, 2]


def test_c_iou_matches_reference_golden(oracle_mod, golden):
    b, q = golden["iou_boxes"], golden["iou_query"]
    assert np.array_equal(oracle_mod.clib.bbox_overlaps(b, q), golden["iou_out"])
    assert np.array_equal(oracle_mod.clib.bbox_overlaps(b, q, ui=True), golden["iou_ui_out"])


def test_layers_bbox_transform_matches_reference_golden(oracle_mod, golden):
    L = oracle_mod.layers
    inv = L.bbox_transform_inv(golden["bt_boxes"], golden["bt_deltas"])
    assert inv.dtype == np.float32
    # np.exp is the only non-exact step; same numpy here as when the fixture was made
    np.testing.assert_allclose(inv, golden["bt_inv"], rtol=1e-6, atol=1e-4)
    clipped = L.clip_boxes(golden["bt_inv"].copy(), np.array([600, 800], np.float32))
    assert np.array_equal(clipped, golden["bt_clip"])
    t = L.bbox_transform(golden["bt_ex"], golden["bt_gt"]).astype(np.float32)
    np.testing.assert_allclose(t, golden["bt_targets"], rtol=1e-6, atol=1e-6)


@pytest.mark.skipif(not REFERENCE_PRESENT, reason="/root/reference only exists in the authoring container")
def test_layers_vs_reference_module_by_path(oracle_mod):
    spec = importlib.util.spec_from_file_location(
        "ref_bt", "/root/reference/code/lib/fast_rcnn/bbox_transform.py")
    bt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bt)
    L = oracle_mod.layers
    rng = np.random.default_rng(11)
    boxes = syn.random_boxes(12, 500).astype(np.float64)
    deltas = (rng.standard_normal((500, 4)) * 0.5).astype(np.float32)
    assert np.array_equal(L.bbox_transform_inv(boxes, deltas), bt.bbox_transform_inv(boxes, deltas))
    a = L.bbox_transform_inv(boxes, deltas)
    assert np.array_equal(L.clip_boxes(a.copy(), (600, 800)), bt.clip_boxes(a.copy(), (600, 800)))
    assert np.array_equal(L.bbox_transform(boxes, boxes[::-1]), bt.bbox_transform(boxes, boxes[::-1]))


def test_ref_modules_vs_c_restatement(oracle_mod):
    ref, clib = oracle_mod.ref, oracle_mod.clib
    if not ref.available():
        pytest.skip("oracle/_ref not built (reference absent)")
    for seed, n, kw in ((21, 700, {}), (22, 900, {"clustered": True})):
        d = syn.dets(seed, n, **kw)
        for t in (0.3, 0.5, 0.7):
            assert ref.cpu_nms(d, t) == clib.nms(d, t)
        assert ref.nms_new(d, 0.7) == clib.nms(d, 0.7, variant=1)
    b = syn.random_boxes(23, 1500).astype(np.float64)
    q = syn.random_boxes(24, 33).astype(np.float64)
    assert np.array_equal(ref.bbox_overlaps(b, q), clib.bbox_overlaps(b, q))
    assert np.array_equal(ref.bbox_overlaps_ui(b, q), clib.bbox_overlaps(b, q, ui=True))


def test_nms_zero_union_raises(oracle_mod):
    d = np.array([[5, 5, 4, 4, 0.9], [7, 7, 6, 6, 0.8]], np.float32)   # areas 0, no overlap
    with pytest.raises(ZeroDivisionError):
        oracle_mod.clib.nms(d, 0.5)
    if oracle_mod.ref.available():
        with pytest.raises(ZeroDivisionError):
            oracle_mod.ref.cpu_nms(d, 0.5)


# ---- RoI pooling: hand-derived expectations
def _bins(clib, roi_h_cells, PH, mode):
    """Which rows each ph pools, observed through argmax on a map whose value is its row."""
    H, W, C = 80, 2, 1
    bottom = np.zeros((1, H, W, C), np.float32)
    bottom[0, :, :, 0] = np.arange(1, H + 1, dtype=np.float32)[:, None]   # increasing in h
    roi = np.array([[0, 0, 0, 0, (roi_h_cells - 1) * 16]], np.float32)
    top, arg = clib.roi_pool_fwd(bottom, roi, PH, 1, 1.0 / 16, bin_mode=mode)
    # max of an increasing map = last row of the bin; recover first row from a decreasing map
    bottom2 = bottom.copy()
    bottom2[0, :, :, 0] = np.arange(H, 0, -1, dtype=np.float32)[:, None]
    top2, arg2 = clib.roi_pool_fwd(bottom2, roi, PH, 1, 1.0 / 16, bin_mode=mode)
    out = []
    for ph in range(PH):
        if arg[0, ph, 0, 0] < 0:
            out.append(None)
        else:
            out.append((int(arg2[0, ph, 0, 0]) // (W * C), int(arg[0, ph, 0, 0]) // (W * C) + 1))
    return out


def test_roi_pool_bin_tables(oracle_mod):
    clib = oracle_mod.clib
    T, G = clib.CPU_TRUNC, clib.GPU_CEIL
    assert _bins(clib, 10, 7, T) == [(0, 1), (1, 2), (2, 4), (4, 5), (5, 7), (7, 8), (8, 10)]
    assert _bins(clib, 10, 7, G) == [(0, 2), (1, 3), (2, 5), (4, 6), (5, 8), (7, 9), (8, 10)]
    assert _bins(clib, 3, 7, T) == [None, None, (0, 1), None, (1, 2), None, (2, 3)]
    assert _bins(clib, 3, 7, G) == [(0, 1), (0, 1), (0, 2), (1, 2), (1, 3), (2, 3), (2, 3)]
    assert _bins(clib, 14, 7, T) == [(2 * i, 2 * i + 2) for i in range(7)]
    assert _bins(clib, 14, 7, G) == [(2 * i, 2 * i + 2) for i in range(7)]
    # fp32 edge products: 7*fl(31/7) = 30.999998 -> CPU last edge 30, GPU 31
    assert _bins(clib, 31, 7, T)[-1][1] == 30
    assert _bins(clib, 31, 7, G)[-1][1] == 31
    # 7*fl(57/7) = 57.000004 -> CPU 57, GPU 58 (one row past the RoI)
    assert _bins(clib, 57, 7, T)[-1][1] == 57
    assert _bins(clib, 57, 7, G)[-1][1] == 58


def test_roi_pool_rounding_and_ties(oracle_mod):
    clib = oracle_mod.clib
    # round() is half away from zero: x=8 -> 0.5 -> 1 ; x=24 -> 1.5 -> 2 (np.round would give 0, 2)
    bottom = np.zeros((1, 6, 6, 1), np.float32)
    bottom[0, :, :, 0] = np.arange(36, dtype=np.float32).reshape(6, 6)
    top, arg = clib.roi_pool_fwd(bottom, np.array([[0, 8, 8, 24, 24]], np.float32), 1, 1, 1 / 16.)
    assert top[0, 0, 0, 0] == 14.0 and arg[0, 0, 0, 0] == 14     # cells 1..2 x 1..2 -> (2,2)
    # ties: all-zero map -> first scanned cell wins, NOT -1
    z = np.zeros((1, 6, 6, 2), np.float32)
    top, arg = clib.roi_pool_fwd(z, np.array([[0, 16, 32, 64, 80]], np.float32), 1, 1, 1 / 16.)
    assert top[0, 0, 0].tolist() == [0.0, 0.0]
    assert arg[0, 0, 0].tolist() == [(2 * 6 + 1) * 2, (2 * 6 + 1) * 2 + 1]
    # values <= -FLT_MAX never win: (-FLT_MAX, -1)
    m = np.full((1, 2, 2, 1), -np.inf, np.float32)
    top, arg = clib.roi_pool_fwd(m, np.array([[0, 0, 0, 31, 31]], np.float32), 1, 1, 1 / 16.)
    assert top[0, 0, 0, 0] == -np.finfo(np.float32).max and arg[0, 0, 0, 0] == -1


def test_roi_pool_bwd_literal_equals_fast(oracle_mod):
    clib = oracle_mod.clib
    rng = np.random.default_rng(31)
    B, H, W, C, PH, PW = 2, 12, 14, 8, 7, 7
    bottom = syn.feature_map(32, B, H, W, C)
    rois = np.concatenate([syn.rois_for_pool(33, 40, B, im_w=W * 16, im_h=H * 16),
                           syn.adversarial_rois(B, W, H)])
    for mode in (clib.CPU_TRUNC, clib.GPU_CEIL):
        top, arg = clib.roi_pool_fwd(bottom, rois, PH, PW, 1 / 16., bin_mode=mode)
        g = rng.standard_normal(top.shape).astype(np.float32)
        a = clib.roi_pool_bwd(g, arg, rois, bottom.shape, 1 / 16., literal=True)
        b = clib.roi_pool_bwd(g, arg, rois, bottom.shape, 1 / 16., literal=False)
        assert np.array_equal(a, b)
        # scatter-through-argmax equals the gather except for malformed RoIs
        wellformed = (rois[:, 3] >= rois[:, 1]) & (rois[:, 4] >= rois[:, 2])
        ref = np.zeros((B, H * W * C), np.float64)
        for r in np.where(wellformed)[0]:
            bi = int(rois[r, 0])
            idx = arg[r].reshape(-1)
            ok = idx >= 0
            # GPU_CEIL can pool one row past the RoI (fl edge products); the gather's in-RoI
            # test drops those, so only compare CPU_TRUNC against the plain scatter
            np.add.at(ref[bi], idx[ok], g[r].reshape(-1)[ok].astype(np.float64))
        if mode == clib.CPU_TRUNC:
            np.testing.assert_allclose(a.reshape(B, -1), ref, rtol=1e-5, atol=1e-5)
    # malformed RoIs: forward pools one cell, backward contributes nothing
    bad = np.array([[0, 100, 40, 50, 120]], np.float32)   # x2 < x1, inside the 12x14 map
    top, arg = clib.roi_pool_fwd(bottom, bad, PH, PW, 1 / 16.)
    assert (arg >= 0).any()
    gz = clib.roi_pool_bwd(np.ones_like(top), arg, bad, bottom.shape, 1 / 16., literal=True)
    assert not gz.any()


def test_proposal_layer_restatement_properties(oracle_mod):
    L = oracle_mod.layers
    cls, reg, info = syn.rpn_outputs(41, 2, 10, 12, 9, im_h=160, im_w=192)
    cfg = dict(L.TEST, RPN_PRE_NMS_TOP_N=300, RPN_POST_NMS_TOP_N=50)
    blob, parts = L.proposal_layer(cls, reg, info, cfg=cfg, return_parts=True)
    assert blob.dtype == np.float32 and blob.shape[1] == 5
    for b, p in enumerate(parts):
        rows = blob[blob[:, 0] == b]
        assert len(rows) == len(p["anchor_idx"]) <= 50
        assert np.all(np.diff(p["scores"]) < 0)                      # descending, unique
        assert np.all(rows[:, 1] >= 0) and np.all(rows[:, 3] <= 191) and np.all(rows[:, 4] <= 159)
        assert np.all(rows[:, 3] - rows[:, 1] + 1 >= 16) and np.all(rows[:, 4] - rows[:, 2] + 1 >= 16)
        # clip is idempotent
        d = p["decoded"].copy()
        assert np.array_equal(L.clip_boxes(d.copy(), info[b, :2]), d)


def test_detections_postprocess_restatement_properties(oracle_mod):
    """Control flow of fast_rcnn/test_bus.py:360-401 (unpinned by the reference: no fixtures):
    survivors are a subset of the thresholded boxes in descending-score order, the cap keeps
    the top max_per_image scores, class-agnostic NMS never adds boxes."""
    R, K = 300, 3
    rois = syn.rois_for_pool(5, R)
    scores, deltas = syn.rcnn_head_outputs(6, R, K)
    assert len(np.unique(scores)) == scores.size
    pb = oracle_mod.layers.im_detect_boxes(rois, deltas, (437, 583, 3), 600.0 / 437)
    assert pb.shape == (R, 4 * K) and pb.dtype == np.float32
    assert (pb[:, 0::4] >= 0).all() and (pb[:, 2::4] <= 582).all() and (pb[:, 3::4] <= 436).all()
    base = oracle_mod.layers.detections_postprocess(scores, pb, max_per_image=0)
    assert base[0].shape == (0, 5)
    for j in range(1, K):
        d = base[j]
        assert (d[:, 4] > 0.05).all() and (np.diff(d[:, 4]) < 0).all()
        assert set(map(tuple, d[:, :4])) <= set(map(tuple, pb[:, 4 * j:4 * j + 4]))
    capped = oracle_mod.layers.detections_postprocess(scores, pb, max_per_image=20)
    assert sum(len(c) for c in capped) == 20
    top = np.sort(np.concatenate([b[:, 4] for b in base]))[-20:]
    assert np.array_equal(np.sort(np.concatenate([c[:, 4] for c in capped])), top)
    agn = oracle_mod.layers.detections_postprocess(scores, pb, max_per_image=0, cls_agnostic_nms=True)
    for j in range(1, K):
        assert set(map(tuple, agn[j])) <= set(map(tuple, base[j]))


def test_voc_eval_restatement_hand_cases(oracle_mod):
    """datasets/voc_eval_bus.py:143-275 restated (no fixtures in the reference): a duplicate
    detection is an FP, a match to a difficult box is ignored, AP07 of a perfect ranking is 1."""
    L = oracle_mod.layers
    gt = [np.array([[10, 10, 50, 50]]), np.array([[0, 0, 20, 20]])]
    dif = [np.array([False]), np.array([True])]
    rec, prec, ap, ni, nok, nfp, per_img = L.voc_eval_arrays(
        [0, 0, 1, 1], [0.9, 0.8, 0.7, 0.3],
        [[10, 10, 50, 50], [12, 12, 50, 50], [0, 0, 20, 20], [100, 100, 120, 120]], gt, dif,
        use_07_metric=True)
    assert rec.tolist() == [1, 1, 1, 1] and np.allclose(prec, [1, .5, .5, 1 / 3.])
    assert abs(ap - 1.0) < 1e-12 and (ni, nok, nfp, per_img) == (2, 2, 0, [0, 0])
    # a confident miss is a FROC false positive of its image; CorLoc drops
    rec, prec, ap, ni, nok, nfp, per_img = L.voc_eval_arrays(
        [0], [0.95], [[200, 200, 260, 260]], gt, dif, use_07_metric=False)
    assert (ni, nok, nfp, per_img) == (2, 0, 1, [1, 0]) and ap == 0
    assert L.voc_eval_arrays([], [], np.zeros((0, 4)), gt, dif)[2] == -1
    assert abs(L.voc_ap(np.array([.5, 1.]), np.array([1., .5]), False) - 0.75) < 1e-12


def _roi_pool_fwd_python(bottom, rois, PH, PW, scale, gpu_bins):
    """Second, independent restatement of the RoiPool body in plain Python loops with numpy
    float32 scalars (roi_pooling_op.cc:141-195; gpu_bins: roi_pooling_op_gpu.cu.cc:51-58).
    Only for tiny cases: it pins the C restatement the GPU parity tests use."""
    import math
    f32 = np.float32
    B, H, W, C = bottom.shape
    R = rois.shape[0]
    top = np.zeros((R, PH, PW, C), np.float32)
    arg = np.zeros((R, PH, PW, C), np.int32)

    def c_round(x):                                   # C round(): half away from zero
        x = float(x)
        return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))

    for n in range(R):
        r = rois[n]
        b = int(r[0])
        sw, sh = c_round(f32(r[1]) * f32(scale)), c_round(f32(r[2]) * f32(scale))
        ew, eh = c_round(f32(r[3]) * f32(scale)), c_round(f32(r[4]) * f32(scale))
        roi_w, roi_h = max(ew - sw + 1, 1), max(eh - sh + 1, 1)
        bin_h, bin_w = f32(roi_h) / f32(PH), f32(roi_w) / f32(PW)
        for ph in range(PH):
            for pw in range(PW):
                if gpu_bins:
                    hs, he = math.floor(f32(ph) * bin_h), math.ceil(f32(ph + 1) * bin_h)
                    ws, we = math.floor(f32(pw) * bin_w), math.ceil(f32(pw + 1) * bin_w)
                else:                                 # static_cast<int> BEFORE floor / ceil
                    hs, he = int(f32(ph) * bin_h), int(f32(ph + 1) * bin_h)
                    ws, we = int(f32(pw) * bin_w), int(f32(pw + 1) * bin_w)
                hs, he = min(max(hs + sh, 0), H), min(max(he + sh, 0), H)
                ws, we = min(max(ws + sw, 0), W), min(max(we + sw, 0), W)
                empty = he <= hs or we <= ws
                for c in range(C):
                    maxval, maxidx = (f32(0) if empty else f32(-3.4028234663852886e38)), -1
                    for h in range(hs, he):
                        for w in range(ws, we):
                            v = bottom[b, h, w, c]
                            if v > maxval:
                                maxval, maxidx = v, (h * W + w) * C + c
                    top[n, ph, pw, c], arg[n, ph, pw, c] = maxval, maxidx
    return top, arg


@pytest.mark.parametrize("gpu_bins", [False, True])
def test_roi_pool_fwd_c_restatement_equals_python_restatement(oracle_mod, gpu_bins):
    """Two restatements of roi_pooling_op.cc written independently (C with float/int casts,
    Python with numpy float32 scalars) must agree bit for bit on random and adversarial RoIs:
    an independent check next to the reference's own compiled kernel (see
    test_roi_pool_restatement_equals_reference_kernel), and besides the hand-derived bin tables
    the only one for the GPU_CEIL bins."""
    from wssdl_bus_b200 import synthetic as syn
    B, H, W, C = 2, 13, 17, 3
    bottom = syn.feature_map(70, B, H, W, C)
    bottom[0, 2:4, 3:6] = -np.inf
    bottom[1, 0, 0, :] = np.nan
    rois = np.concatenate([
        syn.rois_for_pool(71, 40, B, im_w=W * 16, im_h=H * 16),
        syn.adversarial_rois(B, W, H)[:, :],
        np.array([[0, -37.0, -90.0, 500.0, 700.0], [1, 8.0, 24.0, 8.0, 24.0],
                  [1, 100.0, 60.0, 20.0, 10.0]], np.float32)])
    for PH, PW in ((7, 7), (3, 5), (1, 1), (14, 14)):
        want_t, want_a = _roi_pool_fwd_python(bottom, rois, PH, PW, 1 / 16., gpu_bins)
        got_t, got_a = oracle_mod.clib.roi_pool_fwd(bottom, rois, PH, PW, 1 / 16.,
                                                    bin_mode=1 if gpu_bins else 0)
        assert np.array_equal(got_a, want_a), (PH, PW)
        assert np.array_equal(got_t, want_t, equal_nan=True), (PH, PW)


def _roi_pool_bwd_python(top_diff, argmax, rois, shape, scale):
    """Independent restatement of the RoiPoolGrad body (roi_pooling_op.cc:387-457) in plain
    Python with numpy float32 scalars: a gather over input cells, float32 accumulation in
    (roi, ph, pw) order."""
    import math
    f32 = np.float32
    B, H, W, C = shape
    R, PH, PW, _ = top_diff.shape
    out = np.zeros(shape, np.float32)

    def c_round(x):
        x = float(x)
        return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))

    geo = []
    for r in rois:
        sw, sh = c_round(f32(r[1]) * f32(scale)), c_round(f32(r[2]) * f32(scale))
        ew, eh = c_round(f32(r[3]) * f32(scale)), c_round(f32(r[4]) * f32(scale))
        bin_h = f32(max(eh - sh + 1, 1)) / f32(PH)
        bin_w = f32(max(ew - sw + 1, 1)) / f32(PW)
        geo.append((int(r[0]), sw, sh, ew, eh, bin_h, bin_w))
    for n in range(B):
        for h in range(H):
            for w in range(W):
                for c in range(C):
                    grad = f32(0)
                    for roi_n, (b, sw, sh, ew, eh, bin_h, bin_w) in enumerate(geo):
                        if n != b or not (sw <= w <= ew and sh <= h <= eh):
                            continue
                        phs = math.floor(f32(h - sh) / bin_h)
                        phe = math.ceil(f32(h - sh + 1) / bin_h)
                        pws = math.floor(f32(w - sw) / bin_w)
                        pwe = math.ceil(f32(w - sw + 1) / bin_w)
                        phs, phe = min(max(phs, 0), PH), min(max(phe, 0), PH)
                        pws, pwe = min(max(pws, 0), PW), min(max(pwe, 0), PW)
                        for ph in range(phs, phe):
                            for pw in range(pws, pwe):
                                if argmax[roi_n, ph, pw, c] == (h * W + w) * C + c:
                                    grad = f32(grad + top_diff[roi_n, ph, pw, c])
                    out[n, h, w, c] = grad
    return out


def test_roi_pool_bwd_c_restatement_equals_python_restatement(oracle_mod):
    """Same idea as the forward cross-check, for RoiPoolGrad: the C restatement (both its
    literal gather and its fast form) against an independent Python restatement, bit for bit
    (the accumulation order is the reference's, so float32 sums agree exactly)."""
    clib = oracle_mod.clib
    rng = np.random.default_rng(81)
    B, H, W, C, PH, PW = 2, 7, 9, 2, 3, 3
    bottom = syn.feature_map(82, B, H, W, C)
    rois = np.concatenate([
        syn.rois_for_pool(83, 14, B, im_w=W * 16, im_h=H * 16),
        np.array([[0, -20.0, -30.0, 200.0, 150.0], [1, 16.0, 16.0, 16.0, 16.0],
                  [1, 100.0, 60.0, 20.0, 10.0], [0, 0.0, 0.0, 47.0, 47.0]], np.float32)])
    for mode in (clib.CPU_TRUNC, clib.GPU_CEIL):
        top, arg = clib.roi_pool_fwd(bottom, rois, PH, PW, 1 / 16., bin_mode=mode)
        g = rng.standard_normal(top.shape).astype(np.float32)
        want = _roi_pool_bwd_python(g, arg, rois, bottom.shape, 1 / 16.)
        for literal in (True, False):
            got = clib.roi_pool_bwd(g, arg, rois, bottom.shape, 1 / 16., literal=literal)
            assert np.array_equal(got, want), (mode, literal)
    # arbitrary argmax (never produced by a forward): the gather's feasibility tests decide
    arg = rng.integers(-1, H * W * C, size=(rois.shape[0], PH, PW, C)).astype(np.int32)
    g = rng.standard_normal(arg.shape).astype(np.float32)
    want = _roi_pool_bwd_python(g, arg, rois, bottom.shape, 1 / 16.)
    assert np.array_equal(clib.roi_pool_bwd(g, arg, rois, bottom.shape, 1 / 16., literal=True), want)


# ---- glue layers against fixtures produced by the reference's own layer code
# (tests/golden/make_layers_golden.py runs rpn_msr/*_tf_bus.py from /root/reference under a
# mechanical py2->py3 shim; inputs are seeded, numpy.random is seeded before every call)
@pytest.fixture(scope="module")
def layers_golden():
    path = os.path.join(os.path.dirname(__file__), "golden", "reference_layers_golden.npz")
    g = np.load(path)
    if str(g["numpy_version"]) != np.__version__:
        pytest.skip("fixture made with numpy %s (np.exp / RandomState streams are only "
                    "guaranteed identical on the same numpy)" % g["numpy_version"])
    return g


def test_proposal_layer_restatement_equals_reference_output(oracle_mod, layers_golden):
    g, L = layers_golden, oracle_mod.layers
    cls, reg, info = g["prop_cls"], g["prop_reg"], g["prop_info"]
    assert np.array_equal(L.proposal_layer(cls, reg, info, is_training=False), g["prop_blob_test"])
    assert np.array_equal(L.proposal_layer(cls, reg, info, is_training=True), g["prop_blob_train"])
    c1 = syn.rpn_outputs(int(g["prop_c1_seed"]), 1, 38, 50, 9)
    blob = L.proposal_layer(c1[0], c1[1], c1[2], is_training=False)
    assert blob.shape == (300, 5) and np.array_equal(blob, g["prop_c1_blob_test"])


def _anchor_layer_from_blocks(L, H, W, gts, infos, seed, dataset="SNUBH"):
    """anchor_target_layer[_joint] output layout (anchor_target_layer_tf_bus.py:568-611)
    assembled from the oracle's per-image blocks, with numpy.random as the injected RNG."""
    A = 9
    np.random.seed(seed)
    outs = [[], [], [], []]
    for gt, info in zip(gts, infos):
        lab = L.anchor_labels(H, W, gt, info, dataset=dataset)
        labels, targets, inside, outside = L.anchor_targets_from_labels(lab, np.random)
        outs[0].append(labels.reshape((1, H, W, A)).transpose(0, 3, 1, 2).reshape((1, 1, A * H, W)))
        for k, v in ((1, targets), (2, inside), (3, outside)):
            outs[k].append(v.reshape((1, H, W, A * 4)).transpose(0, 3, 1, 2))
    return [np.concatenate(o) for o in outs]


def test_anchor_target_layer_restatement_equals_reference_output(oracle_mod, layers_golden):
    g, L = layers_golden, oracle_mod.layers
    # joint layer: two supervised images + one weakly supervised image (labels -1, zeros)
    gt, num, info = g["at_gt"], g["at_num"], g["prop_info"]
    got = _anchor_layer_from_blocks(L, 12, 16, [gt[i, :num[i]] for i in range(2)], info, 1234)
    for k, name in enumerate(("labels", "targets", "inside", "outside")):
        want = g["at_joint_" + name]
        assert np.array_equal(got[k], want[:2]), name
        assert np.all(want[2] == (-1 if name == "labels" else 0))
    assert (got[0] == 1).sum() > 0 and (got[0] == 0).sum() > 0
    # plain layer at the full 38x50 map: fg/bg subsampling through npr.choice
    gt1, num1 = g["at1_gt"], g["at1_num"]
    got = _anchor_layer_from_blocks(L, 38, 50, [gt1[0, :num1[0]]], g["at1_info"], 4321)
    for k, name in enumerate(("labels", "targets", "inside", "outside")):
        assert np.array_equal(got[k], g["at1_" + name]), name
    # cases where the npr.choice draws happen (256 = RPN_BATCHSIZE labelled anchors remain):
    # the general (non-SNUBH) branch, and SNUBH with a background box over most of the image
    got = _anchor_layer_from_blocks(L, 38, 50, [gt1[0, :num1[0]]], g["at1_info"], 2468, dataset="VOC")
    for k, name in enumerate(("labels", "targets", "inside", "outside")):
        assert np.array_equal(got[k], g["at2_" + name]), name
    assert (got[0] == 0).sum() + (got[0] == 1).sum() == 256 and (got[0] == 0).sum() > 200
    gt3, num3 = g["at3_gt"], g["at3_num"]
    got = _anchor_layer_from_blocks(L, 38, 50, [gt3[0, :num3[0]]], g["at1_info"], 1357)
    for k, name in enumerate(("labels", "targets", "inside", "outside")):
        assert np.array_equal(got[k], g["at3_" + name]), name
    assert (got[0] == 0).sum() + (got[0] == 1).sum() == 256 and (got[0] == 0).sum() > 200


def test_proposal_target_layer_restatement_equals_reference_output(oracle_mod, layers_golden):
    g, L = layers_golden, oracle_mod.layers
    rois_in, gt, num = g["prop_blob_train"], g["at_gt"], g["at_num"]
    np.random.seed(777)
    got = [[], [], [], []]
    for i in range(2):
        all_rois = rois_in[rois_in[:, 0] == i]
        t_gt = gt[i, :num[i]]
        fg_gt = t_gt[:int(np.sum(t_gt[:, 4] != 0))]
        # GT boxes join the candidates (proposal_target_layer_tf_bus.py:45-50)
        extra = np.hstack((np.full((fg_gt.shape[0], 1), i, fg_gt.dtype), fg_gt[:, :-1]))
        all_rois = np.vstack((all_rois, extra))
        labels, rois, targets, inside, _ = L.sample_rois(all_rois, fg_gt, 32, 128, 3, np.random)
        for k, v in enumerate((rois, labels.reshape(-1, 1), targets, inside)):
            got[k].append(v)
    got = [np.concatenate(v) for v in got]
    for k, name in enumerate(("rois", "labels", "targets", "inside")):
        assert np.array_equal(got[k], g["ptl_" + name]), name
    assert np.array_equal((got[3] > 0).astype(np.float32), g["ptl_outside"])


def test_voc_eval_restatement_equals_reference_output(oracle_mod, layers_golden):
    """voc_eval_arrays against the reference's voc_eval_bus() run on a synthetic VOC-style tree
    (XML annotations, a detection text file) in make_layers_golden.py: rec / prec arrays, AP in
    both flavours, CorLoc counts and FROC false-positive counts, all bit for bit."""
    g, L = layers_golden, oracle_mod.layers
    counts = g["eval_gt_counts"]
    ofs = np.concatenate(([0], np.cumsum(counts)))
    gt_bbox = [g["eval_gt_bbox"][ofs[i]:ofs[i + 1]] for i in range(len(counts))]
    gt_diff = [g["eval_gt_difficult"][ofs[i]:ofs[i + 1]] for i in range(len(counts))]
    assert g["eval_gt_difficult"].any() and (counts == 0).any()
    for tag, use07 in (("area", False), ("voc07", True)):
        rec, prec, ap, ni, nok, nfp, fp_per_img = L.voc_eval_arrays(
            g["eval_image_ids"], g["eval_confidence"], g["eval_BB"], gt_bbox, gt_diff,
            ovthresh=0.5, use_07_metric=use07, score_thresh=0.5)
        assert np.array_equal(rec, g["eval_%s_rec" % tag])
        assert np.array_equal(prec, g["eval_%s_prec" % tag])
        assert [ap, ni, nok, nfp] == g["eval_%s_scalars" % tag].tolist()
        assert list(fp_per_img) == g["eval_%s_fp_per_img" % tag].tolist()


def test_detection_postprocess_restatement_equals_reference_output(oracle_mod, layers_golden):
    """im_detect_boxes / detections_postprocess against the reference's own code blocks
    (fast_rcnn/test_bus.py:207-223 and :359-401, executed as fragments by
    make_layers_golden.py): regressed + clipped boxes, per-class NMS 0.3 lists, the
    class-agnostic pass and the max_per_image cap, bit for bit."""
    g, L = layers_golden, oracle_mod.layers
    im_h, im_w, im_scale = g["det_meta"]
    pred = L.im_detect_boxes(g["det_rois"], g["det_deltas"], (int(im_h), int(im_w), 3), float(im_scale))
    assert pred.dtype == g["det_pred_boxes"].dtype and np.array_equal(pred, g["det_pred_boxes"])
    for tag, agnostic, cap in (("plain", False, 300), ("agnostic_cap", True, 40)):
        out = L.detections_postprocess(g["det_scores"], pred, thresh=0.05, max_per_image=cap,
                                       cls_agnostic_nms=agnostic)
        for j in (1, 2):
            want = g["det_%s_cls%d" % (tag, j)]
            assert out[j].shape == want.shape and np.array_equal(out[j], want), (tag, j)
    assert sum(len(g["det_agnostic_cap_cls%d" % j]) for j in (1, 2)) == 40
    assert len(g["det_plain_cls1"]) > 40


# ---- the reference's own RoiPool / RoiPoolGrad CPU kernels (roi_pooling_op.cc compiled
# unmodified against oracle/tf_stub into oracle/_ref/ref_roi_pool.so)
def _need_ref_roi_pool(oracle_mod):
    if not oracle_mod.ref.roi_pool_available():
        pytest.skip("oracle/_ref/ref_roi_pool.so not built (reference absent)")


def test_roi_pool_restatement_equals_reference_kernel(oracle_mod):
    """The C restatement the GPU parity tests check against (oracle/hotpath_ref.c, CPU_TRUNC
    bins) versus the object code of the reference's own RoiPoolOp / RoiPoolGradOp: bit for
    bit on proposal-like, random and adversarial RoIs, several pooled sizes and thread counts,
    NaN / -inf cells, malformed RoIs."""
    _need_ref_roi_pool(oracle_mod)
    clib, ref = oracle_mod.clib, oracle_mod.ref
    rng = np.random.default_rng(91)
    for trial, (B, H, W, C) in enumerate([(2, 38, 50, 16), (3, 13, 17, 5), (1, 7, 9, 3), (4, 20, 24, 8)]):
        bottom = syn.feature_map(92 + trial, B, H, W, C)
        bottom[0, H // 3:H // 2, W // 4:W // 2] = -np.inf
        bottom[B - 1, 0, 0, :] = np.nan
        rois = np.concatenate([syn.rois_for_pool(93 + trial, 60, B, im_w=W * 16, im_h=H * 16),
                               syn.adversarial_rois(B, W, H)])
        for PH, PW in ((7, 7), (14, 14), (3, 5), (1, 1)):
            for threads in (1, 5):
                top, arg = ref.roi_pool_fwd(bottom, rois, PH, PW, 1 / 16., threads=threads)
                wt, wa = clib.roi_pool_fwd(bottom, rois, PH, PW, 1 / 16., bin_mode=clib.CPU_TRUNC)
                assert np.array_equal(arg, wa), (trial, PH, PW)
                assert np.array_equal(top, wt, equal_nan=True), (trial, PH, PW)
            g = rng.standard_normal(top.shape).astype(np.float32)
            want = ref.roi_pool_bwd(g, arg, rois, bottom.shape, 1 / 16., threads=3)
            for literal in (True, False):
                got = clib.roi_pool_bwd(g, arg, rois, bottom.shape, 1 / 16., literal=literal)
                assert np.array_equal(got, want), (trial, PH, PW, literal)
        # arbitrary argmax tensors (never produced by a forward)
        arg = rng.integers(-1, H * W * C, size=(rois.shape[0], 3, 3, C)).astype(np.int32)
        g = rng.standard_normal(arg.shape).astype(np.float32)
        assert np.array_equal(clib.roi_pool_bwd(g, arg, rois, bottom.shape, 1 / 16., literal=True),
                              ref.roi_pool_bwd(g, arg, rois, bottom.shape, 1 / 16.))


def test_reference_kernel_attribute_checks(oracle_mod):
    """The op's own checks surface through the driver (roi_pooling_op.cc:73-82)."""
    _need_ref_roi_pool(oracle_mod)
    bottom = np.zeros((1, 4, 4, 2), np.float32)
    with pytest.raises(RuntimeError, match="pooled_height"):
        oracle_mod.ref.roi_pool_fwd(bottom, np.zeros((1, 5), np.float32), -1, 7, 1 / 16.)


def test_gpu_ceil_restatement_equals_reference_cuda_kernel_source(oracle_mod):
    """bin_mode GPU_CEIL of the C restatement versus the reference's CUDA kernel
    (roi_pooling_op_gpu.cu.cc ROIPoolForward), whose body is compiled for the host and run once
    per emulated CUDA thread (oracle/build_ref.py:build_cuda_twin, oracle/tf_stub/cuda_emu.h):
    bit for bit, and different from the CPU op where the two bin rules differ."""
    if not oracle_mod.ref.cuda_twin_available():
        pytest.skip("oracle/_ref/ref_roi_pool_cudatwin.so not built (reference absent)")
    clib, ref = oracle_mod.clib, oracle_mod.ref
    differs = 0
    for trial, (B, H, W, C) in enumerate([(2, 38, 50, 8), (3, 13, 17, 5), (1, 60, 61, 2)]):
        bottom = syn.feature_map(96 + trial, B, H, W, C)
        bottom[0, 1:3, 2:5] = -np.inf
        bottom[B - 1, 0, 0, :] = np.nan
        rois = np.concatenate([syn.rois_for_pool(97 + trial, 80, B, im_w=W * 16, im_h=H * 16),
                               syn.adversarial_rois(B, W, H)])
        for PH, PW in ((7, 7), (14, 14), (3, 5), (1, 1)):
            top, arg = ref.roi_pool_fwd_cuda_twin(bottom, rois, PH, PW, 1 / 16.)
            wt, wa = clib.roi_pool_fwd(bottom, rois, PH, PW, 1 / 16., bin_mode=clib.GPU_CEIL)
            assert np.array_equal(arg, wa), (trial, PH, PW)
            assert np.array_equal(top, wt, equal_nan=True), (trial, PH, PW)
            _, ca = clib.roi_pool_fwd(bottom, rois, PH, PW, 1 / 16., bin_mode=clib.CPU_TRUNC)
            differs += int(not np.array_equal(ca, wa))
    assert differs >= 6          # the fork of SURVEY.md section 0.1 is real


def test_reference_cuda_nms_host_build_agrees_with_its_python_twin(oracle_mod, golden):
    """nms/nms_kernel.cu (the '>' NMS behind gpu_nms: devIoU, the 64-bit bitmask kernel and the
    host sweep _nms) built for the host (oracle/build_ref.py:build_cuda_nms) versus the
    reference's pure-Python twin nms/py_cpu_nms.py loaded by path: same keep lists, and the
    documented fork against cpu_nms ('>=' on a double threshold) at IoU == 0.3f / 0.7f."""
    if not oracle_mod.ref.gpu_nms_available():
        pytest.skip("oracle/_ref/ref_gpu_nms_hostbuild.so not built (reference absent)")
    ref = oracle_mod.ref
    fork = golden["nms_fork_dets"]
    assert [int(k) for k in ref.gpu_nms(fork, 0.7)] == [0, 1, 2, 3]
    assert [int(k) for k in ref.gpu_nms(fork, 0.3)] == [0, 2, 3]          # cpu_nms: [0, 2]
    assert golden["nms_fork_keep_03"].tolist() == [0, 2]
    if not REFERENCE_PRESENT:
        return
    spec = importlib.util.spec_from_file_location(
        "ref_py_cpu_nms", "/root/reference/code/lib/nms/py_cpu_nms.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    for seed, n, kw in ((300, 2500, {"clustered": True}), (301, 1000, {}), (302, 65, {}), (303, 64, {}),
                        (304, 1, {}), (305, 129, {"clustered": True})):
        d = syn.dets(seed, n, **kw)
        for t in (0.3, 0.5, 0.7):
            assert [int(k) for k in ref.gpu_nms(d, t)] == [int(k) for k in m.py_cpu_nms(d, t)]


def test_philox_restatement_against_random123_known_answers():
    """oracle/philox.py: Philox4x32-10 against the known-answer vectors Random123 publishes
    (kat_vectors: zeros, all ones, digits of pi), and the selection helpers' basic properties."""
    from oracle import philox as ph
    kat = (((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)))
    for ctr, key, want in kat:
        got = tuple(int(np.asarray(v).reshape(-1)[0]) for v in ph.philox4x32_10(*ctr, *key))
        assert got == want
    # vectorised counters give the scalar results
    v = ph.philox4x32_10(np.arange(5), 7, 1, 0, 123, 456)
    for i in range(5):
        assert tuple(int(a[i]) for a in v) == tuple(
            int(np.asarray(a).reshape(-1)[0]) for a in ph.philox4x32_10(i, 7, 1, 0, 123, 456))
    labels = np.array([1] * 10 + [0] * 20 + [-1] * 5, np.float32)
    out = ph.anchor_subsample(labels, 0, 42, num_fg=4, batchsize=12)
    assert (out == 1).sum() == 4 and (out == 0).sum() == 8 and np.all(out[labels == -1] == -1)
    assert np.all(labels[out == 1] == 1) and np.all(labels[out == 0] == 0)
    f, b = ph.roi_select(10, 50, 1, 42, 4, 16)
    assert len(f) == 4 and len(b) == 12 and len(set(f)) == 4 and len(set(b)) == 12
    assert np.array_equal(ph.roi_select(10, 50, 1, 42, 4, 16)[1], b)
    assert not np.array_equal(ph.roi_select(10, 50, 2, 42, 4, 16)[1], b)
