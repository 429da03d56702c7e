"""Fused proposal layer on the GPU against the oracle.

Decoded boxes: <= 1e-5 relative (np.exp vs correctly-rounded exp).  Everything discrete
after the decode -- min-size filter, top-N order, NMS keep, final blob -- is compared
BIT-EXACTLY by feeding the device-decoded boxes into the oracle's remaining steps
(SURVEY.md section 7.8)."""
import numpy as np
import pytest

from wssdl_bus_b200 import ops, synthetic as syn
from wssdl_bus_b200.rpn_msr.generate_anchors import generate_anchors

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(autouse=True, params=["cta", "cluster2", "cluster4", "cluster8"])
def proposals_variant(request, tuning):
    """Every test runs four times: one CTA per image, and thread-block clusters of 2 / 4 / 8 CTAs
    per image (csrc/proposal.cu: both stages of the keep-list NMS rounds are split over the
    cluster, alive bits and column masks are exchanged through distributed shared memory).  The
    default picks by batch size."""
    tuning("proposals_cluster", 0 if request.param == "cta" else int(request.param[7:]))
    return request.param


def _run(oracle_mod, seed, B, H, W, pre, post, thresh=0.7, im_h=600, im_w=800, scale=1.0,
         sigma=None):
    cls, reg, info = syn.rpn_outputs(seed, B, H, W, 9, im_h=im_h, im_w=im_w, im_scale=scale)
    if sigma is not None:
        reg *= sigma
    base = generate_anchors()
    out = ops.proposals(cls, reg, info, base, 16, pre, post, thresh, 16, want_decoded=True)
    dec = out["decoded"].cpu().numpy()
    cfg = dict(oracle_mod.layers.TEST, RPN_PRE_NMS_TOP_N=pre, RPN_POST_NMS_TOP_N=post,
               RPN_NMS_THRESH=thresh)
    # 1) decode tolerance against the pure-numpy path
    _, parts_np = oracle_mod.layers.proposal_layer(cls, reg, info, cfg=cfg, return_parts=True)
    for b in range(B):
        np.testing.assert_allclose(dec[b], parts_np[b]["decoded"], rtol=RTOL, atol=1e-3)
    # 2) discrete steps bit-exact on identical boxes
    blob, parts = oracle_mod.layers.proposal_layer(cls, reg, info, cfg=cfg, return_parts=True,
                                                   decoded_override=dec)
    counts = out["counts"].cpu().numpy()
    rois = out["rois"].cpu().numpy().reshape(B, post, 5)
    scores = out["scores"].cpu().numpy().reshape(B, post)
    aidx = out["anchor_idx"].cpu().numpy().reshape(B, post)
    for b in range(B):
        n = len(parts[b]["anchor_idx"])
        assert counts[b] == n
        assert np.array_equal(aidx[b, :n], parts[b]["anchor_idx"])
        assert np.array_equal(scores[b, :n], parts[b]["scores"])
        assert np.array_equal(rois[b, :n], blob[blob[:, 0] == b])
        assert not rois[b, n:].any() and np.all(aidx[b, n:] == -1)
    got_blob = ops.compact_rois(out).cpu().numpy()
    assert np.array_equal(got_blob, blob)
    return counts


def test_c1_test_config_bit_exact(oracle_mod):
    counts = _run(oracle_mod, 600, 1, 38, 50, 6000, 300)
    assert counts[0] == 300


def test_c2_train_shapes_2000(oracle_mod):
    _run(oracle_mod, 601, 1, 38, 50, 2000, 2000)


def test_reference_train_defaults_12000_2000(oracle_mod):
    _run(oracle_mod, 602, 1, 38, 50, 12000, 2000)


@pytest.mark.parametrize("thresh", [0.3, 0.5])
def test_batch_and_thresholds(oracle_mod, thresh):
    _run(oracle_mod, 603, 5, 38, 50, 6000, 300, thresh=thresh)


def test_heavy_suppression_small_deltas(oracle_mod):
    # tiny deltas: proposals stay close to their anchors -> clusters of 9 per cell, many
    # candidates visited per kept box, several NMS chunks
    _run(oracle_mod, 604, 2, 38, 50, 6000, 300, sigma=0.05)


def test_small_maps_and_filtering(oracle_mod):
    # few anchors (< pre_nms_topN), scaled image: min_size*scale = 40 filters many boxes
    _run(oracle_mod, 605, 3, 10, 12, 6000, 300, im_h=160, im_w=192, scale=2.5)
    _run(oracle_mod, 606, 2, 5, 4, 50, 7, im_h=80, im_w=64)
    _run(oracle_mod, 607, 1, 38, 50, 300, 300)
    _run(oracle_mod, 608, 1, 19, 25, 1, 1, im_h=300, im_w=400)


def test_score_ties_follow_documented_order(oracle_mod):
    cls, reg, info = syn.rpn_outputs(609, 1, 12, 12, 9, im_h=192, im_w=192)
    cls[..., 9:] = np.round(cls[..., 9:] * 16) / 16          # massive ties
    base = generate_anchors()
    out = ops.proposals(cls, reg, info, base, 16, 200, 50, 0.7, 16, want_decoded=True)
    dec = out["decoded"].cpu().numpy()[0]
    sc = cls[0, :, :, 9:].reshape(-1)
    keep = oracle_mod.layers.filter_boxes(dec, np.float32(16))
    order = keep[sc[keep].argsort(kind="stable")[::-1][:200]]   # (score desc, index desc)
    d = np.hstack([dec[order], sc[order][:, None]])
    k = oracle_mod.clib.nms(d, 0.7, order=np.arange(len(d)))[:50]
    n = int(out["counts"][0])
    assert n == len(k)
    assert np.array_equal(out["anchor_idx"].cpu().numpy()[:n], order[k])


def test_dropin_proposal_layer_numpy(oracle_mod):
    from wssdl_bus_b200.rpn_msr.proposal_layer_tf_bus import proposal_layer
    cls, reg, info = syn.rpn_outputs(610, 2, 38, 50, 9, info_cols=4)
    blob = proposal_layer(cls, reg, info, False, False, [16, ], [8, 16, 32])
    assert isinstance(blob, np.ndarray) and blob.dtype == np.float32 and blob.shape == (600, 5)
    want = oracle_mod.layers.proposal_layer(cls, reg, info)
    assert blob.shape == want.shape
    # same RoIs up to the exp tolerance whenever the discrete decisions agree (they do here)
    np.testing.assert_allclose(blob, want, rtol=RTOL, atol=1e-3)


def test_no_truncation_and_deep_batches(oracle_mod):
    """The candidates are selected + sorted lazily, 1024 (or 2*post) at a time.  Cases that need
    more than one batch: heavy suppression at a low threshold (hundreds of candidates per kept
    box), pre_nms_topN <= 0 = no truncation (proposal_layer_tf_bus.py:130: all ~17100 valid
    anchors are candidates), a keep list that never fills (every candidate is visited), and
    massive score ties across the batch boundaries."""
    _run(oracle_mod, 620, 2, 38, 50, 6000, 300, thresh=0.3, sigma=0.02)
    _run(oracle_mod, 621, 1, 38, 50, 0, 300)
    _run(oracle_mod, 622, 1, 38, 50, -1, 300, thresh=0.3, sigma=0.02)
    _run(oracle_mod, 623, 1, 38, 50, 5000, 4096, thresh=0.9)
    _run(oracle_mod, 624, 2, 20, 24, 0, 64, thresh=0.1, sigma=0.01)


def test_batch_boundaries_with_score_ties(oracle_mod):
    """Ties at the thresholds between two batches: (score desc, index desc) must hold across
    the boundary, and nothing may be visited twice or skipped."""
    cls, reg, info = syn.rpn_outputs(625, 1, 38, 50, 9)
    cls[..., 9:] = np.round(cls[..., 9:] * 8) / 8            # 9 distinct scores, ~1900 anchors each
    reg *= 0.02
    base = generate_anchors()
    out = ops.proposals(cls, reg, info, base, 16, 6000, 300, 0.3, 16, want_decoded=True)
    dec = out["decoded"].cpu().numpy()[0]
    sc = cls[0, :, :, 9:].reshape(-1)
    keep = oracle_mod.layers.filter_boxes(dec, np.float32(16))
    order = keep[sc[keep].argsort(kind="stable")[::-1][:6000]]   # (score desc, index desc)
    d = np.hstack([dec[order], sc[order][:, None]])
    k = oracle_mod.clib.nms(d, 0.3, order=np.arange(len(d)))[:300]
    n = int(out["counts"][0])
    assert n == len(k)
    assert np.array_equal(out["anchor_idx"].cpu().numpy()[:n], order[k])


def test_post_nms_topn_zero_means_no_truncation(oracle_mod):
    """proposal_layer_tf_bus.py:139: post_nms_topN <= 0 keeps every survivor; the device blob
    then has stride min(pre_nms_topN, anchors)."""
    cls, reg, info = syn.rpn_outputs(626, 2, 12, 16, 9, im_h=192, im_w=256)
    out = ops.proposals(cls, reg, info, generate_anchors(), 16, 700, 0, 0.7, 16, want_decoded=True)
    assert out["post_nms_topN"] == 700 and out["rois"].shape == (2 * 700, 5)
    cfg = dict(oracle_mod.layers.TEST, RPN_PRE_NMS_TOP_N=700, RPN_POST_NMS_TOP_N=0)
    blob, _ = oracle_mod.layers.proposal_layer(cls, reg, info, cfg=cfg, return_parts=True,
                                               decoded_override=out["decoded"].cpu().numpy())
    assert np.array_equal(ops.compact_rois(out).cpu().numpy(), blob)


def test_gpu_nms_mode_follows_use_gpu_nms(oracle_mod, monkeypatch):
    """cfg.USE_GPU_NMS routes nms_wrapper.nms to gpu_nms ('>' against the float threshold,
    nms_kernel.cu:71) instead of cpu_nms ('>=' against the double one): the fused layer follows.
    Checked on boxes built so that pairs sit exactly AT the threshold."""
    from wssdl_bus_b200.fast_rcnn.config import cfg
    from wssdl_bus_b200.rpn_msr import proposal_layer_tf_bus as pl
    H, W, A = 6, 6, 9
    cls, reg, info = syn.rpn_outputs(627, 1, H, W, A, im_h=H * 16, im_w=W * 16)
    base = generate_anchors()
    out_ge = ops.proposals(cls, reg, info, base, 16, 300, 300, 0.5, 2, want_decoded=True)
    out_gt = ops.proposals(cls, reg, info, base, 16, 300, 300, 0.5, 2, nms_mode=ops.NMS_GT_F32)
    dec = out_ge["decoded"].cpu().numpy()[0]
    sc = cls[0, :, :, A:].reshape(-1)
    keep = oracle_mod.layers.filter_boxes(dec, np.float32(2))
    order = keep[sc[keep].argsort(kind="stable")[::-1][:300]]
    d = np.hstack([dec[order], sc[order][:, None]]).astype(np.float32)

    def greedy(dets, keep_if):                       # py_cpu_nms.py:10-38 with a pluggable rule
        x1, y1, x2, y2 = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3]
        areas = (x2 - x1 + 1) * (y2 - y1 + 1)
        alive = np.ones(len(dets), bool)
        kept = []
        for i in range(len(dets)):
            if not alive[i]:
                continue
            kept.append(i)
            w = np.maximum(np.float32(0), np.minimum(x2[i], x2) - np.maximum(x1[i], x1) + 1)
            h = np.maximum(np.float32(0), np.minimum(y2[i], y2) - np.maximum(y1[i], y1) + 1)
            inter = w * h
            ovr = inter / (areas[i] + areas - inter)
            alive &= keep_if(ovr)
            alive[i] = False
        return kept
    k_gt = greedy(d, lambda o: o <= np.float32(0.5))                       # gpu_nms: drop iff > 0.5f
    k_ge = greedy(d, lambda o: ~(o.astype(np.float64) >= 0.5))             # cpu_nms: drop iff >= 0.5
    n_gt, n_ge = int(out_gt["counts"][0]), int(out_ge["counts"][0])
    assert np.array_equal(out_gt["anchor_idx"].cpu().numpy()[:n_gt], order[k_gt])
    assert np.array_equal(out_ge["anchor_idx"].cpu().numpy()[:n_ge], order[k_ge])
    # the drop-in layer follows the config switch
    monkeypatch.setattr(cfg, "USE_GPU_NMS", True)
    assert pl._nms_mode() == ops.NMS_GT_F32
    blob = pl.proposal_layer(cls, reg, info, False, False, [16, ], [8, 16, 32])
    monkeypatch.setattr(cfg.TEST, "RPN_NMS_THRESH", 0.5)
    monkeypatch.setattr(cfg.TEST, "RPN_MIN_SIZE", 2)
    blob = pl.proposal_layer(cls, reg, info, False, False, [16, ], [8, 16, 32])
    assert np.array_equal(blob[:, 1:], dec[order[k_gt]][:300])


def test_limits_are_reported():
    cls, reg, info = syn.rpn_outputs(611, 1, 38, 50, 9)
    with pytest.raises(ops.WssdlError, match="limits"):
        ops.proposals(cls, reg, info, generate_anchors(), 16, 6000, 5000, 0.7, 16)   # keep list > 4096
    with pytest.raises(ops.WssdlError, match="limits"):
        ops.proposals(cls, reg, info, generate_anchors(), 16, 0, 0, 0.7, 16)          # no truncation at all: 17100
