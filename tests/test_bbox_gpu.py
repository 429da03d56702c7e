"""IoU matrices (fp64 bit-exact), box transforms (1e-5 relative) and anchor labels
(bit-exact) on the GPU against the oracle."""
import numpy as np
import pytest
import torch

from wssdl_bus_b200 import ops, synthetic as syn
from wssdl_bus_b200.rpn_msr.generate_anchors import generate_anchors

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # north star: decoded boxes / IoU within 1e-5 relative


def test_iou_golden_bit_exact(golden):
    b, q = golden["iou_boxes"], golden["iou_query"]
    assert np.array_equal(ops.bbox_overlaps(b, q), golden["iou_out"])
    assert np.array_equal(ops.bbox_overlaps_ui(b, q), golden["iou_ui_out"])


@pytest.mark.parametrize("N,K", [(0, 5), (5, 0), (1, 1), (17100, 20), (5944, 3), (2020, 20),
                                 (10000, 1024), (3000, 1500), (777, 1023), (64, 65), (70, 4)])
def test_iou_vs_oracle_bit_exact(oracle_mod, N, K):
    b = syn.random_boxes(400 + N, N).astype(np.float64)
    q = syn.random_boxes(401 + K, K, lo=30, hi=300).astype(np.float64)
    got = ops.bbox_overlaps(b, q)
    assert got.dtype == np.float64 and got.shape == (N, K)
    assert np.array_equal(got, oracle_mod.clib.bbox_overlaps(b, q))
    assert np.array_equal(ops.bbox_overlaps_ui(b, q), oracle_mod.clib.bbox_overlaps(b, q, ui=True))


def test_iou_f32_fast_variant_and_properties(oracle_mod):
    b = syn.random_boxes(410, 4000)
    q = syn.random_boxes(411, 128)
    want = oracle_mod.clib.bbox_overlaps(b.astype(np.float64), q.astype(np.float64))
    got = ops.bbox_overlaps_device(b, q, ops.IOU, torch.float32).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-6)
    # C5-scale N x N: symmetry + unit diagonal, size-independent
    n = 20000
    bb = torch.from_numpy(syn.random_boxes(412, n).astype(np.float64)).cuda()
    m = ops.bbox_overlaps_device(bb, bb[:2048], ops.IOU, torch.float64)
    assert bool(torch.all(m[:2048].diagonal() == 1.0))
    assert bool(torch.equal(m[:2048], m[:2048].t()))
    assert float(m.min()) >= 0.0 and float(m.max()) <= 1.0


def test_bbox_transform_golden(golden):
    inv = ops.bbox_transform_inv(golden["bt_boxes"], golden["bt_deltas"])
    assert inv.dtype == np.float32 and inv.shape == golden["bt_inv"].shape
    np.testing.assert_allclose(inv, golden["bt_inv"], rtol=RTOL, atol=1e-4)
    clipped = ops.clip_boxes(golden["bt_inv"].copy(), (600, 800))
    assert np.array_equal(clipped, golden["bt_clip"])                      # exact: min/max only
    assert np.array_equal(ops.clip_boxes(clipped.copy(), (600, 800)), clipped)   # idempotent
    t = ops.bbox_transform(golden["bt_ex"], golden["bt_gt"])
    np.testing.assert_allclose(t, golden["bt_targets"], rtol=RTOL, atol=1e-6)
    assert ops.bbox_transform_inv(np.zeros((0, 4)), np.zeros((0, 12), np.float32)).shape == (0, 12)


@pytest.mark.parametrize("dataset,mode", [("SNUBH", 0), ("SNUBH_FG", 1), ("UDIAT", 2)])
def test_anchor_labels_bit_exact(oracle_mod, dataset, mode):
    B, H, W = 4, 38, 50
    gt, num = syn.gt_boxes(500, B)
    info = np.tile(np.array([[600, 800, 1.0]], np.float32), (B, 1))
    from wssdl_bus_b200.rpn_msr.generate_anchors import generate_anchors
    base = generate_anchors(scales=np.array([8, 16, 32]))
    labels, argmax, maxov = ops.anchor_labels(gt, num, info, H, W, base, 16, dataset_mode=mode)
    labels, argmax, maxov = labels.cpu().numpy(), argmax.cpu().numpy(), maxov.cpu().numpy()
    for b in range(B):
        lab = oracle_mod.layers.anchor_labels(H, W, gt[b, :num[b]], info[b], dataset=dataset)
        ins = lab["inds_inside"]
        assert len(ins) == 5944                                   # SURVEY probe for 600x800
        full = np.full(H * W * 9, -1, np.float32)
        full[ins] = lab["labels"]
        assert np.array_equal(labels[b], full)
        assert np.array_equal(argmax[b][ins], lab["argmax_overlaps"])
        assert np.all(np.delete(argmax[b], ins) == -1)
        assert np.array_equal(maxov[b][ins], lab["max_overlaps"])
        assert (lab["labels"] == 1).sum() > 0


def test_anchor_and_proposal_target_layers_match_oracle(oracle_mod):
    """Full layers with the host RNG seeded identically on both sides."""
    import numpy.random as npr
    from wssdl_bus_b200.rpn_msr import anchor_target_layer_tf_bus as atl
    from wssdl_bus_b200.rpn_msr import proposal_target_layer_tf_bus as ptl
    H, W, A = 38, 50, 9
    gt, num = syn.gt_boxes(510, 1)
    info = np.array([[600, 800, 1.0]], np.float32)
    score = np.zeros((1, H, W, 2 * A), np.float32)
    npr.seed(3)
    lab, tgt, iw, ow = atl.anchor_target_layer(score, gt, num, info, None, [16, ], [8, 16, 32], "SNUBH")
    npr.seed(3)
    o = oracle_mod.layers.anchor_labels(H, W, gt[0, :num[0]], info[0], dataset="SNUBH")
    ol, ot, oiw, oow = oracle_mod.layers.anchor_targets_from_labels(o, npr)
    want_lab = ol.reshape((1, H, W, A)).transpose(0, 3, 1, 2).reshape((1, 1, A * H, W))
    assert lab.shape == (1, 1, A * H, W) and np.array_equal(lab, want_lab)
    want_t = ot.reshape((1, H, W, A * 4)).transpose(0, 3, 1, 2)
    np.testing.assert_allclose(tgt, want_t, rtol=RTOL, atol=1e-6)
    assert np.array_equal(iw, oiw.reshape((1, H, W, A * 4)).transpose(0, 3, 1, 2))
    np.testing.assert_allclose(ow, oow.reshape((1, H, W, A * 4)).transpose(0, 3, 1, 2), rtol=1e-6)
    # joint flavour: 1 supervised + 2 weakly supervised images
    lab3, t3, _, _ = atl.anchor_target_layer_joint(np.zeros((3, H, W, 2 * A), np.float32),
                                                   np.tile(gt, (3, 1, 1)), np.tile(num, 3),
                                                   np.tile(info, (3, 1)), None, True, [16, ],
                                                   [8, 16, 32], "SNUBH")
    assert lab3.shape == (3, 1, A * H, W) and np.all(lab3[1:] == -1) and not t3[1:].any()
    # proposal target layer
    rois = syn.rois_for_pool(511, 300)
    npr.seed(4)
    r, l, t, iw2, ow2 = ptl.proposal_target_layer(rois, gt, num, 3, True, False)
    npr.seed(4)
    pos = gt[0, :int((gt[0, :num[0], 4] != 0).sum())]
    allr = np.vstack((rois, np.hstack((np.zeros((len(pos), 1), np.float32), pos[:, :4]))))
    ol, orois, ot, oiw, _ = oracle_mod.layers.sample_rois(allr, pos, 32, 128, 3, npr)
    assert np.array_equal(r, orois) and np.array_equal(l.ravel(), ol)
    np.testing.assert_allclose(t, ot, rtol=RTOL, atol=1e-6)
    assert np.array_equal(iw2, oiw) and np.array_equal(ow2, (oiw > 0).astype(np.float32))


def test_target_layers_match_reference_generated_golden(monkeypatch):
    """The drop-in layers (device IoU / labels / targets, host npr.choice) against fixtures
    produced by the reference's OWN anchor_target_layer[_joint] / proposal_target_layer /
    proposal_layer code (tests/golden/make_layers_golden.py), numpy.random seeded the same:
    labels, sampled RoIs and weights bit-exact, regression targets within 1e-5."""
    import os
    import numpy.random as npr
    from wssdl_bus_b200.fast_rcnn.config import cfg
    from wssdl_bus_b200.rpn_msr import anchor_target_layer_tf_bus as atl
    from wssdl_bus_b200.rpn_msr import proposal_target_layer_tf_bus as ptl
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_layers_golden.npz"))
    if str(g["numpy_version"]) != np.__version__:
        pytest.skip("fixture made with numpy %s" % g["numpy_version"])
    names = ("labels", "targets", "inside", "outside")

    def check(got, prefix):
        assert np.array_equal(got[0], g[prefix + "labels"])
        np.testing.assert_allclose(got[1], g[prefix + "targets"], rtol=RTOL, atol=1e-6)
        assert np.array_equal(got[2], g[prefix + "inside"])
        np.testing.assert_allclose(got[3], g[prefix + "outside"], rtol=1e-6)

    # joint layer: 2 supervised + 1 weakly supervised image on a 12x16 map
    monkeypatch.setattr(cfg.TRAIN, "IMS_PER_BATCH", 2)
    monkeypatch.setattr(cfg.TRAIN, "WS_IMS_PER_BATCH", 1)
    npr.seed(1234)
    got = atl.anchor_target_layer_joint(np.zeros((3, 12, 16, 18), np.float32), g["at_gt"],
                                        g["at_num"], g["prop_info"], None, True, [16, ],
                                        [8, 16, 32], "SNUBH")
    check(got, "at_joint_")
    # plain layer, 38x50 map: SNUBH, the general branch (bg draw happens), SNUBH with a
    # background box over most of the image (bg draw happens)
    score = np.zeros((1, 38, 50, 18), np.float32)
    for seed, prefix, gt, num, ds in ((4321, "at1_", g["at1_gt"], g["at1_num"], "SNUBH"),
                                      (2468, "at2_", g["at1_gt"], g["at1_num"], "VOC"),
                                      (1357, "at3_", g["at3_gt"], g["at3_num"], "SNUBH")):
        npr.seed(seed)
        check(atl.anchor_target_layer(score, gt, num, g["at1_info"], None, [16, ], [8, 16, 32], ds),
              prefix)
    # proposal target layer on the reference's own TRAIN proposals (573 RoIs of 2 images)
    monkeypatch.setattr(cfg.TRAIN, "WS_IMS_PER_BATCH", 0)
    npr.seed(777)
    r, l, t, iw, ow = ptl.proposal_target_layer(g["prop_blob_train"], g["at_gt"], g["at_num"], 3,
                                                True, False)
    assert np.array_equal(r, g["ptl_rois"]) and np.array_equal(l, g["ptl_labels"])
    np.testing.assert_allclose(t, g["ptl_targets"], rtol=RTOL, atol=1e-6)
    assert np.array_equal(iw, g["ptl_inside"]) and np.array_equal(ow, g["ptl_outside"])
    # proposal layer (drop-in signature): same RoIs as the reference within the decode
    # tolerance (np.exp vs correctly rounded exp), same count
    from wssdl_bus_b200.rpn_msr.proposal_layer_tf_bus import proposal_layer
    blob = proposal_layer(g["prop_cls"], g["prop_reg"], g["prop_info"], False, False)
    assert blob.shape == g["prop_blob_test"].shape
    np.testing.assert_allclose(blob, g["prop_blob_test"], rtol=RTOL, atol=1e-3)
    del names


def test_anchor_targets_device_sampler_philox(oracle_mod):
    """sampler="philox": the subsampling is drawn on the device (no host trip).  The draw differs
    from numpy's, so the check is structural: quotas met exactly, survivors are a subset of the
    pre-subsample labels, targets and inside weights independent of the sampler, outside weights
    1 / #examples, deterministic per seed, different across seeds and images; device tensors in ->
    device tensors out."""
    import torch
    from wssdl_bus_b200.fast_rcnn.config import cfg
    from wssdl_bus_b200.rpn_msr import anchor_target_layer_tf_bus as atl
    H, W, A, B = 38, 50, 9, 3
    gt, num = syn.gt_boxes(520, B)
    gt[1, 2] = [10, 10, 780, 590, 0]                     # a huge explicit background box: many bg anchors
    num[1] = max(num[1], 3)
    info = np.tile(np.array([[600, 800, 1.0]], np.float32), (B, 1))
    score = np.zeros((B, H, W, 2 * A), np.float32)
    np.random.seed(11)
    host = atl.anchor_target_layer(score, gt, num, info, None, [16, ], [8, 16, 32], "VOC")
    dev = atl.anchor_target_layer(score, torch.from_numpy(gt).cuda(), torch.from_numpy(num).cuda(),
                                  torch.from_numpy(info).cuda(), None, [16, ], [8, 16, 32], "VOC",
                                  sampler="philox", seed=1234)
    assert all(torch.is_tensor(t) and t.is_cuda for t in dev)
    lab, tgt, iw, ow = [t.cpu().numpy() for t in dev]
    again = atl.anchor_target_layer(score, gt, num, info, None, [16, ], [8, 16, 32], "VOC",
                                    sampler="philox", seed=1234)
    other = atl.anchor_target_layer(score, gt, num, info, None, [16, ], [8, 16, 32], "VOC",
                                    sampler="philox", seed=99)
    assert all(np.array_equal(a, b) for a, b in zip((lab, tgt, iw, ow), again))
    assert not np.array_equal(lab, other[0])
    assert np.array_equal(tgt, host[1])                   # targets do not depend on the draws
    num_fg = int(cfg.TRAIN.RPN_FG_FRACTION * cfg.TRAIN.RPN_BATCHSIZE)
    labels_pre, _, _ = ops.anchor_labels(gt, num, info, H, W, generate_anchors(), 16, dataset_mode=2,
                                         positive_overlap=cfg.TRAIN.RPN_POSITIVE_OVERLAP,
                                         negative_overlap=cfg.TRAIN.RPN_NEGATIVE_OVERLAP,
                                         clobber_positives=cfg.TRAIN.RPN_CLOBBER_POSITIVES)
    pre = labels_pre.cpu().numpy().reshape(B, H, W, A).transpose(0, 3, 1, 2).reshape(B, 1, A * H, W)
    drew = 0
    for b in range(B):
        n_fg, n_bg = int((pre[b] == 1).sum()), int((pre[b] == 0).sum())
        fg, bg = int((lab[b] == 1).sum()), int((lab[b] == 0).sum())
        assert fg == min(n_fg, num_fg) and bg == min(n_bg, cfg.TRAIN.RPN_BATCHSIZE - fg)
        assert np.all(pre[b][lab[b] == 1] == 1) and np.all(pre[b][lab[b] == 0] == 0)
        drew += (n_fg > fg) + (n_bg > bg)
        lab4 = np.repeat(lab[b].reshape(A, H, W), 4, axis=0)
        assert np.array_equal(iw[b] != 0, lab4 == 1)
        assert np.array_equal(ow[b][lab4 >= 0], np.full(int((lab4 >= 0).sum()), np.float32(1.0 / (fg + bg))))
        assert not ow[b][lab4 < 0].any()
    assert drew >= 2                                      # the sampler actually had to draw
    assert not np.array_equal(lab[0], lab[2]) or not np.array_equal(pre[0], pre[2])
    # the documented stream (include/wssdl_b200.h): the numpy restatement of the Philox selection
    # (oracle/philox.py, pinned by Random123's known-answer vectors) gives the same labels, exactly
    flat = labels_pre.cpu().numpy()
    for b in range(B):
        want = oracle_mod.philox.anchor_subsample(flat[b], b, 1234, num_fg, cfg.TRAIN.RPN_BATCHSIZE)
        want = want.reshape(H, W, A).transpose(2, 0, 1).reshape(1, A * H, W)
        assert np.array_equal(lab[b], want), b


def test_proposal_targets_device_sampler_philox(oracle_mod):
    """sampler="philox" for the proposal-target layer: fixed row stride, no host trip.  Structural
    checks against the oracle's candidate sets: the fg rows are distinct fg candidates, the bg rows
    distinct bg candidates, quotas as :243 / :256-258, labels / targets / weights of every row
    what the oracle computes for that row; deterministic per seed, different across seeds."""
    import torch
    from wssdl_bus_b200.fast_rcnn.config import cfg
    from wssdl_bus_b200.rpn_msr import proposal_target_layer_tf_bus as ptl
    B, K = 3, 3
    gt, num = syn.gt_boxes(530, B)
    rois = np.concatenate([np.hstack((np.full((300 + 50 * i, 1), i, np.float32),
                                      syn.rois_for_pool(531 + i, 300 + 50 * i)[:, 1:])) for i in range(B)])
    rois = rois[np.random.RandomState(5).permutation(len(rois))]        # interleaved images
    dev = torch.device("cuda:0")
    out = ptl.proposal_target_layer(torch.from_numpy(rois).to(dev), torch.from_numpy(gt).to(dev),
                                    torch.from_numpy(num.astype(np.int32)).to(dev), K, True, False,
                                    sampler="philox", seed=9)
    r, l, t, iw, ow, cnt = out
    assert all(x.is_cuda for x in out)
    S = cfg.TRAIN.BATCH_SIZE
    fgq = int(np.round(cfg.TRAIN.FG_FRACTION * S))
    assert r.shape == (B * S, 5) and l.shape == (B * S, 1) and t.shape == (B * S, 4 * K)
    r, l, t, iw, ow, cnt = (x.cpu().numpy() for x in out)
    for i in range(B):
        pos = gt[i, :int((gt[i, :num[i], 4] != 0).sum())]
        allr = np.vstack((rois[rois[:, 0] == i], np.hstack((np.full((len(pos), 1), i, np.float32), pos[:, :4]))))
        ov = oracle_mod.clib.bbox_overlaps(allr[:, 1:5].astype(np.float64), pos[:, :4].astype(np.float64))
        mx, am = ov.max(axis=1), ov.argmax(axis=1)
        fg = np.where(mx >= cfg.TRAIN.FG_THRESH)[0]
        bg = np.where((mx < cfg.TRAIN.BG_THRESH_HI) & (mx >= cfg.TRAIN.BG_THRESH_LO))[0]
        nf, nb = min(fgq, fg.size), min(S - min(fgq, fg.size), bg.size)
        assert tuple(cnt[i]) == (nf, nb)
        rows = r[i * S:i * S + nf + nb]
        key = {tuple(v): j for j, v in enumerate(map(tuple, allr))}
        idx = np.array([key[tuple(v)] for v in rows])
        assert len(set(idx[:nf])) == nf and set(idx[:nf]) <= set(fg)
        assert len(set(idx[nf:])) == nb and set(idx[nf:]) <= set(bg)
        want_l = pos[am[idx], 4].copy()
        want_l[nf:] = 0
        assert np.array_equal(l[i * S:i * S + nf + nb, 0], want_l)
        tg = oracle_mod.layers.bbox_transform(rows[:, 1:5], pos[am[idx], :4])
        if cfg.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED:
            tg = (tg - np.array(cfg.TRAIN.BBOX_NORMALIZE_MEANS)) / np.array(cfg.TRAIN.BBOX_NORMALIZE_STDS)
        for j in range(nf + nb):
            c = int(want_l[j])
            row_t, row_i = t[i * S + j].reshape(K, 4), iw[i * S + j].reshape(K, 4)
            if c > 0:
                np.testing.assert_allclose(row_t[c], tg[j], rtol=RTOL, atol=1e-6)
                assert np.array_equal(row_i[c], np.asarray(cfg.TRAIN.BBOX_INSIDE_WEIGHTS, np.float32))
            assert not row_t[[k for k in range(K) if k != c or c == 0]].any()
        assert np.array_equal(ow, (iw > 0).astype(np.float32))
        assert not r[i * S + nf + nb:(i + 1) * S].any() and not t[i * S + nf + nb:(i + 1) * S].any()
        # the documented stream: exactly the rows, in the order, the numpy restatement selects
        sel_fg, sel_bg = oracle_mod.philox.roi_select(fg.size, bg.size, i, 9, fgq, S)
        assert np.array_equal(idx, np.concatenate((fg[sel_fg], bg[sel_bg]))), i
    again = ptl.proposal_target_layer(rois, gt, num, K, True, False, sampler="philox", seed=9)
    assert np.array_equal(again[0], r) and np.array_equal(again[2], t)
    other = ptl.proposal_target_layer(rois, gt, num, K, True, False, sampler="philox", seed=10)
    assert not np.array_equal(other[0], r)
