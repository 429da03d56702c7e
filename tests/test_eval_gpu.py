"""VOC AP / CorLoc / FROC evaluation on the detection blob (datasets/voc_eval_bus.py:143-275)
against the numpy restatement (oracle.layers.voc_eval_arrays)."""
import numpy as np
import pytest

from wssdl_bus_b200 import ops, synthetic as syn

pytestmark = pytest.mark.gpu


def _blob(seed, B, K, S, G=20):
    """Detections that overlap the GT often enough to exercise every branch: jittered copies
    of GT boxes (duplicates -> FP), random boxes (FP), difficult GT."""
    rng = np.random.default_rng(seed)
    gt = np.zeros((B, G, 5), np.float32)
    num = np.zeros((B,), np.int32)
    diff = np.zeros((B, G), np.uint8)
    dets = np.zeros((B, K, S, 5), np.float32)
    counts = np.zeros((B, K), np.int32)
    scores = rng.permutation(B * K * S).astype(np.float32).reshape(B, K, S)
    scores = (scores + 0.5) / scores.size
    for b in range(B):
        n = int(rng.integers(0, 6))
        num[b] = n
        gt[b, :n, :4] = syn.random_boxes(seed * 7 + b, n, lo=30, hi=250) if n else 0
        gt[b, :n, 4] = rng.integers(1, K, n)
        diff[b, :n] = rng.random(n) < 0.25
        for j in range(1, K):
            c = int(rng.integers(0, S + 1))
            counts[b, j] = c
            boxes = syn.random_boxes(seed * 13 + b * K + j, c, lo=20, hi=300) if c else np.zeros((0, 4), np.float32)
            own = np.where(gt[b, :n, 4] == j)[0]
            for i in range(c):
                if len(own) and rng.random() < 0.5:          # near-copy of a GT box of this class
                    g = gt[b, own[int(rng.integers(len(own)))], :4]
                    boxes[i] = g + rng.normal(0, 6, 4).astype(np.float32)
            order = np.argsort(-scores[b, j, :c])
            dets[b, j, :c, :4] = boxes
            dets[b, j, :c, 4] = scores[b, j, :c][order]       # class lists are in descending score
    return dets, counts, gt, num, diff


@pytest.mark.parametrize("B,K,S,use07", [(1, 2, 5, True), (16, 3, 40, True), (64, 3, 300, False),
                                         (7, 5, 33, True)])
def test_eval_matches_reference_restatement(oracle_mod, B, K, S, use07):
    from wssdl_bus_b200.datasets import voc_eval_bus
    dets, counts, gt, num, diff = _blob(1000 + B + S, B, K, S)
    res = voc_eval_bus.evaluate_detections_blob(dets, counts, gt, num, diff, ovthresh=0.5,
                                                use_07_metric=use07, score_thresh=0.5)
    assert len(res) == K - 1
    for r in res:
        j = r["cls"]
        ids, conf, BB = [], [], []
        for b in range(B):
            c = counts[b, j]
            ids += [b] * c
            conf += dets[b, j, :c, 4].tolist()
            BB += dets[b, j, :c, :4].tolist()
        gtb = [gt[b, :num[b], :4][gt[b, :num[b], 4] == j] for b in range(B)]
        gtd = [diff[b, :num[b]][gt[b, :num[b], 4] == j].astype(bool) for b in range(B)]
        rec, prec, ap, ni, nok, nfp, per_img = oracle_mod.layers.voc_eval_arrays(
            ids, conf, BB, gtb, gtd, ovthresh=0.5, use_07_metric=use07, score_thresh=0.5)
        assert (r["ni"], r["nok"]) == (ni, nok)
        if len(ids) == 0:
            assert r["ap"] == -1
            continue
        assert r["num_fp_per_img"] == per_img and r["num_all_fps"] == nfp
        npos = sum(int((~d).sum()) for d in gtd)
        if npos == 0:       # rec = tp / 0: nan/inf on both sides
            assert np.array_equal(r["prec"], prec)
            continue
        assert np.array_equal(r["rec"], rec) and np.array_equal(r["prec"], prec)
        assert r["ap"] == ap


def test_eval_hand_case_and_limits():
    """Duplicate detection -> FP, difficult match -> ignored, low score miss -> FP but not FROC."""
    from wssdl_bus_b200 import _lib
    dets = np.zeros((2, 2, 3, 5), np.float32)
    dets[0, 1, 0] = [10, 10, 50, 50, 0.9]
    dets[0, 1, 1] = [12, 12, 50, 50, 0.8]
    dets[1, 1, 0] = [0, 0, 20, 20, 0.7]
    dets[1, 1, 1] = [100, 100, 120, 120, 0.3]
    counts = np.array([[0, 2], [0, 2]], np.int32)
    gt = np.zeros((2, 4, 5), np.float32)
    gt[0, 0] = [10, 10, 50, 50, 1]
    gt[1, 0] = [0, 0, 20, 20, 1]
    diff = np.zeros((2, 4), np.uint8)
    diff[1, 0] = 1
    m = ops.eval_match(dets, counts, gt, np.array([1, 1], np.int32), diff)
    assert m["tp"].cpu().numpy()[:, 1, :2].tolist() == [[1, 0], [0, 0]]
    assert m["fp"].cpu().numpy()[:, 1, :2].tolist() == [[0, 1], [0, 1]]
    assert m["fp_froc"].cpu().numpy()[:, 1].sum() == 0
    assert m["npos"].cpu().numpy().tolist() == [0, 1]
    assert m["img_stats"].cpu().numpy()[:, 1].tolist() == [[1, 1], [1, 1]]
    with pytest.raises(_lib.WssdlError):
        ops.eval_match(dets, counts, np.zeros((2, 65, 5), np.float32), np.array([1, 1], np.int32))


def test_eval_matches_reference_generated_golden():
    """evaluate_detections_blob (matching on the device) against the output of the reference's
    OWN voc_eval_bus() on a synthetic VOC-style tree (tests/golden/make_layers_golden.py):
    rec / prec / AP (both flavours), CorLoc counts, FROC false positives, bit for bit."""
    import os
    from wssdl_bus_b200.datasets import voc_eval_bus
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_layers_golden.npz"))
    if str(g["numpy_version"]) != np.__version__:
        pytest.skip("fixture made with numpy %s" % g["numpy_version"])
    counts_gt = g["eval_gt_counts"]
    B = len(counts_gt)
    ofs = np.concatenate(([0], np.cumsum(counts_gt)))
    G = 4
    gt = np.zeros((B, G, 5), np.float32)
    num = counts_gt.astype(np.int32)
    diff = np.zeros((B, G), np.uint8)
    for b in range(B):
        n = num[b]
        gt[b, :n, :4] = g["eval_gt_bbox"][ofs[b]:ofs[b + 1]]
        gt[b, :n, 4] = 1
        diff[b, :n] = g["eval_gt_difficult"][ofs[b]:ofs[b + 1]]
    ids, conf, BB = g["eval_image_ids"], g["eval_confidence"], g["eval_BB"]
    S = int(np.bincount(ids, minlength=B).max())
    dets = np.zeros((B, 2, S, 5), np.float32)
    counts = np.zeros((B, 2), np.int32)
    for b in range(B):
        sel = np.where(ids == b)[0]
        sel = sel[np.argsort(-conf[sel])]                 # class lists are in descending score
        counts[b, 1] = len(sel)
        dets[b, 1, :len(sel), :4] = BB[sel]
        dets[b, 1, :len(sel), 4] = conf[sel]
    # the fixture's boxes and scores are exactly representable decimals x.y / 0.xyz parsed as
    # float64; the blob is float32: box coordinates k/10 below 1024 and scores k/1000 are not
    # all exact in float32, so compare what float32 can carry: the discrete outcomes and AP
    for tag, use07 in (("area", False), ("voc07", True)):
        r = voc_eval_bus.evaluate_detections_blob(dets, counts, gt, num, diff, ovthresh=0.5,
                                                  use_07_metric=use07, score_thresh=0.5)[0]
        ap, ni, nok, nfp = g["eval_%s_scalars" % tag].tolist()
        assert (r["ni"], r["nok"], r["num_all_fps"]) == (ni, nok, nfp)
        assert list(r["num_fp_per_img"]) == g["eval_%s_fp_per_img" % tag].tolist()
        assert np.array_equal(r["rec"], g["eval_%s_rec" % tag])
        assert np.array_equal(r["prec"], g["eval_%s_prec" % tag])
        assert r["ap"] == ap
