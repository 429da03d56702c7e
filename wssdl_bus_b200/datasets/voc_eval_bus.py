"""datasets/voc_eval_bus.py twin for the gathered detection blob.

voc_ap(rec, prec, use_07_metric)         -- :37-66, unchanged arithmetic (numpy, O(n))
voc_eval_bus_blob(dets, counts, ...)     -- voc_eval_bus (:68-281) from the point where the
    reference has parsed its result file and annotation XMLs: the per-detection IoU matching,
    TP/FP/FROC marking and the CorLoc pass run on the device (csrc/eval.cu,
    wssdl_eval_match); the global argsort by confidence, the cumulative sums and the AP use the
    reference's own numpy calls on the returned flags.
evaluate_detections_blob(...)            -- the per-class loop of bus._do_python_eval
    (datasets/bus.py:263-392): AP, CorLoc, FROC counts per class.

Deviation (documented): the reference round-trips detections through a text file
('{:.3f}' scores, '{:.1f}' coordinates, datasets/bus.py:245-262); here the float32 values of
the blob are used as they are.
"""
import numpy as np

from wssdl_bus_b200 import ops


def voc_ap(rec, prec, use_07_metric=False):
    """ap = voc_ap(rec, prec, [use_07_metric]) -- voc_eval_bus.py:37-66."""
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            if np.sum(rec >= t) == 0:
                p = 0
            else:
                p = np.max(prec[rec >= t])
            ap = ap + p / 11.
    else:
        mrec = np.concatenate(([0.], rec, [1.]))
        mpre = np.concatenate(([0.], prec, [0.]))
        for i in range(mpre.size - 1, 0, -1):
            mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
        i = np.where(mrec[1:] != mrec[:-1])[0]
        ap = np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])
    return ap


def voc_eval_bus_blob(dets, counts, gt_boxes, num_gt, cls, difficult=None, ovthresh=0.5,
                      use_07_metric=False, score_thresh=0.5, match=None):
    """One class of voc_eval_bus on the blob.  dets [B,K,S,5] / counts [B,K] (device or numpy),
    gt_boxes [B,G,5] (x1,y1,x2,y2,cls), num_gt [B].  Returns the reference's tuple without
    arr_ok: (rec, prec, ap, ni, nok, num_all_fps, num_fp_per_img).  `match` may carry the
    result of ops.eval_match so several classes share one launch."""
    if match is None:
        match = ops.eval_match(dets, counts, gt_boxes, num_gt, difficult, ovthresh, score_thresh)
    cnt = (counts.cpu().numpy() if hasattr(counts, "cpu") else np.asarray(counts))[:, cls]
    d = dets.cpu().numpy() if hasattr(dets, "cpu") else np.asarray(dets)
    B, S = cnt.shape[0], d.shape[2]
    valid = np.arange(S)[None, :] < cnt[:, None]                     # file order: image, then rank
    image_ids = np.broadcast_to(np.arange(B)[:, None], (B, S))[valid]
    confidence = d[:, cls, :, 4][valid].astype(float)
    stats = match["img_stats"].cpu().numpy()[:, cls]
    ni, nok = int(stats[:, 0].sum()), int(stats[:, 1].sum())
    if confidence.size == 0:                                         # :276-279
        return -1, -1, -1, ni, nok, 0, [0] * B
    tp = match["tp"].cpu().numpy()[:, cls][valid].astype(float)
    fp = match["fp"].cpu().numpy()[:, cls][valid].astype(float)
    fp_froc = match["fp_froc"].cpu().numpy()[:, cls][valid].astype(float)
    npos = int(match["npos"].cpu().numpy()[cls])
    sorted_ind = np.argsort(-confidence)                             # :155
    tp, fp, fp_froc, image_ids = tp[sorted_ind], fp[sorted_ind], fp_froc[sorted_ind], image_ids[sorted_ind]
    num_all_fps = np.sum(fp_froc)                                    # :254
    num_fp_per_img = [int(np.sum(fp_froc[image_ids == i])) for i in range(B)]
    fp = np.cumsum(fp)                                               # :265-271
    tp = np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    ap = voc_ap(rec, prec, use_07_metric)
    return rec, prec, ap, ni, nok, num_all_fps, num_fp_per_img


def evaluate_detections_blob(dets, counts, gt_boxes, num_gt, difficult=None, ovthresh=0.5,
                             use_07_metric=True, score_thresh=0.5):
    """Per-class loop of bus._do_python_eval (datasets/bus.py:263-392): one device launch for
    all classes, then AP / CorLoc / FROC bookkeeping per class.  Returns a list of dicts for
    classes 1..K-1."""
    match = ops.eval_match(dets, counts, gt_boxes, num_gt, difficult, ovthresh, score_thresh)
    K = match["npos"].shape[0]
    out = []
    for cls in range(1, K):
        rec, prec, ap, ni, nok, nfp, per_img = voc_eval_bus_blob(
            dets, counts, gt_boxes, num_gt, cls, difficult, ovthresh, use_07_metric, score_thresh,
            match=match)
        out.append(dict(cls=cls, rec=rec, prec=prec, ap=ap, ni=ni, nok=nok,
                        corloc=(nok / float(ni)) if ni else float("nan"),
                        num_all_fps=nfp, num_fp_per_img=per_img))
    return out
