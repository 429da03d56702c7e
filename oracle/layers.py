"""numpy restatements of the reference's python glue on the hot path.

TEST INFRASTRUCTURE ONLY.  The reference modules are python-2 (print statements,
xrange, implicit relative imports, `cfg` from easydict) and cannot be imported under
python 3; `fast_rcnn/bbox_transform.py` can be loaded by file path in the authoring
container but does not exist on the GPU box.  Each function below restates one
reference function and cites it; tests/test_oracle.py checks the restatements against
the real module (when /root/reference is present) and against tests/golden fixtures
generated from it.  Glue parity (proposal_layer / anchor_target / proposal_target
composition) is UNPINNED by the reference, which ships no fixtures for it; it is
composed here from the pinned pieces (oracle.ref.*) so only control flow is restated.
"""
import numpy as np

from . import clib, ref

# hot-path constants: fast_rcnn/config.py
TEST = dict(RPN_NMS_THRESH=0.7, RPN_PRE_NMS_TOP_N=6000, RPN_POST_NMS_TOP_N=300, RPN_MIN_SIZE=16,
            NMS=0.3)
TRAIN = dict(RPN_NMS_THRESH=0.7, RPN_PRE_NMS_TOP_N=12000, RPN_POST_NMS_TOP_N=2000,
             RPN_MIN_SIZE=16, RPN_POSITIVE_OVERLAP=0.7, RPN_NEGATIVE_OVERLAP=0.3,
             RPN_CLOBBER_POSITIVES=False, RPN_FG_FRACTION=0.5, RPN_BATCHSIZE=256,
             BATCH_SIZE=128, FG_FRACTION=0.25, FG_THRESH=0.5, BG_THRESH_HI=0.5,
             BG_THRESH_LO=0.0, RPN_BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0),
             RPN_POSITIVE_WEIGHT=-1.0, BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0))


def _nms(dets, thresh):
    """cpu_nms: the reference binary when built, else its C restatement."""
    if ref.available():
        return ref.cpu_nms(dets, thresh)
    return clib.nms(dets, thresh)


def _overlaps(boxes, query, ui=False):
    if ref.available():
        return (ref.bbox_overlaps_ui if ui else ref.bbox_overlaps)(boxes, query)
    return clib.bbox_overlaps(boxes, query, ui=ui)


# ---------------------------------------------------------------- generate_anchors.py:37-97
def _whctrs(anchor):
    w = anchor[2] - anchor[0] + 1
    h = anchor[3] - anchor[1] + 1
    return w, h, anchor[0] + 0.5 * (w - 1), anchor[1] + 0.5 * (h - 1)


def _mkanchors(ws, hs, x_ctr, y_ctr):
    ws = ws[:, np.newaxis]
    hs = hs[:, np.newaxis]
    return np.hstack((x_ctr - 0.5 * (ws - 1), y_ctr - 0.5 * (hs - 1),
                      x_ctr + 0.5 * (ws - 1), y_ctr + 0.5 * (hs - 1)))


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=2 ** np.arange(3, 6)):
    """generate_anchors.py:37-48 (ratio enum :75-86, scale enum :88-97)."""
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales)
    base_anchor = np.array([1, 1, base_size, base_size]) - 1
    w, h, x_ctr, y_ctr = _whctrs(base_anchor)
    size_ratios = (w * h) / ratios
    ws = np.round(np.sqrt(size_ratios))
    hs = np.round(ws * ratios)
    ratio_anchors = _mkanchors(ws, hs, x_ctr, y_ctr)
    out = []
    for i in range(ratio_anchors.shape[0]):
        w, h, x_ctr, y_ctr = _whctrs(ratio_anchors[i, :])
        out.append(_mkanchors(w * scales, h * scales, x_ctr, y_ctr))
    return np.vstack(out)


def shifted_anchors(height, width, feat_stride=16, anchor_scales=(8, 16, 32)):
    """proposal_layer_tf_bus.py:49-71 == anchor_target_layer_tf_bus.py:369-382.
    Returns (K*A, 4) float64, row order (h, w, a)."""
    _anchors = generate_anchors(scales=np.array(anchor_scales))
    A = _anchors.shape[0]
    shift_x = np.arange(0, width) * feat_stride
    shift_y = np.arange(0, height) * feat_stride
    shift_x, shift_y = np.meshgrid(shift_x, shift_y)
    shifts = np.vstack((shift_x.ravel(), shift_y.ravel(),
                        shift_x.ravel(), shift_y.ravel())).transpose()
    K = shifts.shape[0]
    anchors = _anchors.reshape((1, A, 4)) + shifts.reshape((1, K, 4)).transpose((1, 0, 2))
    return anchors.reshape((K * A, 4)), A


# ---------------------------------------------------------------- fast_rcnn/bbox_transform.py
def bbox_transform(ex_rois, gt_rois):
    """bbox_transform.py:10-28."""
    ex_widths = ex_rois[:, 2] - ex_rois[:, 0] + 1.0
    ex_heights = ex_rois[:, 3] - ex_rois[:, 1] + 1.0
    ex_ctr_x = ex_rois[:, 0] + 0.5 * ex_widths
    ex_ctr_y = ex_rois[:, 1] + 0.5 * ex_heights
    gt_widths = gt_rois[:, 2] - gt_rois[:, 0] + 1.0
    gt_heights = gt_rois[:, 3] - gt_rois[:, 1] + 1.0
    gt_ctr_x = gt_rois[:, 0] + 0.5 * gt_widths
    gt_ctr_y = gt_rois[:, 1] + 0.5 * gt_heights
    targets_dx = (gt_ctr_x - ex_ctr_x) / ex_widths
    targets_dy = (gt_ctr_y - ex_ctr_y) / ex_heights
    targets_dw = np.log(gt_widths / ex_widths)
    targets_dh = np.log(gt_heights / ex_heights)
    return np.vstack((targets_dx, targets_dy, targets_dw, targets_dh)).transpose()


def bbox_transform_inv(boxes, deltas):
    """bbox_transform.py:30-61 (boxes are cast to the deltas' dtype, :34)."""
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), dtype=deltas.dtype)
    boxes = boxes.astype(deltas.dtype, copy=False)
    widths = boxes[:, 2] - boxes[:, 0] + 1.0
    heights = boxes[:, 3] - boxes[:, 1] + 1.0
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    dx, dy, dw, dh = deltas[:, 0::4], deltas[:, 1::4], deltas[:, 2::4], deltas[:, 3::4]
    pred_ctr_x = dx * widths[:, np.newaxis] + ctr_x[:, np.newaxis]
    pred_ctr_y = dy * heights[:, np.newaxis] + ctr_y[:, np.newaxis]
    pred_w = np.exp(dw) * widths[:, np.newaxis]
    pred_h = np.exp(dh) * heights[:, np.newaxis]
    pred_boxes = np.zeros(deltas.shape, dtype=deltas.dtype)
    pred_boxes[:, 0::4] = pred_ctr_x - 0.5 * pred_w
    pred_boxes[:, 1::4] = pred_ctr_y - 0.5 * pred_h
    pred_boxes[:, 2::4] = pred_ctr_x + 0.5 * pred_w
    pred_boxes[:, 3::4] = pred_ctr_y + 0.5 * pred_h
    return pred_boxes


def clip_boxes(boxes, im_shape):
    """bbox_transform.py:63-77 (in place)."""
    boxes[:, 0::4] = np.maximum(np.minimum(boxes[:, 0::4], im_shape[1] - 1), 0)
    boxes[:, 1::4] = np.maximum(np.minimum(boxes[:, 1::4], im_shape[0] - 1), 0)
    boxes[:, 2::4] = np.maximum(np.minimum(boxes[:, 2::4], im_shape[1] - 1), 0)
    boxes[:, 3::4] = np.maximum(np.minimum(boxes[:, 3::4], im_shape[0] - 1), 0)
    return boxes


def filter_boxes(boxes, min_size):
    """proposal_layer_tf_bus.py:151-156."""
    ws = boxes[:, 2] - boxes[:, 0] + 1
    hs = boxes[:, 3] - boxes[:, 1] + 1
    return np.where((ws >= min_size) & (hs >= min_size))[0]


# ---------------------------------------------------------------- rpn_msr/proposal_layer_tf_bus.py
def proposal_layer(rpn_cls_prob_reshape, rpn_bbox_pred, im_info, is_training=False,
                   feat_stride=16, anchor_scales=(8, 16, 32), cfg=None, return_parts=False,
                   decoded_override=None):
    """proposal_layer_tf_bus.py:19-148.  Inputs NHWC: [B,H,W,2A], [B,H,W,4A], [B,>=3].
    Returns the (sum R, 5) float32 blob; with return_parts also per-image dicts holding
    the decoded+clipped proposals (row order (h,w,a)), the pre-NMS order and scores.
    `decoded_override[i]` (optional, (K*A,4) f32) replaces step 1-2 for image i so that
    a caller can feed device-decoded boxes and compare the discrete steps bit-exactly."""
    cfg = cfg or (TRAIN if is_training else TEST)
    pre_nms_topN = cfg["RPN_PRE_NMS_TOP_N"]
    post_nms_topN = cfg["RPN_POST_NMS_TOP_N"]
    nms_thresh = cfg["RPN_NMS_THRESH"]
    min_size = cfg["RPN_MIN_SIZE"]
    # :34-35 NHWC -> NCHW
    cls = np.transpose(rpn_cls_prob_reshape, [0, 3, 1, 2])
    reg = np.transpose(rpn_bbox_pred, [0, 3, 1, 2])
    batch_size = im_info.shape[0]
    height, width = cls.shape[-2:]
    anchors, A = shifted_anchors(height, width, feat_stride, anchor_scales)
    blob = np.zeros((0, 5), dtype=np.float32)
    parts = []
    for i in range(batch_size):
        t_im_info = im_info[i, :]
        scores = cls[[i], A:, :, :]                                   # :86
        bbox_deltas = reg[[i], :, :, :]
        bbox_deltas = bbox_deltas.transpose((0, 2, 3, 1)).reshape((-1, 4))   # :106
        scores = scores.transpose((0, 2, 3, 1)).reshape((-1, 1))             # :113
        if decoded_override is not None:
            proposals = np.array(decoded_override[i], dtype=np.float32, copy=True)
        else:
            proposals = bbox_transform_inv(anchors, bbox_deltas)             # :116
            proposals = clip_boxes(proposals, t_im_info[:2])                 # :119
        all_props = proposals
        keep = filter_boxes(proposals, min_size * t_im_info[2])              # :123
        proposals = proposals[keep, :]
        scores = scores[keep]
        order = scores.ravel().argsort()[::-1]                               # :129
        if pre_nms_topN > 0:
            order = order[:pre_nms_topN]
        proposals = proposals[order, :]
        scores = scores[order]
        pre_idx = keep[order]
        if proposals.shape[0] == 0:                                          # nms_wrapper.py:16
            keep_nms = []
        else:
            keep_nms = _nms(np.hstack((proposals, scores)), nms_thresh)      # :138
        if post_nms_topN > 0:
            keep_nms = keep_nms[:post_nms_topN]
        keep_nms = np.asarray(keep_nms, dtype=np.int64)
        proposals = proposals[keep_nms, :]
        scores = scores[keep_nms]
        batch_inds = np.ones((proposals.shape[0], 1), dtype=np.float32) * i
        t_blob = np.hstack((batch_inds, proposals.astype(np.float32, copy=False)))
        blob = np.concatenate((blob, t_blob))
        parts.append(dict(decoded=all_props, pre_nms_anchor_idx=pre_idx,
                          anchor_idx=pre_idx[keep_nms], scores=scores.ravel().copy()))
    if return_parts:
        return blob, parts
    return blob


# ---------------------------------------------------------------- anchor_target_layer_tf_bus.py
def anchor_labels(height, width, gt_boxes, im_info, feat_stride=16, anchor_scales=(8, 16, 32),
                  dataset="SNUBH", cfg=None):
    """Pre-subsample labels of anchor_target_layer[_joint] for ONE image.

    Follows anchor_target_layer_tf_bus.py:410-467 (SNUBH) and :470-509 (other datasets):
    inside filter, IoU (fp64) against fg GT rows, uni-directional overlap against the
    explicit background rows, labels -1/0/1.  Returns a dict with inds_inside, labels
    (len(inds_inside),) f32, argmax_overlaps, max_overlaps (fg) -- everything that is
    deterministic before the npr.choice subsampling at :512-527."""
    cfg = cfg or TRAIN
    all_anchors, A = shifted_anchors(height, width, feat_stride, anchor_scales)
    t_im_info = im_info
    t_gt_boxes = gt_boxes
    inds_inside = np.where(
        (all_anchors[:, 0] >= 0) & (all_anchors[:, 1] >= 0) &
        (all_anchors[:, 2] < t_im_info[1]) & (all_anchors[:, 3] < t_im_info[0]))[0]
    anchors = all_anchors[inds_inside, :]
    labels = np.empty((len(inds_inside),), dtype=np.float32)
    labels.fill(-1)
    out = dict(inds_inside=inds_inside, anchors=anchors, A=A, total_anchors=all_anchors.shape[0])
    if dataset == "SNUBH":
        b_pos = np.transpose(t_gt_boxes[:, 4] != 0)
        num_pos = int(np.sum(b_pos))
        exist_neg = (t_gt_boxes.shape[0] != num_pos)
        overlaps_pos = _overlaps(np.ascontiguousarray(anchors, dtype=np.float64),
                                 np.ascontiguousarray(t_gt_boxes[:num_pos, :4], dtype=np.float64))
        argmax_overlaps_pos = overlaps_pos.argmax(axis=1)
        max_overlaps_pos = overlaps_pos[np.arange(len(inds_inside)), argmax_overlaps_pos]
        gt_argmax_overlaps_pos = overlaps_pos.argmax(axis=0)
        gt_max_overlaps_pos = overlaps_pos[gt_argmax_overlaps_pos, np.arange(overlaps_pos.shape[1])]
        gt_argmax_overlaps_pos = np.where(overlaps_pos == gt_max_overlaps_pos)[0]
        if exist_neg:
            overlaps_neg = _overlaps(np.ascontiguousarray(anchors, dtype=np.float64),
                                     np.ascontiguousarray(t_gt_boxes[num_pos:, :4], dtype=np.float64),
                                     ui=True)
            argmax_overlaps_neg = overlaps_neg.argmax(axis=1)
            max_overlaps_neg = overlaps_neg[np.arange(len(inds_inside)), argmax_overlaps_neg]
            out["max_overlaps_neg"] = max_overlaps_neg
        if not cfg["RPN_CLOBBER_POSITIVES"] and exist_neg:
            labels[max_overlaps_neg >= cfg["RPN_POSITIVE_OVERLAP"]] = 0       # :461
        labels[gt_argmax_overlaps_pos] = 1                                   # :464
        labels[max_overlaps_pos >= cfg["RPN_POSITIVE_OVERLAP"]] = 1           # :467
        out.update(argmax_overlaps=argmax_overlaps_pos, max_overlaps=max_overlaps_pos,
                   gt_max_overlaps=gt_max_overlaps_pos, num_pos=num_pos)
    else:
        if dataset == "SNUBH_FG":
            num_pos = int(np.sum(t_gt_boxes[:, 4] != 0))
            t_gt_boxes = t_gt_boxes[:num_pos, :]
        overlaps = _overlaps(np.ascontiguousarray(anchors, dtype=np.float64),
                             np.ascontiguousarray(t_gt_boxes[:, :4], dtype=np.float64))
        argmax_overlaps = overlaps.argmax(axis=1)
        max_overlaps = overlaps[np.arange(len(inds_inside)), argmax_overlaps]
        gt_argmax_overlaps = overlaps.argmax(axis=0)
        gt_max_overlaps = overlaps[gt_argmax_overlaps, np.arange(overlaps.shape[1])]
        gt_argmax_overlaps = np.where(overlaps == gt_max_overlaps)[0]
        if not cfg["RPN_CLOBBER_POSITIVES"]:
            labels[max_overlaps < cfg["RPN_NEGATIVE_OVERLAP"]] = 0           # :499
        labels[gt_argmax_overlaps] = 1
        labels[max_overlaps >= cfg["RPN_POSITIVE_OVERLAP"]] = 1
        if cfg["RPN_CLOBBER_POSITIVES"]:
            labels[max_overlaps < cfg["RPN_NEGATIVE_OVERLAP"]] = 0
        out.update(argmax_overlaps=argmax_overlaps, max_overlaps=max_overlaps,
                   gt_max_overlaps=gt_max_overlaps, num_pos=t_gt_boxes.shape[0])
    out["labels"] = labels
    out["gt_used"] = t_gt_boxes
    return out


def anchor_targets_from_labels(lab, rng, cfg=None):
    """anchor_target_layer_tf_bus.py:512-571 for one image: subsample with the injected
    RNG (`rng.choice` stands for npr.choice), regression targets, weights, _unmap."""
    cfg = cfg or TRAIN
    labels = lab["labels"].copy()
    anchors, inds_inside, total = lab["anchors"], lab["inds_inside"], lab["total_anchors"]
    num_fg = int(cfg["RPN_FG_FRACTION"] * cfg["RPN_BATCHSIZE"])
    fg_inds = np.where(labels == 1)[0]
    if len(fg_inds) > num_fg:
        labels[rng.choice(fg_inds, size=(len(fg_inds) - num_fg), replace=False)] = -1
    num_bg = cfg["RPN_BATCHSIZE"] - np.sum(labels == 1)
    bg_inds = np.where(labels == 0)[0]
    if len(bg_inds) > num_bg:
        labels[rng.choice(bg_inds, size=(len(bg_inds) - num_bg), replace=False)] = -1
    gt = lab["gt_used"]
    bbox_targets = bbox_transform(anchors, gt[lab["argmax_overlaps"], :][:, :4]).astype(
        np.float32, copy=False)                                              # :533, :645-653
    inside = np.zeros((len(inds_inside), 4), dtype=np.float32)
    inside[labels == 1, :] = np.array(cfg["RPN_BBOX_INSIDE_WEIGHTS"])
    outside = np.zeros((len(inds_inside), 4), dtype=np.float32)
    num_examples = np.sum(labels >= 0)
    w = np.ones((1, 4)) * 1.0 / num_examples
    outside[labels == 1, :] = w
    outside[labels == 0, :] = w

    def unmap(data, fill):
        if data.ndim == 1:
            ret = np.empty((total,), dtype=np.float32)
            ret.fill(fill)
            ret[inds_inside] = data
        else:
            ret = np.empty((total,) + data.shape[1:], dtype=np.float32)
            ret.fill(fill)
            ret[inds_inside, :] = data
        return ret
    return unmap(labels, -1), unmap(bbox_targets, 0), unmap(inside, 0), unmap(outside, 0)


# ---------------------------------------------------------------- proposal_target_layer_tf_bus.py
def sample_rois_deterministic(all_rois, gt_boxes, cfg=None):
    """The deterministic part of _sample_rois (proposal_target_layer_tf_bus.py:233-254):
    IoU, gt assignment, max overlap, candidate fg/bg index sets (before npr.choice)."""
    cfg = cfg or TRAIN
    overlaps = _overlaps(np.ascontiguousarray(all_rois[:, 1:5], dtype=np.float64),
                         np.ascontiguousarray(gt_boxes[:, :4], dtype=np.float64))
    gt_assignment = overlaps.argmax(axis=1)
    max_overlaps = overlaps.max(axis=1)
    labels = gt_boxes[gt_assignment, 4]
    fg_inds = np.where(max_overlaps >= cfg["FG_THRESH"])[0]
    bg_inds = np.where((max_overlaps < cfg["BG_THRESH_HI"]) &
                       (max_overlaps >= cfg["BG_THRESH_LO"]))[0]
    return dict(overlaps=overlaps, gt_assignment=gt_assignment, max_overlaps=max_overlaps,
                labels=labels, fg_inds=fg_inds, bg_inds=bg_inds)


def sample_rois(all_rois, gt_boxes, fg_rois_per_image, rois_per_image, num_classes, rng,
                cfg=None):
    """_sample_rois (proposal_target_layer_tf_bus.py:228-280) with an injected RNG."""
    cfg = cfg or TRAIN
    d = sample_rois_deterministic(all_rois, gt_boxes, cfg)
    fg_inds, bg_inds = d["fg_inds"], d["bg_inds"]
    fg_n = min(fg_rois_per_image, fg_inds.size)
    if fg_inds.size > 0:
        fg_inds = rng.choice(fg_inds, size=fg_n, replace=False)
    bg_n = min(rois_per_image - fg_n, bg_inds.size)
    if bg_inds.size > 0:
        bg_inds = rng.choice(bg_inds, size=bg_n, replace=False)
    keep_inds = np.append(fg_inds, bg_inds).astype(np.int64)
    labels = d["labels"][keep_inds].copy()
    labels[fg_n:] = 0
    rois = all_rois[keep_inds]
    targets = bbox_transform(rois[:, 1:5], gt_boxes[d["gt_assignment"][keep_inds], :4])
    data = np.hstack((labels[:, np.newaxis], targets)).astype(np.float32, copy=False)
    bbox_targets = np.zeros((labels.size, 4 * num_classes), dtype=np.float32)
    inside = np.zeros(bbox_targets.shape, dtype=np.float32)
    for ind in np.where(data[:, 0] > 0)[0]:
        cls = int(data[ind, 0])
        bbox_targets[ind, 4 * cls:4 * cls + 4] = data[ind, 1:]
        inside[ind, 4 * cls:4 * cls + 4] = cfg["BBOX_INSIDE_WEIGHTS"]
    return labels, rois, bbox_targets, inside, keep_inds


# --------------------------------------------------------------- detection post-processing
def im_detect_boxes(rois, bbox_pred, im_shape, im_scale, bbox_reg=True, num_classes=None):
    """Tail of im_detect (fast_rcnn/test_bus.py:207-223): RoIs of ONE image back to the
    original image frame, per-class box regression, clip.

    rois [R,5] f32 (batch, x1,y1,x2,y2) in the scaled frame; bbox_pred [R,4K] f32;
    im_shape = shape of the unscaled image (H, W[, ch]); im_scale python float.
    Returns pred_boxes [R,4K] f32."""
    boxes = rois[:, 1:5] / im_scale                                   # :209 (f32 / py float -> f32)
    if bbox_reg:
        pred_boxes = bbox_transform_inv(boxes, bbox_pred)             # :222
        # _clip_boxes (:124-134): one-sided clips, unlike bbox_transform.clip_boxes
        pred_boxes[:, 0::4] = np.maximum(pred_boxes[:, 0::4], 0)
        pred_boxes[:, 1::4] = np.maximum(pred_boxes[:, 1::4], 0)
        pred_boxes[:, 2::4] = np.minimum(pred_boxes[:, 2::4], im_shape[1] - 1)
        pred_boxes[:, 3::4] = np.minimum(pred_boxes[:, 3::4], im_shape[0] - 1)
    else:
        pred_boxes = np.tile(boxes, (1, num_classes))                 # :226
    return pred_boxes


def detections_postprocess(scores, boxes, thresh=0.05, nms_thresh=None, max_per_image=300,
                         cls_agnostic_nms=False):
    """Per-image body of test_net (fast_rcnn/test_bus.py:360-401).

    scores [R,K] f32, boxes [R,4K] f32 -> list over classes (entry 0 = background = empty)
    of [n_j,5] f32 (x1,y1,x2,y2,score) in descending-score order."""
    nms_thresh = TEST['NMS'] if nms_thresh is None else nms_thresh
    K = scores.shape[1]
    out = [np.zeros((0, 5), np.float32) for _ in range(K)]
    for j in range(1, K):                                             # :360 skip background
        inds = np.where(scores[:, j] > thresh)[0]                     # :361
        cls_scores = scores[inds, j]
        cls_boxes = boxes[inds, j * 4:(j + 1) * 4]
        cls_dets = np.hstack((cls_boxes, cls_scores[:, np.newaxis])).astype(np.float32, copy=False)
        keep = _nms(cls_dets, nms_thresh)                             # :366 utils.cython_nms.nms
        out[j] = cls_dets[keep, :]
    if cls_agnostic_nms:                                              # :371-386
        all_dets = np.zeros((0, 6), dtype=np.float32)
        for j in range(1, K):
            all_dets = np.concatenate(
                (all_dets, np.hstack((out[j], j * np.ones((out[j].shape[0], 1), dtype=np.float32)))),
                axis=0)
        keep = _nms(all_dets, nms_thresh)
        all_dets = all_dets[keep, :]
        for j in range(1, K):
            inds = np.where(all_dets[:, 5] == j)[0]
            out[j] = all_dets[inds, :5]
    if max_per_image > 0:                                             # :394-401
        image_scores = np.hstack([out[j][:, -1] for j in range(1, K)])
        if len(image_scores) > max_per_image:
            image_thresh = np.sort(image_scores)[-max_per_image]
            for j in range(1, K):
                keep = np.where(out[j][:, -1] >= image_thresh)[0]
                out[j] = out[j][keep, :]
    return out


# --------------------------------------------------------------- VOC / CorLoc / FROC evaluation
def voc_ap(rec, prec, use_07_metric=False):
    """datasets/voc_eval_bus.py:37-66."""
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


def voc_eval_arrays(image_ids, confidence, BB, gt_bbox, gt_difficult, ovthresh=0.5,
                    use_07_metric=False, score_thresh=0.5):
    """voc_eval_bus (datasets/voc_eval_bus.py:68-281) for ONE class, from the point where the
    reference has parsed its text/XML files (:129-160): image_ids [nd] ints (index into the
    image list, in file order), confidence [nd], BB [nd,4]; gt_bbox[i] [G_i,4] and
    gt_difficult[i] [G_i] bool for every image i (objects of this class only).
    Returns (rec, prec, ap, ni, nok, num_all_fps, num_fp_per_img) like :281 (arr_ok omitted)."""
    n_images = len(gt_bbox)
    class_recs = [dict(bbox=np.asarray(gt_bbox[i], dtype=float).reshape(-1, 4),
                       difficult=np.asarray(gt_difficult[i], dtype=bool).reshape(-1),
                       det=[False] * len(gt_bbox[i])) for i in range(n_images)]
    npos = sum(int(np.sum(~r['difficult'])) for r in class_recs)                  # :135
    image_ids = np.asarray(image_ids, dtype=np.int64)
    confidence = np.asarray(confidence, dtype=float)
    BB = np.asarray(BB, dtype=float).reshape(-1, 4)
    if len(image_ids) == 0:                                                          # :276-279
        return -1, -1, -1, 0, 0, 0, [0] * n_images
    sorted_ind = np.argsort(-confidence)                                             # :155
    sorted_scores = np.sort(-confidence)
    BB = BB[sorted_ind, :]
    image_ids = image_ids[sorted_ind]
    # CorLoc :161-204
    ni = nok = 0
    for i in range(n_images):
        BBGT = class_recs[i]['bbox']
        if BBGT.shape[0] > 0:
            ni += 1
            inds = np.where((image_ids == i) & (sorted_scores <= -score_thresh))[0]
            if len(inds) == 0:
                continue
            bb = BB[inds, :]
            bok = False
            for j in range(BBGT.shape[0]):
                ixmin = np.maximum(bb[:, 0], BBGT[j, 0])
                iymin = np.maximum(bb[:, 1], BBGT[j, 1])
                ixmax = np.minimum(bb[:, 2], BBGT[j, 2])
                iymax = np.minimum(bb[:, 3], BBGT[j, 3])
                iw = np.maximum(ixmax - ixmin + 1., 0.)
                ih = np.maximum(iymax - iymin + 1., 0.)
                inters = iw * ih
                uni = ((BBGT[j, 2] - BBGT[j, 0] + 1.) * (BBGT[j, 3] - BBGT[j, 1] + 1.) +
                       (bb[:, 2] - bb[:, 0] + 1.) * (bb[:, 3] - bb[:, 1] + 1.) - inters)
                if np.max(inters / uni) > ovthresh:
                    bok = True
            if bok:
                nok += 1
    # TP / FP / FROC marking :206-247
    nd = len(image_ids)
    tp = np.zeros(nd)
    fp = np.zeros(nd)
    fp_froc = np.zeros(nd)
    for d in range(nd):
        R = class_recs[image_ids[d]]
        bb = BB[d, :]
        ovmax = -np.inf
        BBGT = R['bbox']
        if BBGT.size > 0:
            ixmin = np.maximum(BBGT[:, 0], bb[0])
            iymin = np.maximum(BBGT[:, 1], bb[1])
            ixmax = np.minimum(BBGT[:, 2], bb[2])
            iymax = np.minimum(BBGT[:, 3], bb[3])
            iw = np.maximum(ixmax - ixmin + 1., 0.)
            ih = np.maximum(iymax - iymin + 1., 0.)
            inters = iw * ih
            uni = ((bb[2] - bb[0] + 1.) * (bb[3] - bb[1] + 1.) +
                   (BBGT[:, 2] - BBGT[:, 0] + 1.) * (BBGT[:, 3] - BBGT[:, 1] + 1.) - inters)
            overlaps = inters / uni
            ovmax = np.max(overlaps)
            jmax = np.argmax(overlaps)
        if ovmax > ovthresh:
            if not R['difficult'][jmax]:
                if not R['det'][jmax]:
                    tp[d] = 1.
                    R['det'][jmax] = 1
                else:
                    fp[d] = 1.
        else:
            fp[d] = 1.
        if sorted_scores[d] <= -score_thresh:
            if ovmax <= ovthresh:
                fp_froc[d] = 1.
    num_all_fps = np.sum(fp_froc)
    num_fp_per_img = [int(np.sum(fp_froc[image_ids == i])) for i in range(n_images)]   # :252-262
    fp = np.cumsum(fp)                                                                   # :265-271
    tp = np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    ap = voc_ap(rec, prec, use_07_metric)
    return rec, prec, ap, ni, nok, num_all_fps, num_fp_per_img
