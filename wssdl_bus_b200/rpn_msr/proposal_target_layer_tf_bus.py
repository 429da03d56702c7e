"""rpn_msr/proposal_target_layer_tf_bus.py twin: proposal_target_layer (:15-97),
proposal_target_layer_joint (:99-184), _sample_rois (:228-280).

Device: the RoI x GT IoU matrix (fp64, bit-exact) and the regression targets.  Host: the
fg/bg npr.choice draws (:250, :261) so the random stream matches the reference's.
"""
import numpy as np
import numpy.random as npr

from wssdl_bus_b200 import ops
from wssdl_bus_b200.fast_rcnn.config import cfg


def _get_bbox_regression_labels(bbox_target_data, num_classes):
    """:187-210 -- expand N x (cls, tx, ty, tw, th) into the 4-of-4K layout."""
    clss = bbox_target_data[:, 0]
    bbox_targets = np.zeros((clss.size, 4 * num_classes), dtype=np.float32)
    bbox_inside_weights = np.zeros(bbox_targets.shape, dtype=np.float32)
    for ind in np.where(clss > 0)[0]:
        start = 4 * int(clss[ind])
        bbox_targets[ind, start:start + 4] = bbox_target_data[ind, 1:]
        bbox_inside_weights[ind, start:start + 4] = cfg.TRAIN.BBOX_INSIDE_WEIGHTS
    return bbox_targets, bbox_inside_weights


def _compute_targets(ex_rois, gt_rois, labels):
    """:213-226."""
    targets = ops.bbox_transform(np.ascontiguousarray(ex_rois, np.float32),
                                 np.ascontiguousarray(gt_rois, np.float32))
    if cfg.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED:
        targets = ((targets - np.array(cfg.TRAIN.BBOX_NORMALIZE_MEANS))
                   / np.array(cfg.TRAIN.BBOX_NORMALIZE_STDS))
    return np.hstack((labels[:, np.newaxis], targets)).astype(np.float32, copy=False)


def _sample_rois(all_rois, gt_boxes, fg_rois_per_image, rois_per_image, num_classes):
    """:228-280."""
    overlaps = ops.bbox_overlaps(np.ascontiguousarray(all_rois[:, 1:5], dtype=np.float64),
                                 np.ascontiguousarray(gt_boxes[:, :4], dtype=np.float64))
    gt_assignment = overlaps.argmax(axis=1)
    max_overlaps = overlaps.max(axis=1)
    labels = gt_boxes[gt_assignment, 4]
    fg_inds = np.where(max_overlaps >= cfg.TRAIN.FG_THRESH)[0]
    fg_rois_per_this_image = min(fg_rois_per_image, fg_inds.size)
    if fg_inds.size > 0:
        fg_inds = npr.choice(fg_inds, size=fg_rois_per_this_image, replace=False)
    bg_inds = np.where((max_overlaps < cfg.TRAIN.BG_THRESH_HI) &
                       (max_overlaps >= cfg.TRAIN.BG_THRESH_LO))[0]
    bg_rois_per_this_image = min(rois_per_image - fg_rois_per_this_image, bg_inds.size)
    if bg_inds.size > 0:
        bg_inds = npr.choice(bg_inds, size=bg_rois_per_this_image, replace=False)
    keep_inds = np.append(fg_inds, bg_inds).astype(np.int64)
    labels = labels[keep_inds]
    labels[fg_rois_per_this_image:] = 0
    rois = all_rois[keep_inds]
    bbox_target_data = _compute_targets(rois[:, 1:5], gt_boxes[gt_assignment[keep_inds], :4], labels)
    bbox_targets, bbox_inside_weights = _get_bbox_regression_labels(bbox_target_data, num_classes)
    be_sel = np.zeros((all_rois.shape[0],), dtype=bool)
    be_sel[keep_inds] = True
    return labels, rois, bbox_targets, bbox_inside_weights, be_sel


def _sample_rois_ws(all_rois, num_classes):
    """:282-295 -- weakly supervised image: every RoI, zero labels/targets."""
    labels = np.zeros((all_rois.shape[0],), dtype=np.float32)
    bbox_targets = np.zeros((all_rois.shape[0], 4 * num_classes), dtype=np.float32)
    return labels, all_rois, bbox_targets, np.zeros(bbox_targets.shape, dtype=np.float32)


def _layer(rpn_rois, gt_boxes, num_gt_boxes, _num_classes, n_supervised, ws_range, add_gt,
           all_ws=False):
    batch_rois = np.zeros((0, 5), dtype=rpn_rois.dtype)
    batch_labels = np.zeros((0, 1), dtype=gt_boxes.dtype)
    batch_t = np.zeros((0, _num_classes * 4), dtype=np.float32)
    batch_iw = np.zeros((0, _num_classes * 4), dtype=np.float32)
    batch_ow = np.zeros((0, _num_classes * 4), dtype=np.float32)
    for i in range(n_supervised):
        all_rois = rpn_rois[rpn_rois[:, 0] == i, :]
        t_gt_boxes = gt_boxes[i, :num_gt_boxes[i], :]
        num_pos = int(np.sum(t_gt_boxes[:, 4] != 0))
        temp_gt_boxes = t_gt_boxes[:num_pos, :]
        if add_gt:   # include GT boxes among the candidates (:127-132)
            idx = np.ones((temp_gt_boxes.shape[0], 1), dtype=temp_gt_boxes.dtype) * i
            all_rois = np.vstack((all_rois, np.hstack((idx, temp_gt_boxes[:, :-1]))))
        rois_per_image = cfg.TRAIN.BATCH_SIZE // 1     # py2 integer division (:135)
        fg_rois_per_image = int(np.round(cfg.TRAIN.FG_FRACTION * rois_per_image))
        if all_ws:
            labels, rois, bbox_targets, bbox_inside_weights = _sample_rois_ws(all_rois, _num_classes)
        else:
            labels, rois, bbox_targets, bbox_inside_weights, _ = _sample_rois(
                all_rois, temp_gt_boxes, fg_rois_per_image, rois_per_image, _num_classes)
        rois = rois.reshape(-1, 5)
        labels = labels.reshape(-1, 1)
        bbox_targets = bbox_targets.reshape(-1, _num_classes * 4)
        bbox_inside_weights = bbox_inside_weights.reshape(-1, _num_classes * 4)
        bbox_outside_weights = np.array(bbox_inside_weights > 0).astype(np.float32)
        batch_rois = np.concatenate((batch_rois, rois))
        batch_labels = np.concatenate((batch_labels, labels))
        batch_t = np.concatenate((batch_t, bbox_targets))
        batch_iw = np.concatenate((batch_iw, bbox_inside_weights))
        batch_ow = np.concatenate((batch_ow, bbox_outside_weights))
    for i in ws_range:   # weakly supervised images: every proposal, un-sampled (:162-182)
        batch_rois = np.concatenate((batch_rois, rpn_rois[rpn_rois[:, 0] == i, :].reshape(-1, 5)))
    return batch_rois, batch_labels, batch_t, batch_iw, batch_ow


def proposal_target_layer(rpn_rois, gt_boxes, num_gt_boxes, _num_classes, is_training=True,
                          is_ws=False):
    """:15-97 -- GT boxes join the candidates when training a supervised batch (:45-50); a
    weakly supervised training batch passes every RoI through un-sampled (:62-65)."""
    return _layer(rpn_rois, gt_boxes, num_gt_boxes, _num_classes, gt_boxes.shape[0], (),
                  bool(is_training) and not bool(is_ws), all_ws=bool(is_training) and bool(is_ws))


def proposal_target_layer_joint(rpn_rois, gt_boxes, num_gt_boxes, _num_classes, is_training):
    """:99-184."""
    n = cfg.TRAIN.IMS_PER_BATCH
    ws = range(n, n + cfg.TRAIN.WS_IMS_PER_BATCH) if is_training else ()
    return _layer(rpn_rois, gt_boxes, num_gt_boxes, _num_classes, n, ws, bool(is_training))
