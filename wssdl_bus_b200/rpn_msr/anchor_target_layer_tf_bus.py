"""rpn_msr/anchor_target_layer_tf_bus.py twin: anchor_target_layer (:19-303) and
anchor_target_layer_joint (:328-628).

Device: inside filter, fp64 IoU / uni-directional overlap, row/column maxima, labels
(csrc/anchor_target.cu, one launch pair for the batch) and the regression targets
(bbox_transform).  Host: the npr.choice subsampling (:512-527), exactly where and how the
reference draws it, so a seeded run consumes the same random stream.
"""
import numpy as np
import numpy.random as npr

from wssdl_bus_b200 import ops
from wssdl_bus_b200.fast_rcnn.config import cfg
from wssdl_bus_b200.rpn_msr.generate_anchors import generate_anchors, shifted_anchors

_MODES = {'SNUBH': 0, 'SNUBH_FG': 1}


def _stride(feat_stride):
    return int(feat_stride[0]) if isinstance(feat_stride, (list, tuple, np.ndarray)) else int(feat_stride)


def _layer(rpn_cls_score, gt_boxes, num_gt_boxes, im_info, n_supervised, n_ws, _feat_stride,
           anchor_scales, dataset):
    base = generate_anchors(scales=np.array(anchor_scales))
    A = base.shape[0]
    height, width = rpn_cls_score.shape[1:3]
    fs = _stride(_feat_stride)
    gt = np.ascontiguousarray(gt_boxes[:n_supervised], dtype=np.float32)
    ng = np.ascontiguousarray(num_gt_boxes[:n_supervised], dtype=np.int32)
    info = np.ascontiguousarray(im_info[:n_supervised], dtype=np.float32)
    labels_d, argmax_d, _ = ops.anchor_labels(
        gt, ng, info, height, width, base, fs, dataset_mode=_MODES.get(dataset, 2),
        positive_overlap=cfg.TRAIN.RPN_POSITIVE_OVERLAP,
        negative_overlap=cfg.TRAIN.RPN_NEGATIVE_OVERLAP,
        clobber_positives=cfg.TRAIN.RPN_CLOBBER_POSITIVES, want_max_overlap=False)
    labels_all = labels_d.cpu().numpy()
    argmax_all = argmax_d.cpu().numpy()
    all_anchors = shifted_anchors(height, width, fs, base)
    total = all_anchors.shape[0]

    rpn_labels, rpn_t, rpn_iw, rpn_ow = [], [], [], []
    for i in range(n_supervised):
        inds_inside = np.where(argmax_all[i] >= 0)[0]
        labels = labels_all[i][inds_inside].copy()
        # subsample positives / negatives (:512-527), host RNG like the reference
        num_fg = int(cfg.TRAIN.RPN_FG_FRACTION * cfg.TRAIN.RPN_BATCHSIZE)
        fg_inds = np.where(labels == 1)[0]
        if len(fg_inds) > num_fg:
            labels[npr.choice(fg_inds, size=(len(fg_inds) - num_fg), replace=False)] = -1
        num_bg = cfg.TRAIN.RPN_BATCHSIZE - np.sum(labels == 1)
        bg_inds = np.where(labels == 0)[0]
        if len(bg_inds) > num_bg:
            labels[npr.choice(bg_inds, size=(len(bg_inds) - num_bg), replace=False)] = -1
        # regression targets against the best fg GT (:533, :645-653), on the device
        t_gt = gt[i, :ng[i], :]
        if dataset in ('SNUBH', 'SNUBH_FG'):
            t_gt = t_gt[:int(np.sum(t_gt[:, 4] != 0)), :]
        anchors = all_anchors[inds_inside]
        if len(inds_inside) and t_gt.shape[0]:
            targets = ops.bbox_transform(anchors.astype(np.float32),
                                         t_gt[argmax_all[i][inds_inside], :4])
        else:
            targets = np.zeros((len(inds_inside), 4), np.float32)
        inside_w = np.zeros((len(inds_inside), 4), dtype=np.float32)
        inside_w[labels == 1, :] = np.array(cfg.TRAIN.RPN_BBOX_INSIDE_WEIGHTS)
        outside_w = np.zeros((len(inds_inside), 4), dtype=np.float32)
        if cfg.TRAIN.RPN_POSITIVE_WEIGHT < 0:
            num_examples = np.sum(labels >= 0)
            pos_w = neg_w = np.ones((1, 4)) * 1.0 / num_examples
        else:
            pos_w = cfg.TRAIN.RPN_POSITIVE_WEIGHT / np.sum(labels == 1)
            neg_w = (1.0 - cfg.TRAIN.RPN_POSITIVE_WEIGHT) / np.sum(labels == 0)
        outside_w[labels == 1, :] = pos_w
        outside_w[labels == 0, :] = neg_w

        def unmap(data, fill):
            ret = np.empty((total,) + data.shape[1:], dtype=np.float32)
            ret.fill(fill)
            ret[inds_inside] = data
            return ret
        lab = unmap(labels, -1).reshape((1, height, width, A)).transpose(0, 3, 1, 2)
        rpn_labels.append(lab.reshape((1, 1, A * height, width)))
        rpn_t.append(unmap(targets, 0).reshape((1, height, width, A * 4)).transpose(0, 3, 1, 2))
        rpn_iw.append(unmap(inside_w, 0).reshape((1, height, width, A * 4)).transpose(0, 3, 1, 2))
        rpn_ow.append(unmap(outside_w, 0).reshape((1, height, width, A * 4)).transpose(0, 3, 1, 2))
    if n_ws:
        # weakly-supervised images carry no RPN supervision (:613-626)
        lab = np.empty((n_ws, 1, A * height, width), dtype=np.float32)
        lab.fill(-1)
        rpn_labels.append(lab)
        z = np.zeros((n_ws, A * 4, height, width), dtype=np.float32)
        rpn_t.append(z)
        rpn_iw.append(z.copy())
        rpn_ow.append(z.copy())

    def cat(parts, shape):
        return np.concatenate(parts) if parts else np.zeros(shape, np.float32)
    return (cat(rpn_labels, (0, 1, A * height, width)), cat(rpn_t, (0, A * 4, height, width)),
            cat(rpn_iw, (0, A * 4, height, width)), cat(rpn_ow, (0, A * 4, height, width)))


def anchor_target_layer(rpn_cls_score, gt_boxes, num_gt_boxes, im_info, data=None,
                        _feat_stride=[16, ], anchor_scales=[4, 8, 16, 32], dataset='SNUBH'):
    """:19-303 -- every image of the batch is supervised."""
    return _layer(rpn_cls_score, gt_boxes, num_gt_boxes, im_info, rpn_cls_score.shape[0], 0,
                  _feat_stride, anchor_scales, dataset)


def anchor_target_layer_ws(rpn_cls_score, gt_boxes, num_gt_boxes, im_info, data=None,
                           _feat_stride=[16, ], anchor_scales=[4, 8, 16, 32]):
    """:306-325 -- weakly supervised batch: all labels -1, zero targets and weights."""
    A = generate_anchors(scales=np.array(anchor_scales)).shape[0]
    height, width = rpn_cls_score.shape[1:3]
    batch_size = rpn_cls_score.shape[0]
    labels = np.empty((batch_size, 1, A * height, width), dtype=np.float32)
    labels.fill(-1)
    z = np.zeros((batch_size, A * 4, height, width), dtype=np.float32)
    return labels, z, z.copy(), z.copy()


def anchor_target_layer_joint(rpn_cls_score, gt_boxes, num_gt_boxes, im_info, data=None,
                              is_training=True, _feat_stride=[16, ], anchor_scales=[4, 8, 16, 32],
                              dataset='SNUBH'):
    """:328-628 -- the first IMS_PER_BATCH images are supervised, the WS_IMS_PER_BATCH
    weakly-supervised ones get all-don't-care labels when training."""
    n_ws = cfg.TRAIN.WS_IMS_PER_BATCH if is_training else 0
    return _layer(rpn_cls_score, gt_boxes, num_gt_boxes, im_info, cfg.TRAIN.IMS_PER_BATCH, n_ws,
                  _feat_stride, anchor_scales, dataset)
