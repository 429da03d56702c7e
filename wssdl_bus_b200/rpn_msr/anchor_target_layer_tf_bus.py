"""rpn_msr/anchor_target_layer_tf_bus.py twin: anchor_target_layer (:19-303) and
anchor_target_layer_joint (:328-628), device resident.

Two launches for the whole batch: labels (csrc/anchor_target.cu: inside filter, fp64 IoU /
uni-directional overlap, row / column maxima, label rules) and targets (csrc/targets.cu: fg / bg
subsampling, fp64 regression targets, inside / outside weights, `_unmap`, final layouts).
Nothing is copied per image.  The subsampling (:512-527) has two modes:
  sampler="host"    the reference's npr.choice draws, from numpy.random in the reference's order, so
                    a seeded run consumes the same random stream and reproduces its outputs: the
                    device reports the two population sizes per image (one [B,2] int32 read, the
                    only host trip), the host draws RANKS, the device applies them;
  sampler="philox"  drawn on the device from a documented Philox4x32-10 stream (include/
                    wssdl_b200.h): no host trip at all.
"""
import numpy as np
import numpy.random as npr
import torch

from wssdl_bus_b200 import ops
from wssdl_bus_b200.fast_rcnn.config import cfg
from wssdl_bus_b200.rpn_msr.generate_anchors import generate_anchors

_MODES = {'SNUBH': 0, 'SNUBH_FG': 1}


def _stride(feat_stride):
    return int(feat_stride[0]) if isinstance(feat_stride, (list, tuple, np.ndarray)) else int(feat_stride)


def _layer(rpn_cls_score, gt_boxes, num_gt_boxes, im_info, n_supervised, n_ws, _feat_stride,
           anchor_scales, dataset, sampler="host", seed=0, return_device=False):
    base = generate_anchors(scales=np.array(anchor_scales))
    A = base.shape[0]
    height, width = rpn_cls_score.shape[1:3]
    fs = _stride(_feat_stride)
    as_np = not torch.is_tensor(gt_boxes)
    gt = ops._cuda(gt_boxes[:n_supervised], torch.float32)
    ng = ops._cuda(num_gt_boxes[:n_supervised], torch.int32, gt.device)
    info = ops._cuda(im_info[:n_supervised], torch.float32, gt.device)
    num_fg = int(cfg.TRAIN.RPN_FG_FRACTION * cfg.TRAIN.RPN_BATCHSIZE)            # :513
    batchsize = int(cfg.TRAIN.RPN_BATCHSIZE)
    if n_supervised:
        labels_d, argmax_d, _ = ops.anchor_labels(
            gt, ng, info, height, width, base, fs, dataset_mode=_MODES.get(dataset, 2),
            positive_overlap=cfg.TRAIN.RPN_POSITIVE_OVERLAP,
            negative_overlap=cfg.TRAIN.RPN_NEGATIVE_OVERLAP,
            clobber_positives=cfg.TRAIN.RPN_CLOBBER_POSITIVES, want_max_overlap=False)
    else:
        labels_d = torch.zeros((0, height * width * A), dtype=torch.float32, device=gt.device)
        argmax_d = torch.zeros((0, height * width * A), dtype=torch.int32, device=gt.device)
    kw = {}
    if sampler == "host":
        # the population sizes are all the host RNG needs: npr.choice(pop, size, replace=False)
        # draws permutation(len(pop))[:size], whatever pop holds (:514-518, :524-528)
        counts = ops.anchor_label_counts(labels_d).cpu().numpy() if n_supervised else np.zeros((0, 2), int)
        ranks, off = [], [0]
        for i in range(n_supervised):
            n_fg, n_bg = int(counts[i, 0]), int(counts[i, 1])
            if n_fg > num_fg:
                ranks.append(npr.choice(n_fg, size=(n_fg - num_fg), replace=False))
            off.append(off[-1] + (n_fg - num_fg if n_fg > num_fg else 0))
            num_bg = batchsize - min(n_fg, num_fg)
            if n_bg > num_bg:
                ranks.append(npr.choice(n_bg, size=(n_bg - num_bg), replace=False))
            off.append(off[-1] + (n_bg - num_bg if n_bg > num_bg else 0))
        kw = dict(ranks=np.concatenate(ranks).astype(np.int32) if ranks else np.zeros((0,), np.int32),
                  rank_off=np.asarray(off, np.int32))
    elif sampler == "philox":
        kw = dict(seed=int(seed))
    else:
        raise ValueError("sampler must be 'host' or 'philox'")
    labels, targets, inside, outside, _ = ops.anchor_targets(
        labels_d, argmax_d, gt, n_supervised + n_ws, height, width, base, fs, num_fg, batchsize,
        inside_weights=cfg.TRAIN.RPN_BBOX_INSIDE_WEIGHTS,
        positive_weight=cfg.TRAIN.RPN_POSITIVE_WEIGHT, **kw)
    if as_np and not return_device:
        return tuple(t.cpu().numpy() for t in (labels, targets, inside, outside))
    return labels, targets, inside, outside


def anchor_target_layer(rpn_cls_score, gt_boxes, num_gt_boxes, im_info, data=None,
                        _feat_stride=[16, ], anchor_scales=[4, 8, 16, 32], dataset='SNUBH',
                        sampler="host", seed=0, return_device=False):
    """:19-303 -- every image of the batch is supervised."""
    return _layer(rpn_cls_score, gt_boxes, num_gt_boxes, im_info, rpn_cls_score.shape[0], 0,
                  _feat_stride, anchor_scales, dataset, sampler, seed, return_device)


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
                              dataset='SNUBH', sampler="host", seed=0, return_device=False):
    """:328-628 -- the first IMS_PER_BATCH images are supervised, the WS_IMS_PER_BATCH
    weakly-supervised ones get all-don't-care labels when training."""
    n_ws = cfg.TRAIN.WS_IMS_PER_BATCH if is_training else 0
    return _layer(rpn_cls_score, gt_boxes, num_gt_boxes, im_info, cfg.TRAIN.IMS_PER_BATCH, n_ws,
                  _feat_stride, anchor_scales, dataset, sampler, seed, return_device)
