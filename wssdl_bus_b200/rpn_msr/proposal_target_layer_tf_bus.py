"""rpn_msr/proposal_target_layer_tf_bus.py twin: proposal_target_layer (:15-97),
proposal_target_layer_joint (:99-184), _sample_rois (:228-280), device resident.

Two launches for all supervised images (csrc/targets.cu): roi_match (the image's candidate RoIs
(+ GT rows when training), fp64 IoU, max / argmax, fg / bg candidacy) and roi_targets (gather of the
selected rows, labels, fp32 regression targets, 4-of-4K expansion, weights).  The selection
(:243-261) has two modes, like the anchor-target layer:
  sampler="host"    npr.choice(n_fg, fg_this, replace=False) then npr.choice(n_bg, bg_this,
                    replace=False) per image from numpy.random, in the reference's order -- the
                    draws depend on the population SIZES only, which the device reports in one
                    [Bs,4] int32 read (the only host trip); a seeded run consumes the same random
                    stream and reproduces the reference's outputs bit for bit;
  sampler="philox"  drawn on the device (include/wssdl_b200.h documents the stream); no host trip,
                    image i owns rows [i*BATCH_SIZE, (i+1)*BATCH_SIZE) and `counts` says how many
                    of them are fg / bg.
"""
import numpy as np
import numpy.random as npr
import torch

from wssdl_bus_b200 import ops
from wssdl_bus_b200.fast_rcnn.config import cfg


def _sample_rois_ws(all_rois, num_classes):
    """:282-295 -- weakly supervised image: every RoI, zero labels/targets."""
    labels = np.zeros((all_rois.shape[0],), dtype=np.float32)
    bbox_targets = np.zeros((all_rois.shape[0], 4 * num_classes), dtype=np.float32)
    return labels, all_rois, bbox_targets, np.zeros(bbox_targets.shape, dtype=np.float32)


def _rows_of(rpn_rois, i):
    return rpn_rois[rpn_rois[:, 0] == i, :].reshape(-1, 5)


def _layer(rpn_rois, gt_boxes, num_gt_boxes, _num_classes, n_supervised, ws_range, add_gt,
           all_ws=False, sampler="host", seed=0, return_device=False):
    K = int(_num_classes)
    as_np = not torch.is_tensor(rpn_rois)
    rois_per_image = int(cfg.TRAIN.BATCH_SIZE // 1)            # py2 integer division (:57)
    fg_rois_per_image = int(np.round(cfg.TRAIN.FG_FRACTION * rois_per_image))
    if all_ws:                                                 # :62-65: every RoI, un-sampled, zero targets
        r = rpn_rois.cpu().numpy() if torch.is_tensor(rpn_rois) else rpn_rois
        g = gt_boxes.cpu().numpy() if torch.is_tensor(gt_boxes) else gt_boxes
        rows = [_rows_of(r, i) for i in range(n_supervised)]
        rois = np.concatenate([np.zeros((0, 5), r.dtype)] + rows)
        z = np.zeros((rois.shape[0], 4 * K), np.float32)
        return rois, np.zeros((rois.shape[0], 1), g.dtype), z, z.copy(), z.copy()
    gt = ops._cuda(gt_boxes[:n_supervised], torch.float32)
    dev = gt.device
    rois_d = ops._cuda(rpn_rois, torch.float32, dev)
    ng = ops._cuda(num_gt_boxes[:n_supervised], torch.int32, dev)
    m = ops.roi_match(rois_d, gt, ng, add_gt, cfg.TRAIN.FG_THRESH, cfg.TRAIN.BG_THRESH_HI,
                      cfg.TRAIN.BG_THRESH_LO)
    norm = {}
    if cfg.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED:
        norm = dict(normalize_means=cfg.TRAIN.BBOX_NORMALIZE_MEANS,
                    normalize_stds=cfg.TRAIN.BBOX_NORMALIZE_STDS)
    counts_out = None
    if sampler == "host":
        counts = m.counts.cpu().numpy()
        sel, sel_off, row_off = [], [0], [0]
        for i in range(n_supervised):
            ncand, n_fg, n_bg, npos = (int(v) for v in counts[i])
            if npos == 0 or ncand == 0:                       # overlaps.argmax(axis=1) on [N,0]
                raise ValueError("attempt to get argmax of an empty sequence")
            fg_this = min(fg_rois_per_image, n_fg)            # :243
            if n_fg > 0:
                sel.append(npr.choice(n_fg, size=fg_this, replace=False))   # :250
            sel_off.append(sel_off[-1] + fg_this)
            bg_this = min(rois_per_image - fg_this, n_bg)     # :256-258
            if n_bg > 0:
                sel.append(npr.choice(n_bg, size=bg_this, replace=False))   # :261
            sel_off.append(sel_off[-1] + bg_this)
            row_off.append(row_off[-1] + fg_this + bg_this)
        sel = np.concatenate(sel).astype(np.int32) if sel else np.zeros((0,), np.int32)
        out = ops.roi_targets(m, K, cfg.TRAIN.BBOX_INSIDE_WEIGHTS, sel=sel, sel_off=sel_off,
                              row_off=row_off, **norm)
    elif sampler == "philox":
        out = ops.roi_targets(m, K, cfg.TRAIN.BBOX_INSIDE_WEIGHTS, seed=int(seed),
                              fg_rois_per_image=fg_rois_per_image, rois_per_image=rois_per_image,
                              **norm)
        counts_out = out[5]
    else:
        raise ValueError("sampler must be 'host' or 'philox'")
    rois, labels, t, iw, ow = out[:5]
    ws_range = tuple(ws_range)
    if ws_range:   # weakly supervised images: every proposal, un-sampled (:162-182)
        rois = torch.cat([rois] + [_rows_of(rois_d, i) for i in ws_range])
    if as_np and not return_device:
        rdt = rpn_rois.dtype
        ldt = gt_boxes.dtype if isinstance(gt_boxes, np.ndarray) else np.float32
        res = (rois.cpu().numpy().astype(rdt, copy=False), labels.cpu().numpy().astype(ldt, copy=False),
               t.cpu().numpy(), iw.cpu().numpy(), ow.cpu().numpy())
    else:
        res = (rois, labels, t, iw, ow)
    return res + (counts_out,) if sampler == "philox" else res


def proposal_target_layer(rpn_rois, gt_boxes, num_gt_boxes, _num_classes, is_training=True,
                          is_ws=False, sampler="host", seed=0, return_device=False):
    """:15-97 -- GT boxes join the candidates when training a supervised batch (:45-50); a
    weakly supervised training batch passes every RoI through un-sampled (:62-65)."""
    return _layer(rpn_rois, gt_boxes, num_gt_boxes, _num_classes, gt_boxes.shape[0], (),
                  bool(is_training) and not bool(is_ws), all_ws=bool(is_training) and bool(is_ws),
                  sampler=sampler, seed=seed, return_device=return_device)


def proposal_target_layer_joint(rpn_rois, gt_boxes, num_gt_boxes, _num_classes, is_training,
                                sampler="host", seed=0, return_device=False):
    """:99-184."""
    n = cfg.TRAIN.IMS_PER_BATCH
    ws = range(n, n + cfg.TRAIN.WS_IMS_PER_BATCH) if is_training else ()
    return _layer(rpn_rois, gt_boxes, num_gt_boxes, _num_classes, n, ws, bool(is_training),
                  sampler=sampler, seed=seed, return_device=return_device)
