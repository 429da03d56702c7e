"""fast_rcnn/test_bus.py twin for the detection post-processing that follows the RCNN head.

Only the hot-path pieces of the reference module are mirrored (the session loop, blobs,
plotting and pickling are out of scope):

  detect_boxes(...)            tail of im_detect (test_bus.py:207-223): RoIs back to the
                               unscaled frame, per-class box regression, _clip_boxes
  postprocess_detections(...)  per-image body of test_net (test_bus.py:360-401): score
                               threshold, per-class nms(cfg.TEST.NMS), optional
                               class-agnostic NMS, cap at max_per_image
  test_net_batch(...)          both, batched over images on the device; returns the
                               reference's all_boxes[cls][image] nested lists

Everything runs in ONE kernel launch per batch (csrc/detect.cu); nothing here computes on
the CPU.
"""
import numpy as np
import torch

from wssdl_bus_b200 import ops
from wssdl_bus_b200.fast_rcnn.config import cfg


def _meta(im_shapes, im_scales):
    im_shapes = np.asarray(im_shapes, dtype=np.float32).reshape(-1, np.shape(im_shapes)[-1])[:, :2]
    im_scales = np.asarray(im_scales, dtype=np.float32).reshape(-1)
    return np.concatenate([im_shapes, im_scales[:, None]], axis=1)


def detect_boxes(rois, bbox_pred, im_shape, im_scale):
    """im_detect tail for ONE image: -> pred_boxes [R,4K] (numpy in, numpy out)."""
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    R = rois.shape[0]
    K = bbox_pred.shape[1] // 4
    if R == 0:
        return np.zeros((0, 4 * K), np.float32)
    out = ops.detect_postprocess(rois, np.zeros((R, K), np.float32), bbox_pred,
                                 _meta([im_shape], [im_scale]), roi_stride=R,
                                 max_per_image=0, want_pred_boxes=True)
    return out["pred_boxes"].cpu().numpy()


def test_net_batch(rois, scores, bbox_pred, im_shapes, im_scales, roi_counts=None,
                   roi_stride=None, max_per_image=300, thresh=0.05):
    """Batched test_net body.  Returns (all_boxes, out): all_boxes[cls][image] = [n,5] numpy
    arrays like the reference's nested lists (test_bus.py:306-307), out = the device blob
    (dets [B,K,S,5], counts [B,K]) for the all-gather."""
    out = ops.detect_postprocess(rois, scores, bbox_pred, _meta(im_shapes, im_scales),
                                 roi_counts=roi_counts, roi_stride=roi_stride,
                                 score_thresh=thresh, nms_thresh=cfg.TEST.NMS,
                                 max_per_image=max_per_image,
                                 cls_agnostic=cfg.TEST.CLS_AGNOSTIC_NMS)
    dets = out["dets"].cpu().numpy()
    counts = out["counts"].cpu().numpy()
    if int(out["status"].item()):
        raise ZeroDivisionError("float division")      # utils/nms.pyx raises on a zero union
    B, K = counts.shape
    all_boxes = [[dets[i, j, :counts[i, j]].copy() for i in range(B)] for j in range(K)]
    return all_boxes, out


test_net_batch.__test__ = False        # not a pytest test despite the reference's name


def postprocess_detections(scores, boxes, max_per_image=300, thresh=0.05):
    """test_net body for ONE image given already regressed boxes [R,4K] (numpy in/out):
    list over classes of [n_j,5] arrays.  The boxes pass through the device path unchanged
    (zero deltas on a unit scale would re-clip them, so the NMS entry point is used)."""
    scores = np.asarray(scores, dtype=np.float32)
    boxes = np.asarray(boxes, dtype=np.float32)
    K = scores.shape[1]
    out = [np.zeros((0, 5), np.float32) for _ in range(K)]
    for j in range(1, K):
        inds = np.where(scores[:, j] > np.float32(thresh))[0]
        cls_dets = np.hstack((boxes[inds, j * 4:(j + 1) * 4], scores[inds, j][:, None]))
        cls_dets = np.ascontiguousarray(cls_dets, dtype=np.float32)
        keep = ops.nms(cls_dets, cfg.TEST.NMS) if len(inds) else []
        out[j] = cls_dets[keep, :]
    if cfg.TEST.CLS_AGNOSTIC_NMS:
        all_dets = np.concatenate([np.hstack((out[j], np.full((out[j].shape[0], 1), j, np.float32)))
                                   for j in range(1, K)], axis=0)
        keep = ops.nms(np.ascontiguousarray(all_dets), cfg.TEST.NMS) if len(all_dets) else []
        all_dets = all_dets[keep, :]
        for j in range(1, K):
            out[j] = all_dets[all_dets[:, 5] == j, :5]
    if max_per_image > 0:
        image_scores = np.hstack([out[j][:, -1] for j in range(1, K)])
        if len(image_scores) > max_per_image:
            image_thresh = np.sort(image_scores)[-max_per_image]
            for j in range(1, K):
                out[j] = out[j][out[j][:, -1] >= image_thresh, :]
    return out
