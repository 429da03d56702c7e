"""rpn_msr/proposal_layer_tf_bus.py:19-148 twin: the whole layer is ONE fused kernel launch
for the batch (csrc/proposal.cu) instead of a per-image numpy pipeline ending in cpu_nms."""
import numpy as np
import torch

from wssdl_bus_b200 import ops
from wssdl_bus_b200.fast_rcnn.config import cfg
from wssdl_bus_b200.rpn_msr.generate_anchors import generate_anchors


def _stride(feat_stride):
    return int(feat_stride[0]) if isinstance(feat_stride, (list, tuple, np.ndarray)) else int(feat_stride)


def _nms_mode():
    """nms_wrapper.nms (fast_rcnn/nms_wrapper.py:13-21): gpu_nms ('>' in fp32) with
    cfg.USE_GPU_NMS, cpu_nms ('>=' against the double threshold) without."""
    return ops.NMS_GT_F32 if cfg.USE_GPU_NMS else ops.NMS_GE_F64


def proposal_layer(rpn_cls_prob_reshape, rpn_bbox_pred, im_info, is_training, is_ws=False,
                   _feat_stride=[16, ], anchor_scales=[8, 16, 32], return_device=False):
    """Inputs NHWC as the reference's py_func receives them: [B,H,W,2A], [B,H,W,4A], im_info
    [B,3|4].  Returns the (sum R, 5) float32 blob (batch_idx, x1, y1, x2, y2): numpy for numpy
    inputs, a CUDA tensor otherwise (or always with return_device=True).  `is_ws` is accepted
    and ignored, as in the reference."""
    del is_ws
    cfg_key = 'TRAIN' if is_training else 'TEST'
    base = generate_anchors(scales=np.array(anchor_scales))
    as_np = isinstance(rpn_cls_prob_reshape, np.ndarray)
    out = ops.proposals(rpn_cls_prob_reshape, rpn_bbox_pred, im_info, base, _stride(_feat_stride),
                        cfg[cfg_key].RPN_PRE_NMS_TOP_N, cfg[cfg_key].RPN_POST_NMS_TOP_N,
                        cfg[cfg_key].RPN_NMS_THRESH, cfg[cfg_key].RPN_MIN_SIZE, nms_mode=_nms_mode())
    blob = ops.compact_rois(out)
    if as_np and not return_device:
        return blob.cpu().numpy()
    return blob


def proposal_layer_batched(rpn_cls_prob_reshape, rpn_bbox_pred, im_info, is_training,
                           _feat_stride=[16, ], anchor_scales=[8, 16, 32], want_decoded=False):
    """Same computation, fixed-stride device outputs (no host sync): dict with rois
    [B*post,5], scores, anchor_idx, counts -- the form the multi-GPU pipeline consumes."""
    cfg_key = 'TRAIN' if is_training else 'TEST'
    base = generate_anchors(scales=np.array(anchor_scales))
    return ops.proposals(rpn_cls_prob_reshape, rpn_bbox_pred, im_info, base, _stride(_feat_stride),
                         cfg[cfg_key].RPN_PRE_NMS_TOP_N, cfg[cfg_key].RPN_POST_NMS_TOP_N,
                         cfg[cfg_key].RPN_NMS_THRESH, cfg[cfg_key].RPN_MIN_SIZE,
                         want_decoded=want_decoded, nms_mode=_nms_mode())


def _filter_boxes(boxes, min_size):
    """proposal_layer_tf_bus.py:151-156 (kept for callers that import it; device tensors)."""
    boxes = ops._cuda(boxes, torch.float32)
    ws = boxes[:, 2] - boxes[:, 0] + 1
    hs = boxes[:, 3] - boxes[:, 1] + 1
    return torch.nonzero((ws >= min_size) & (hs >= min_size)).reshape(-1)
