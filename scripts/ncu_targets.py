#!/usr/bin/env python
"""Small driver for ncu captures of the secondary kernels (one shape each, few launches):
   python scripts/ncu_targets.py nms|iou|detect|bwd|roi4[:kernel[:slices[:chunks[:threads]]]]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wssdl_bus_b200 import ops, synthetic as syn  # noqa: E402

what = sys.argv[1]
if what == "nms":
    d = torch.from_numpy(syn.dets(7, 20000)).cuda()
    for _ in range(3):
        ops.nms_device(d, 0.7)
elif what == "iou":
    b = torch.from_numpy(syn.random_boxes(3, 20000).astype(np.float64)).cuda()
    for dt in (torch.float64, torch.float32):
        for _ in range(3):
            ops.bbox_overlaps_device(b.to(dt), b.to(dt), ops.IOU, dt)
elif what == "detect":
    B, S, K = 256, 300, 3
    rois = np.concatenate([syn.rois_for_pool(9 + b, S) for b in range(B)])
    scores, deltas = syn.rcnn_head_outputs(9, B * S, K)
    meta = np.tile(np.array([[437, 583, 600.0 / 437]], np.float32), (B, 1))
    args = [torch.from_numpy(v).cuda() for v in (rois, scores, deltas, meta)]
    for _ in range(3):
        ops.detect_postprocess(*args, roi_stride=S)
elif what == "bwd":
    B, H, W, C = 16, 38, 50, 1024
    x = torch.from_numpy(syn.feature_map(1, B, H, W, C)).cuda()
    r = torch.from_numpy(syn.rois_for_pool(5, B * 300, B)).cuda()
    top, arg = ops.roi_pool_forward(x, r, 14, 14, 1 / 16.)
    g = torch.randn_like(top)
    for _ in range(3):
        ops.roi_pool_backward((B, H, W, C), r, arg, g, 14, 14, 1 / 16.)
elif what.startswith("prop"):
    # the proposals kernel on the C4 shapes: prop:<images>[:heavy]
    from wssdl_bus_b200.pipeline import HotPath
    parts = what.split(":")
    nimg = int(parts[1]) if len(parts) > 1 else 32
    cls, reg, info = syn.rpn_outputs(0, nimg, 38, 50, 9)
    if len(parts) > 2 and parts[2] == "heavy":
        reg = (reg * 0.1).astype(np.float32)
    cls, reg, info = [torch.from_numpy(v).cuda() for v in (cls, reg, info)]
    hot = HotPath()
    for _ in range(3):
        ops.proposals(cls, reg, info, hot.base, 16, hot.pre, hot.post, hot.thresh, hot.min_size)
elif what.startswith("roi4"):
    # the C4 RoI-pool forward launch (256 images x 300 proposal RoIs), 3 launches
    from wssdl_bus_b200 import _lib
    from wssdl_bus_b200.pipeline import HotPath
    parts = what.split(":")
    nimg = int(os.environ.get("NCU_IMAGES", "256"))
    cls, reg, info = syn.rpn_outputs(0, nimg, 38, 50, 9)
    hot = HotPath()
    rois = ops.proposals(cls, reg, info, hot.base, 16, hot.pre, hot.post, hot.thresh, hot.min_size)["rois"]
    x = torch.from_numpy(syn.feature_map(1, nimg, 38, 50, 512)).cuda()
    if len(parts) > 1:
        _lib.set_tuning("roi_fwd_kernel", parts[1])
    if len(parts) > 2:
        _lib.set_tuning("roi_fwd_slices", int(parts[2]))
    if len(parts) > 3:
        _lib.set_tuning("roi_fwd_chunks", int(parts[3]))
    if len(parts) > 4:
        _lib.set_tuning("roi_fwd_threads", int(parts[4]))
    for _ in range(3):
        if os.environ.get("NCU_GROUPED"):
            ops.roi_pool_forward_grouped(x, rois, 300, 7, 7, 1 / 16.)
        else:
            ops.roi_pool_forward(x, rois, 7, 7, 1 / 16.)
torch.cuda.synchronize()
