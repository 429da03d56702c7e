#!/usr/bin/env python
"""Every kernel once on small shapes, for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python scripts/sanitize_targets.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wssdl_bus_b200 import _lib, ops, synthetic as syn  # noqa: E402
from wssdl_bus_b200.rpn_msr.generate_anchors import generate_anchors  # noqa: E402

B, H, W, C = 2, 38, 50, 32
feat = torch.from_numpy(syn.feature_map(0, B, H, W, C)).cuda()
rois = np.concatenate([syn.rois_for_pool(1, 90, B), syn.adversarial_rois(B, W, H)])
for kern in ("direct", "tiled", "band", "sorted"):
    _lib.set_tuning("roi_fwd_kernel", kern)
    for mode in ("cpu", "gpu"):
        top, arg = ops.roi_pool_forward(feat, rois, 7, 7, 1 / 16., bin_mode=mode)
# sorted-bins kernel, linear-index variant (C % 128 == 0) incl. bins taller than the band overlap
feat128 = torch.from_numpy(syn.feature_map(9, B, H, W, 128)).cuda()
tall = np.array([[0, 100, -900, 500, 1700], [1, 0, 0, 799, 599]], np.float32)
ops.roi_pool_forward(feat128, np.concatenate([rois, tall]), 7, 7, 1 / 16.)
# counting-sort pre-pass (R > 4096): histogram + scatter kernels, then tiled / band
big = syn.rois_for_pool(2, 4200, B)
for kern in ("tiled", "band", "sorted"):
    _lib.set_tuning("roi_fwd_kernel", kern)
    top2, arg2 = ops.roi_pool_forward(feat, big, 7, 7, 1 / 16.)
_lib.set_tuning("roi_fwd_kernel", "auto")
g = torch.randn_like(top)
for det in (False, True):
    ops.roi_pool_backward((B, H, W, C), rois, arg, g, 7, 7, 1 / 16., deterministic=det)
cls, reg, info = syn.rpn_outputs(3, B, H, W, 9)
for cl in (0, 2, 4, 8):     # one CTA per image / cluster of 2, 4, 8 CTAs per image
    _lib.set_tuning("proposals_cluster", cl)
    ops.proposals(cls, reg, info, generate_anchors(), 16, 6000, 300, 0.7, 16)
    ops.proposals(cls, reg, info, generate_anchors(), 16, 2000, 500, 0.7, 16, want_decoded=True)
_lib.set_tuning("proposals_cluster", -1)
d = syn.dets(4, 5000)
for cl in (0, 1):     # single-CTA sweep / cluster sweep
    _lib.set_tuning("nms_sweep_cluster", cl)
    ops.nms(d, 0.7)
    ops.nms(d, 0.3, mode=ops.NMS_GT_F32 | ops.NMS_CONTAIN)
_lib.set_tuning("nms_sweep_cluster", -1)
b = syn.random_boxes(5, 3000).astype(np.float64)
q = syn.random_boxes(6, 130).astype(np.float64)
ops.bbox_overlaps(b, q)
ops.bbox_overlaps_ui(b, q[:21])
ops.bbox_overlaps_device(b.astype(np.float32), q.astype(np.float32), ops.IOU, torch.float32)
gt, num = syn.gt_boxes(7, B)
ops.anchor_labels(gt, num, info, H, W, generate_anchors(), 16)
S, K = 300, 3
r2 = np.concatenate([syn.rois_for_pool(8 + i, S) for i in range(B)])
sc, dl = syn.rcnn_head_outputs(8, B * S, K)
meta = np.tile(np.array([[437, 583, 600.0 / 437]], np.float32), (B, 1))
ops.detect_postprocess(r2, sc, dl, meta, roi_stride=S)
ops.detect_postprocess(r2, sc, dl, meta, roi_stride=S, cls_agnostic=True, max_per_image=40)
ops.bbox_transform_inv(b[:, :4].astype(np.float32), np.zeros((3000, 12), np.float32))
# fused hot-path entry (PDL chain, grouped bin sort with sub-lists, -1 padding rows) and the
# grouped RoI-pool entry, one and several images
from wssdl_bus_b200.pipeline import HotPath  # noqa: E402
from wssdl_bus_b200.rpn_msr import anchor_target_layer_tf_bus as atl  # noqa: E402
from wssdl_bus_b200.rpn_msr import proposal_target_layer_tf_bus as ptl  # noqa: E402
for nb in (1, 5):
    f512 = torch.from_numpy(syn.feature_map(11, nb, H, W, 64)).cuda()
    c_, r_, i_ = [torch.from_numpy(v).cuda() for v in syn.rpn_outputs(12, nb, H, W, 9)]
    hot = HotPath(pre_nms_topN=200)
    for kern in ("auto", "sorted"):
        _lib.set_tuning("roi_fwd_kernel", kern)
        out = hot.run(f512, c_, r_, i_)
        ops.roi_pool_forward_grouped(f512, out["rois"], 300, 7, 7, 1 / 16.)
_lib.set_tuning("roi_fwd_kernel", "auto")
# device-resident target layers, both samplers
np.random.seed(1)
score = np.zeros((B, H, W, 18), np.float32)
info4 = np.tile(np.array([[600, 800, 1.0]], np.float32), (B, 1))
for sampler in ("host", "philox"):
    atl.anchor_target_layer(score, gt, num, info4, None, [16, ], [8, 16, 32], "SNUBH", sampler=sampler, seed=3)
    atl.anchor_target_layer(score, gt, num, info4, None, [16, ], [8, 16, 32], "VOC", sampler=sampler, seed=3)
    pr = np.concatenate([np.hstack((np.full((260, 1), i, np.float32), syn.rois_for_pool(20 + i, 260)[:, 1:]))
                         for i in range(B)])
    ptl.proposal_target_layer(pr, gt, num, 3, True, False, sampler=sampler, seed=4)
torch.cuda.synchronize()
print("sanitize targets done")
