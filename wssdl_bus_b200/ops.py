"""Tensor-level entry points: torch tensors carry the buffers, libwssdl_b200.so does the work.

Every function enqueues on ``torch.cuda.current_stream()`` and returns device tensors
without synchronising (unless documented otherwise).  numpy inputs are copied to the
current CUDA device first.  Nothing here computes on the CPU: without CUDA or without the
compiled library these functions raise.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import (BIN_CPU_TRUNC, BIN_GPU_CEIL, BWD_ATOMIC, BWD_GATHER, IOU, IOU_UI,  # noqa: F401
                   NMS_CONTAIN, NMS_GE_F64, NMS_GT_F32, WssdlError)

_vp = ctypes.c_void_p


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("wssdl_bus_b200 needs a CUDA device (B200, sm_100a); there is no "
                           "CPU fallback")


def _cuda(x, dtype, device=None):
    """numpy / torch (any device) -> contiguous CUDA tensor of `dtype`."""
    _require_cuda()
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    elif not torch.is_tensor(x):
        x = torch.as_tensor(np.asarray(x))
    if device is None:
        device = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()


def _ptr(t):
    return _vp(t.data_ptr()) if (t is not None and t.numel() > 0) else _vp(None)


def _stream(device):
    return _vp(torch.cuda.current_stream(device).cuda_stream)


_workspaces = {}


def _workspace(nbytes, device):
    """Grow-only scratch per (device, stream); the library never allocates on its own."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _bin_mode(bin_mode):
    if bin_mode in ("cpu", "cpu_trunc", BIN_CPU_TRUNC):
        return BIN_CPU_TRUNC
    if bin_mode in ("gpu", "gpu_ceil", BIN_GPU_CEIL):
        return BIN_GPU_CEIL
    raise ValueError("bin_mode must be 'cpu' (roi_pooling_op.cc) or 'gpu' (roi_pooling_op_gpu.cu.cc)")


# ------------------------------------------------------------------ RoI pooling
def roi_pool_forward(bottom, rois, pooled_height, pooled_width, spatial_scale, bin_mode="cpu",
                     need_argmax=True):
    """RoiPool forward.  bottom [B,H,W,C] f32 NHWC, rois [R,5] -> (top, argmax) [R,PH,PW,C].

    Mirrors the op's checks (roi_pooling_op.cc:73-82, :97-102)."""
    if pooled_height < 0:
        raise ValueError("Need pooled_height >= 0, got %d" % pooled_height)
    if pooled_width < 0:
        raise ValueError("Need pooled_width >= 0, got %d" % pooled_width)
    bottom = _cuda(bottom, torch.float32)
    if bottom.dim() != 4:
        raise ValueError("data must be 4-dimensional")
    rois = _cuda(rois, torch.float32, bottom.device)
    if rois.dim() != 2:
        raise ValueError("rois must be 2-dimensional")
    if rois.shape[1] != 5:
        raise ValueError("rois must be [R,5] rows (batch, x1, y1, x2, y2)")
    B, H, W, C = bottom.shape
    R = rois.shape[0]
    with torch.cuda.device(bottom.device):
        top = torch.empty((R, pooled_height, pooled_width, C), dtype=torch.float32,
                          device=bottom.device)
        argmax = torch.empty_like(top, dtype=torch.int32) if need_argmax else None
        ws = _workspace(_lib.lib().wssdl_roi_pool_fwd_workspace_bytes(B, R, pooled_height, pooled_width), bottom.device)
        rc = _lib.lib().wssdl_roi_pool_fwd(_ptr(bottom), _ptr(rois), B, H, W, C, R, pooled_height,
                                           pooled_width, float(spatial_scale), _bin_mode(bin_mode),
                                           _ptr(top), _ptr(argmax), _vp(ws.data_ptr()), ws.numel(),
                                           _stream(bottom.device))
    _lib.check(rc, "wssdl_roi_pool_fwd")
    return top, argmax


def roi_pool_forward_grouped(bottom, rois, roi_stride, pooled_height, pooled_width, spatial_scale,
                             bin_mode="cpu", need_argmax=True):
    """roi_pool_forward for image-major RoIs (wssdl_roi_pool_fwd_grouped): row r of rois
    [B*roi_stride,5] belongs to image r // roi_stride -- the blob proposals() writes; rows whose
    batch index says otherwise pool to zeros / -1.  Skips the RoI grouping pre-pass."""
    bottom = _cuda(bottom, torch.float32)
    rois = _cuda(rois, torch.float32, bottom.device)
    if bottom.dim() != 4 or rois.dim() != 2 or rois.shape[1] != 5:
        raise ValueError("bottom [B,H,W,C], rois [B*roi_stride,5]")
    B, H, W, C = bottom.shape
    if rois.shape[0] != B * int(roi_stride):
        raise ValueError("rois must hold roi_stride rows per image")
    R = rois.shape[0]
    with torch.cuda.device(bottom.device):
        top = torch.empty((R, pooled_height, pooled_width, C), dtype=torch.float32,
                          device=bottom.device)
        argmax = torch.empty_like(top, dtype=torch.int32) if need_argmax else None
        ws = _workspace(_lib.lib().wssdl_roi_pool_fwd_workspace_bytes(B, R, pooled_height, pooled_width),
                        bottom.device)
        rc = _lib.lib().wssdl_roi_pool_fwd_grouped(
            _ptr(bottom), _ptr(rois), int(roi_stride), B, H, W, C, pooled_height, pooled_width,
            float(spatial_scale), _bin_mode(bin_mode), _ptr(top), _ptr(argmax), _vp(ws.data_ptr()),
            ws.numel(), _stream(bottom.device))
    _lib.check(rc, "wssdl_roi_pool_fwd_grouped")
    return top, argmax


def roi_pool_backward(bottom_shape, rois, argmax, grad, pooled_height, pooled_width,
                      spatial_scale, deterministic=False):
    """RoiPoolGrad.  Returns bottom_diff [B,H,W,C] f32."""
    grad = _cuda(grad, torch.float32)
    argmax = _cuda(argmax, torch.int32, grad.device)
    rois = _cuda(rois, torch.float32, grad.device)
    if grad.dim() != 4:
        raise ValueError("out_backprop must be 4-dimensional")
    if argmax.dim() != 4:
        raise ValueError("argmax_data must be 4-dimensional")
    if rois.dim() != 2:
        raise ValueError("rois must be 2-dimensional")
    B, H, W, C = [int(v) for v in bottom_shape]
    R = rois.shape[0]
    if tuple(grad.shape) != (R, pooled_height, pooled_width, C) or grad.shape != argmax.shape:
        raise ValueError("grad/argmax must be [R,PH,PW,C]")
    with torch.cuda.device(grad.device):
        out = torch.empty((B, H, W, C), dtype=torch.float32, device=grad.device)
        rc = _lib.lib().wssdl_roi_pool_bwd(_ptr(grad), _ptr(argmax), _ptr(rois), B, H, W, C, R,
                                           pooled_height, pooled_width, float(spatial_scale),
                                           BWD_GATHER if deterministic else BWD_ATOMIC, _ptr(out),
                                           _stream(grad.device))
    _lib.check(rc, "wssdl_roi_pool_bwd")
    return out


class _RoiPoolFn(torch.autograd.Function):
    """Gradient registration twin of roi_pooling_op_grad.py:24-44."""

    @staticmethod
    def forward(ctx, bottom, rois, ph, pw, scale, bin_mode, deterministic):
        top, argmax = roi_pool_forward(bottom, rois, ph, pw, scale, bin_mode)
        ctx.save_for_backward(rois, argmax)
        ctx.meta = (tuple(bottom.shape), ph, pw, scale, deterministic)
        ctx.mark_non_differentiable(argmax)
        return top, argmax

    @staticmethod
    def backward(ctx, grad_top, _grad_argmax):
        rois, argmax = ctx.saved_tensors
        shape, ph, pw, scale, det = ctx.meta
        g = roi_pool_backward(shape, rois, argmax, grad_top.contiguous(), ph, pw, scale, det)
        return g, None, None, None, None, None, None


def roi_pool(bottom_data, bottom_rois, pooled_height, pooled_width, spatial_scale, name=None,
             bin_mode="cpu", deterministic_grad=False):
    """Drop-in for roi_pooling_op.roi_pool (roi_pooling_op.py:6): -> (top_data, argmax)."""
    del name
    if torch.is_tensor(bottom_data) and bottom_data.requires_grad and torch.is_grad_enabled():
        return _RoiPoolFn.apply(bottom_data, _cuda(bottom_rois, torch.float32, bottom_data.device),
                                pooled_height, pooled_width, spatial_scale, bin_mode,
                                deterministic_grad)
    return roi_pool_forward(bottom_data, bottom_rois, pooled_height, pooled_width, spatial_scale,
                            bin_mode)


def roi_pool_grad(data, rois, argmax, grad, pooled_height, pooled_width, spatial_scale,
                  deterministic=False):
    """Drop-in for roi_pooling_op.roi_pool_grad (roi_pooling_op.py:7); `data` supplies the
    bottom shape only, as in RoiPoolGradOp (roi_pooling_op.cc:372)."""
    return roi_pool_backward(tuple(data.shape), rois, argmax, grad, pooled_height, pooled_width,
                             spatial_scale, deterministic)


# ------------------------------------------------------------------ NMS
def nms_device(dets, thresh, mode=NMS_GE_F64, max_keep=0):
    """dets: CUDA/numpy [N,>=5] f32.  Returns (keep i32 [min(N,max_keep or N)], num i32 [1],
    status i32 [2]) on the device, no synchronisation."""
    dets = _cuda(dets, torch.float32)
    if dets.dim() != 2 or dets.shape[1] < 5:
        raise ValueError("dets must be [N,5] rows (x1,y1,x2,y2,score)")
    N = dets.shape[0]
    dev = dets.device
    with torch.cuda.device(dev):
        cap = min(N, max_keep) if max_keep > 0 else N
        keep = torch.empty((max(cap, 1),), dtype=torch.int32, device=dev)
        misc = torch.zeros((4,), dtype=torch.int32, device=dev)
        nbytes = _lib.lib().wssdl_nms_workspace_bytes(N)
        ws = _workspace(nbytes, dev)
        rc = _lib.lib().wssdl_nms(_ptr(dets), N, dets.shape[1], float(thresh), int(mode),
                                  int(max_keep), _vp(keep.data_ptr()), _vp(misc.data_ptr()),
                                  _vp(misc.data_ptr() + 4), _vp(ws.data_ptr()), ws.numel(),
                                  _stream(dev))
    _lib.check(rc, "wssdl_nms")
    return keep[:cap], misc[0:1], misc[1:3]


def nms(dets, thresh, mode=NMS_GE_F64, max_keep=0):
    """cpu_nms-shaped call: returns a Python list of kept indices in descending-score
    order.  numpy input goes through the host entry point of the C ABI (one H2D, one D2H);
    CUDA tensors stay on the device until the final read."""
    if isinstance(dets, np.ndarray) or (torch.is_tensor(dets) and not dets.is_cuda):
        _require_cuda()
        d = np.ascontiguousarray(dets.numpy() if torch.is_tensor(dets) else dets, dtype=np.float32)
        if d.ndim != 2 or d.shape[1] < 5:
            raise ValueError("dets must be [N,5] rows (x1,y1,x2,y2,score)")
        n = d.shape[0]
        if n == 0:
            return []
        keep = np.empty((n,), dtype=np.int32)
        num = ctypes.c_int(0)
        rc = _lib.lib().wssdl_nms_host(keep.ctypes.data_as(_vp), ctypes.byref(num),
                                       d.ctypes.data_as(_vp), n, d.shape[1], float(thresh),
                                       int(mode), int(max_keep), torch.cuda.current_device())
        if rc == _lib.EZERODIV:
            raise ZeroDivisionError("float division")
        _lib.check(rc, "wssdl_nms_host")
        return keep[:num.value].tolist()
    keep, num, status = nms_device(dets, thresh, mode, max_keep)
    if dets.shape[0] == 0:
        return []
    n, zero = int(num.item()), int(status[0].item())
    if zero:
        raise ZeroDivisionError("float division")
    return keep[:n].tolist()


def gpu_nms_sorted_host(sorted_dets, thresh, device_id=0):
    """`_nms` twin (nms_kernel.cu:91): host array already sorted, '>' in fp32."""
    _require_cuda()
    d = np.ascontiguousarray(sorted_dets, dtype=np.float32)
    n = d.shape[0]
    if n == 0:
        return np.zeros((0,), np.int32)
    keep = np.empty((n,), dtype=np.int32)
    num = ctypes.c_int(0)
    rc = _lib.lib().wssdl_gpu_nms_host(keep.ctypes.data_as(_vp), ctypes.byref(num),
                                       d.ctypes.data_as(_vp), n, d.shape[1], float(thresh),
                                       int(device_id))
    if rc == _lib.EZERODIV:
        rc = _lib.OK        # the reference's CUDA path never raises
    _lib.check(rc, "wssdl_gpu_nms_host")
    return keep[:num.value]


# ------------------------------------------------------------------ IoU
def bbox_overlaps_device(boxes, query, kind=IOU, dtype=torch.float64):
    boxes = _cuda(boxes, dtype)
    query = _cuda(query, dtype, boxes.device)
    if boxes.dim() != 2 or query.dim() != 2 or boxes.shape[1] < 4 or query.shape[1] < 4:
        raise ValueError("boxes and query_boxes must be [N,4] / [K,4]")
    if boxes.shape[1] != 4:
        boxes = boxes[:, :4].contiguous()
    if query.shape[1] != 4:
        query = query[:, :4].contiguous()
    N, K = boxes.shape[0], query.shape[0]
    with torch.cuda.device(boxes.device):
        out = torch.empty((N, K), dtype=dtype, device=boxes.device)
        fn = (_lib.lib().wssdl_bbox_overlaps_f64 if dtype == torch.float64
              else _lib.lib().wssdl_bbox_overlaps_f32)
        rc = fn(_ptr(boxes), N, _ptr(query), K, kind, _ptr(out), _stream(boxes.device))
    _lib.check(rc, "wssdl_bbox_overlaps")
    return out


def _np_out(fn):
    def wrapped(boxes, query, *a, **k):
        as_np = isinstance(boxes, np.ndarray)
        out = fn(boxes, query, *a, **k)
        return out.cpu().numpy() if as_np else out
    return wrapped


@_np_out
def bbox_overlaps(boxes, query_boxes):
    """utils.cython_bbox.bbox_overlaps: fp64 IoU matrix (numpy in -> numpy out)."""
    return bbox_overlaps_device(boxes, query_boxes, IOU, torch.float64)


@_np_out
def bbox_overlaps_ui(boxes, query_boxes):
    """utils.cython_bbox_ui.bbox_overlaps_ui: intersection / area(boxes[n]), fp64."""
    return bbox_overlaps_device(boxes, query_boxes, IOU_UI, torch.float64)


# ------------------------------------------------------------------ box transforms
def bbox_transform_inv(boxes, deltas):
    as_np = isinstance(deltas, np.ndarray)
    deltas_t = _cuda(deltas, torch.float32)
    boxes_t = _cuda(boxes, torch.float32, deltas_t.device)
    if boxes_t.shape[0] == 0:
        out = torch.zeros((0, deltas_t.shape[1]), dtype=torch.float32, device=deltas_t.device)
        return out.cpu().numpy() if as_np else out
    N, k4 = deltas_t.shape
    if k4 % 4 != 0 or boxes_t.shape != (N, 4):
        raise ValueError("boxes [N,4], deltas [N,4k]")
    with torch.cuda.device(deltas_t.device):
        out = torch.empty_like(deltas_t)
        rc = _lib.lib().wssdl_bbox_transform_inv(_ptr(boxes_t), _ptr(deltas_t), N, k4 // 4,
                                                 _ptr(out), _stream(deltas_t.device))
    _lib.check(rc, "wssdl_bbox_transform_inv")
    return out.cpu().numpy() if as_np else out


def clip_boxes(boxes, im_shape):
    """In place for CUDA tensors; numpy input is clipped on the device and copied back into
    the same array (the reference mutates its argument, bbox_transform.py:69-75)."""
    as_np = isinstance(boxes, np.ndarray)
    t = _cuda(boxes, torch.float32)
    N, k4 = t.shape
    with torch.cuda.device(t.device):
        rc = _lib.lib().wssdl_clip_boxes(_ptr(t), N, k4 // 4, float(im_shape[0]),
                                         float(im_shape[1]), _stream(t.device))
    _lib.check(rc, "wssdl_clip_boxes")
    if as_np:
        boxes[...] = t.cpu().numpy()
        return boxes
    if torch.is_tensor(boxes) and boxes.data_ptr() != t.data_ptr():
        boxes.copy_(t)
        return boxes
    return t


def bbox_transform(ex_rois, gt_rois):
    as_np = isinstance(ex_rois, np.ndarray)
    ex = _cuda(ex_rois, torch.float32)
    gt = _cuda(gt_rois, torch.float32, ex.device)
    ex = ex[:, :4].contiguous()
    gt = gt[:, :4].contiguous()
    N = ex.shape[0]
    with torch.cuda.device(ex.device):
        out = torch.empty((N, 4), dtype=torch.float32, device=ex.device)
        rc = _lib.lib().wssdl_bbox_transform(_ptr(ex), _ptr(gt), N, _ptr(out), _stream(ex.device))
    _lib.check(rc, "wssdl_bbox_transform")
    return out.cpu().numpy() if as_np else out


# ------------------------------------------------------------------ proposals
def proposals(cls_prob, bbox_pred, im_info, base_anchors, feat_stride, pre_nms_topN,
              post_nms_topN, nms_thresh, min_size, want_decoded=False, out=None,
              nms_mode=NMS_GE_F64, pad_rows_invalid=False):
    """Batched fused proposal layer.  cls_prob [B,H,W,2A], bbox_pred [B,H,W,4A] (NHWC),
    im_info [B,>=3].  Returns dict of device tensors: rois [B*post,5], scores [B*post],
    anchor_idx [B*post] i32, counts [B] i32 (+ decoded [B,H*W*A,4] when asked).
    out: optional (rois, scores, counts) tensors to write into (e.g. the views of one
    contiguous detection blob, pipeline.DetectionBlob, so that one all-gather moves them).
    nms_mode: NMS_GE_F64 (cpu_nms, cfg.USE_GPU_NMS False) or NMS_GT_F32 (gpu_nms).
    pre_nms_topN <= 0 / post_nms_topN <= 0: no truncation (proposal_layer_tf_bus.py:130, :139);
    the stride of the blob is then min(pre_nms_topN or H*W*A, H*W*A) (<= 4096 on the device).
    pad_rows_invalid: the hot path's blob convention (wssdl_hot_path_proposals): the unused rows of
    an image's block carry batch index -1 instead of 0 and pool to zeros / -1 (no anchor_idx)."""
    cls_prob = _cuda(cls_prob, torch.float32)
    dev = cls_prob.device
    bbox_pred = _cuda(bbox_pred, torch.float32, dev)
    im_info = _cuda(im_info, torch.float32, dev)
    if im_info.dim() == 1:
        im_info = im_info.reshape(1, -1)
    base = np.ascontiguousarray(base_anchors, dtype=np.float32)
    A = base.shape[0]
    B, H, W, C2 = cls_prob.shape
    if C2 != 2 * A or tuple(bbox_pred.shape) != (B, H, W, 4 * A) or im_info.shape[0] != B:
        raise ValueError("shape mismatch: cls_prob [B,H,W,2A], bbox_pred [B,H,W,4A], im_info [B,3+]")
    NA = H * W * A
    post = int(post_nms_topN)
    if post <= 0:
        post = min(int(pre_nms_topN), NA) if int(pre_nms_topN) > 0 else NA
    with torch.cuda.device(dev):
        if out is not None:
            rois, scores, counts = out
            if (tuple(rois.shape) != (B * post, 5) or tuple(scores.shape) != (B * post,) or
                    tuple(counts.shape) != (B,) or rois.dtype != torch.float32 or
                    scores.dtype != torch.float32 or counts.dtype != torch.int32 or
                    not (rois.is_contiguous() and scores.is_contiguous() and counts.is_contiguous())
                    or rois.device != dev):
                raise ValueError("out = (rois [B*post,5] f32, scores [B*post] f32, counts [B] i32), "
                                 "contiguous, on the inputs' device")
        else:
            rois = torch.empty((B * post, 5), dtype=torch.float32, device=dev)
            scores = torch.empty((B * post,), dtype=torch.float32, device=dev)
            counts = torch.empty((B,), dtype=torch.int32, device=dev)
        aidx = None if pad_rows_invalid else torch.empty((B * post,), dtype=torch.int32, device=dev)
        decoded = (torch.empty((B, H * W * A, 4), dtype=torch.float32, device=dev)
                   if want_decoded else None)
        ws = _workspace(_lib.lib().wssdl_proposals_workspace_bytes(B, H, W, A, int(pre_nms_topN),
                                                                   post), dev)
        if pad_rows_invalid:
            if want_decoded:
                raise ValueError("pad_rows_invalid: no decoded boxes through the hot-path stage entry")
            rc = _lib.lib().wssdl_hot_path_proposals(
                _ptr(cls_prob), _ptr(bbox_pred), _ptr(im_info), im_info.shape[1], B, H, W, A,
                base.ctypes.data_as(_vp), int(feat_stride), int(pre_nms_topN), post,
                float(nms_thresh), int(nms_mode), float(min_size), _ptr(rois), _ptr(scores),
                _ptr(counts), _stream(dev))
            _lib.check(rc, "wssdl_hot_path_proposals")
            return dict(rois=rois, scores=scores, counts=counts, post_nms_topN=post)
        rc = _lib.lib().wssdl_proposals(
            _ptr(cls_prob), _ptr(bbox_pred), _ptr(im_info), im_info.shape[1], B, H, W, A,
            base.ctypes.data_as(_vp), int(feat_stride), int(pre_nms_topN), post,
            float(nms_thresh), int(nms_mode), float(min_size), _ptr(rois), _ptr(scores), _ptr(aidx),
            _ptr(counts), _ptr(decoded), _vp(ws.data_ptr()), ws.numel(), _stream(dev))
    _lib.check(rc, "wssdl_proposals")
    out = dict(rois=rois, scores=scores, anchor_idx=aidx, counts=counts, post_nms_topN=post)
    if want_decoded:
        out["decoded"] = decoded
    return out


def hot_path_forward(feat, cls_prob, bbox_pred, im_info, base_anchors, feat_stride, pre_nms_topN,
                     post_nms_topN, nms_thresh, min_size, pooled_height, pooled_width,
                     spatial_scale, bin_mode="cpu", need_argmax=True, out=None, nms_mode=NMS_GE_F64,
                     rois_ready=None):
    """proposal_layer -> roi_pool as ONE call (wssdl_hot_path_fwd; VGGnet_test_bus.py:57-62).
    feat [B,H,W,C], cls_prob [B,H,W,2A], bbox_pred [B,H,W,4A], im_info [B,>=3].  Returns the dict
    of proposals() (without anchor_idx) + top / argmax [B*post,PH,PW,C].  Unused rows of an image's
    post_nms_topN RoI slots carry batch index -1 and pool to zeros / -1.
    rois_ready: a torch.cuda.Event (recorded at least once before) that the call records between
    its two stages, when rois / scores / counts are final."""
    feat = _cuda(feat, torch.float32)
    dev = feat.device
    cls_prob = _cuda(cls_prob, torch.float32, dev)
    bbox_pred = _cuda(bbox_pred, torch.float32, dev)
    im_info = _cuda(im_info, torch.float32, dev)
    if im_info.dim() == 1:
        im_info = im_info.reshape(1, -1)
    base = np.ascontiguousarray(base_anchors, dtype=np.float32)
    A = base.shape[0]
    if feat.dim() != 4 or cls_prob.dim() != 4:
        raise ValueError("feat [B,H,W,C], cls_prob [B,H,W,2A]")
    B, H, W, C = feat.shape
    if (tuple(cls_prob.shape) != (B, H, W, 2 * A) or tuple(bbox_pred.shape) != (B, H, W, 4 * A)
            or im_info.shape[0] != B):
        raise ValueError("shape mismatch: feat [B,H,W,C], cls_prob [B,H,W,2A], bbox_pred [B,H,W,4A], "
                         "im_info [B,3+]")
    post, PH, PW = int(post_nms_topN), int(pooled_height), int(pooled_width)
    if post <= 0:
        raise ValueError("the fused hot path needs post_nms_topN > 0 (the stride of the RoI blob)")
    with torch.cuda.device(dev):
        if out is not None:
            rois, scores, counts = out
            if (tuple(rois.shape) != (B * post, 5) or tuple(scores.shape) != (B * post,) or
                    tuple(counts.shape) != (B,) or rois.dtype != torch.float32 or
                    scores.dtype != torch.float32 or counts.dtype != torch.int32 or
                    not (rois.is_contiguous() and scores.is_contiguous() and counts.is_contiguous())
                    or rois.device != dev):
                raise ValueError("out = (rois [B*post,5] f32, scores [B*post] f32, counts [B] i32), "
                                 "contiguous, on the inputs' device")
        else:
            rois = torch.empty((B * post, 5), dtype=torch.float32, device=dev)
            scores = torch.empty((B * post,), dtype=torch.float32, device=dev)
            counts = torch.empty((B,), dtype=torch.int32, device=dev)
        top = torch.empty((B * post, PH, PW, C), dtype=torch.float32, device=dev)
        argmax = torch.empty_like(top, dtype=torch.int32) if need_argmax else None
        ws = _workspace(_lib.lib().wssdl_hot_path_fwd_workspace_bytes(B, post, PH, PW), dev)
        rc = _lib.lib().wssdl_hot_path_fwd(
            _ptr(feat), _ptr(cls_prob), _ptr(bbox_pred), _ptr(im_info), im_info.shape[1], B, H, W, C, A,
            base.ctypes.data_as(_vp), int(feat_stride), int(pre_nms_topN), post, float(nms_thresh),
            int(nms_mode), float(min_size), PH, PW, float(spatial_scale), _bin_mode(bin_mode),
            _ptr(rois), _ptr(scores), _ptr(counts), _ptr(top), _ptr(argmax), _vp(ws.data_ptr()),
            ws.numel(), _stream(dev), _vp(rois_ready.cuda_event) if rois_ready is not None else None)
    _lib.check(rc, "wssdl_hot_path_fwd")
    return dict(rois=rois, scores=scores, counts=counts, post_nms_topN=post, top=top, argmax=argmax)


def compact_rois(out):
    """[B*post,5] fixed-stride blob + counts -> the reference's concatenated (sum R,5) blob
    (proposal_layer_tf_bus.py:144-146).  Synchronises (reads counts)."""
    post = out["post_nms_topN"]
    counts = out["counts"].cpu().numpy()
    rois = out["rois"]
    parts = [rois[b * post:b * post + int(c)] for b, c in enumerate(counts)]
    return torch.cat(parts, dim=0) if parts else rois[:0]


# ------------------------------------------------------------------ anchor labels
def anchor_labels(gt_boxes, num_gt, im_info, H, W, base_anchors, feat_stride, dataset_mode=0,
                  positive_overlap=0.7, negative_overlap=0.3, clobber_positives=False,
                  want_max_overlap=True):
    """Deterministic part of anchor_target_layer[_joint].  gt_boxes [B,max_gt,5] f32, num_gt
    [B] i32, im_info [B,>=2].  Returns (labels [B,H*W*A] f32, argmax_gt i32, max_overlap f64)."""
    gt_boxes = _cuda(gt_boxes, torch.float32)
    dev = gt_boxes.device
    num_gt = _cuda(num_gt, torch.int32, dev)
    im_info = _cuda(im_info, torch.float32, dev)
    if im_info.dim() == 1:
        im_info = im_info.reshape(1, -1)
    base = np.ascontiguousarray(base_anchors, dtype=np.float32)
    A = base.shape[0]
    B, max_gt, five = gt_boxes.shape
    if five != 5 or num_gt.shape[0] != B or im_info.shape[0] != B:
        raise ValueError("gt_boxes [B,max_gt,5], num_gt [B], im_info [B,2+]")
    NA = H * W * A
    with torch.cuda.device(dev):
        labels = torch.empty((B, NA), dtype=torch.float32, device=dev)
        argmax = torch.empty((B, NA), dtype=torch.int32, device=dev)
        maxov = torch.empty((B, NA), dtype=torch.float64, device=dev) if want_max_overlap else None
        nbytes = _lib.lib().wssdl_anchor_labels_workspace_bytes(B, H, W, A, max_gt)
        ws = _workspace(nbytes, dev)
        rc = _lib.lib().wssdl_anchor_labels(
            _ptr(gt_boxes), _ptr(num_gt), max_gt, _ptr(im_info), im_info.shape[1], B, H, W, A,
            base.ctypes.data_as(_vp), int(feat_stride), int(dataset_mode),
            float(positive_overlap), float(negative_overlap), int(bool(clobber_positives)),
            _ptr(labels), _ptr(argmax), _ptr(maxov), _vp(ws.data_ptr()), ws.numel(), _stream(dev))
    _lib.check(rc, "wssdl_anchor_labels")
    return labels, argmax, maxov


def anchor_label_counts(labels):
    """labels [B,NA] f32 (as written by anchor_labels) -> counts [B,2] i32 on the device:
    fg (== 1) and bg (== 0) anchors per image: the population sizes the host RNG needs to draw
    the reference's npr.choice subsamples (anchor_target_layer_tf_bus.py:514, :524)."""
    labels = _cuda(labels, torch.float32)
    B, NA = labels.shape
    with torch.cuda.device(labels.device):
        counts = torch.empty((B, 2), dtype=torch.int32, device=labels.device)
        rc = _lib.lib().wssdl_anchor_label_counts(_ptr(labels), B, NA, _ptr(counts),
                                                  _stream(labels.device))
    _lib.check(rc, "wssdl_anchor_label_counts")
    return counts


def anchor_targets(labels_pre, argmax_gt, gt_boxes, B_total, H, W, base_anchors, feat_stride, num_fg,
                   batchsize, inside_weights=(1.0, 1.0, 1.0, 1.0), positive_weight=-1.0, ranks=None,
                   rank_off=None, seed=None):
    """Sampled part of anchor_target_layer[_joint] on the device (csrc/targets.cu): subsampling,
    fp64 regression targets, weights, `_unmap` and the final layouts for the whole batch.

    labels_pre / argmax_gt [Bs,NA], gt_boxes [Bs,max_gt,5]: outputs / input of anchor_labels.
    Sampling: ranks + rank_off (host-drawn disable ranks: the parity mode) or seed (Philox on
    the device, no host trip).  Returns device tensors (labels [B_total,1,A*H,W], targets, inside,
    outside [B_total,4A,H,W], final_counts [Bs,2] i32)."""
    labels_pre = _cuda(labels_pre, torch.float32)
    dev = labels_pre.device
    argmax_gt = _cuda(argmax_gt, torch.int32, dev)
    gt_boxes = _cuda(gt_boxes, torch.float32, dev)
    base = np.ascontiguousarray(base_anchors, dtype=np.float32)
    A = base.shape[0]
    Bs = labels_pre.shape[0]
    NA = H * W * A
    if Bs and (tuple(labels_pre.shape) != (Bs, NA) or tuple(argmax_gt.shape) != (Bs, NA) or
               gt_boxes.dim() != 3 or gt_boxes.shape[0] != Bs or gt_boxes.shape[2] != 5):
        raise ValueError("labels_pre / argmax_gt [Bs,H*W*A], gt_boxes [Bs,max_gt,5]")
    if (ranks is None) == (seed is None):
        raise ValueError("pass either host-drawn ranks (+ rank_off) or a Philox seed")
    iw = np.ascontiguousarray(inside_weights, dtype=np.float32)
    with torch.cuda.device(dev):
        if ranks is not None:
            mode = _lib.SAMPLE_RANKS
            ranks = _cuda(np.asarray(ranks, dtype=np.int32) if not torch.is_tensor(ranks) else ranks,
                          torch.int32, dev)
            rank_off = _cuda(np.asarray(rank_off, dtype=np.int32) if not torch.is_tensor(rank_off)
                             else rank_off, torch.int32, dev)
            if rank_off.numel() != 2 * Bs + 1:
                raise ValueError("rank_off must hold 2*Bs+1 offsets")
        else:
            mode = _lib.SAMPLE_PHILOX
        labels = torch.empty((B_total, 1, A * H, W), dtype=torch.float32, device=dev)
        targets = torch.empty((B_total, 4 * A, H, W), dtype=torch.float32, device=dev)
        inside = torch.empty_like(targets)
        outside = torch.empty_like(targets)
        fc = torch.empty((max(Bs, 1), 2), dtype=torch.int32, device=dev)
        rc = _lib.lib().wssdl_anchor_targets(
            _ptr(labels_pre), _ptr(argmax_gt), _ptr(gt_boxes), max(int(gt_boxes.shape[1]), 1) if Bs else 1,
            Bs, int(B_total), H, W, A, base.ctypes.data_as(_vp), int(feat_stride), int(num_fg),
            int(batchsize), mode, _ptr(ranks), _ptr(rank_off), int(seed or 0) & (2 ** 64 - 1),
            iw.ctypes.data_as(_vp), float(positive_weight), _ptr(labels), _ptr(targets), _ptr(inside),
            _ptr(outside), _ptr(fc), _stream(dev))
    _lib.check(rc, "wssdl_anchor_targets")
    return labels, targets, inside, outside, fc[:Bs]


class RoiMatch(object):
    """Candidate table of roi_match (kept on the device for roi_targets)."""
    __slots__ = ("rois", "gt_boxes", "workspace", "counts", "n_supervised")

    def __init__(self, rois, gt_boxes, workspace, counts, n_supervised):
        self.rois, self.gt_boxes, self.workspace = rois, gt_boxes, workspace
        self.counts, self.n_supervised = counts, n_supervised


def roi_match(rois, gt_boxes, num_gt, add_gt, fg_thresh, bg_thresh_hi, bg_thresh_lo):
    """First half of _sample_rois for every supervised image (csrc/targets.cu roi_match_kernel):
    candidates = the image's RoIs (+ its fg GT rows when add_gt), fp64 IoU, max / argmax, fg / bg
    candidacy and ranks.  rois [R,5], gt_boxes [Bs,max_gt,5], num_gt [Bs].  Returns a RoiMatch whose
    .counts [Bs,4] i32 = (candidates, fg candidates, bg candidates, fg GT rows)."""
    gt_boxes = _cuda(gt_boxes, torch.float32)
    dev = gt_boxes.device
    rois = _cuda(rois, torch.float32, dev)
    num_gt = _cuda(num_gt, torch.int32, dev)
    if rois.dim() != 2 or rois.shape[1] != 5 or gt_boxes.dim() != 3 or gt_boxes.shape[2] != 5:
        raise ValueError("rois [R,5], gt_boxes [Bs,max_gt,5]")
    Bs, max_gt, R = int(gt_boxes.shape[0]), int(gt_boxes.shape[1]), int(rois.shape[0])
    with torch.cuda.device(dev):
        nbytes = int(_lib.lib().wssdl_roi_targets_workspace_bytes(Bs, R, max_gt))
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        counts = torch.zeros((max(Bs, 1), 4), dtype=torch.int32, device=dev)
        rc = _lib.lib().wssdl_roi_match(_ptr(rois), R, _ptr(gt_boxes), _ptr(num_gt), max(max_gt, 1), Bs,
                                        int(bool(add_gt)), float(fg_thresh), float(bg_thresh_hi),
                                        float(bg_thresh_lo), _ptr(ws), nbytes, _ptr(counts), _stream(dev))
    _lib.check(rc, "wssdl_roi_match")
    return RoiMatch(rois, gt_boxes, ws, counts[:Bs], Bs)


def roi_targets(match, num_classes, inside_weights=(1.0, 1.0, 1.0, 1.0), normalize_means=None,
                normalize_stds=None, sel=None, sel_off=None, row_off=None, n_rows=None,
                fg_rois_per_image=0, rois_per_image=0, seed=None):
    """Second half of _sample_rois + _compute_targets + _get_bbox_regression_labels on the device.
    Selection: sel / sel_off / row_off (+ n_rows, host-drawn ranks: the parity mode) or seed with
    fg_rois_per_image / rois_per_image (Philox on the device).  Returns device tensors (rois [N,5],
    labels [N,1], targets, inside, outside [N,4K], out_counts [Bs,2] i32 or None)."""
    dev = match.gt_boxes.device
    Bs, K = match.n_supervised, int(num_classes)
    if (sel is None) == (seed is None):
        raise ValueError("pass either host-drawn ranks (sel, sel_off, row_off) or a Philox seed")
    iw = np.ascontiguousarray(inside_weights, dtype=np.float32)
    means = stds = None
    if normalize_means is not None:
        means = np.ascontiguousarray(normalize_means, dtype=np.float64)
        stds = np.ascontiguousarray(normalize_stds, dtype=np.float64)
    with torch.cuda.device(dev):
        oc = None
        if sel is not None:
            mode = _lib.SAMPLE_RANKS
            packed = np.concatenate([np.asarray(sel_off, np.int32), np.asarray(row_off, np.int32),
                                     np.asarray(sel, np.int32)])
            if len(sel_off) != 2 * Bs + 1 or len(row_off) != Bs + 1:
                raise ValueError("sel_off holds 2*Bs+1 offsets, row_off Bs+1")
            packed_d = _cuda(packed, torch.int32, dev)          # one H2D for all three
            sel_off_d, row_off_d = packed_d[:2 * Bs + 1], packed_d[2 * Bs + 1:3 * Bs + 2]
            sel_d = packed_d[3 * Bs + 2:]
            N = int(row_off[-1]) if n_rows is None else int(n_rows)
        else:
            mode = _lib.SAMPLE_PHILOX
            sel_d = sel_off_d = row_off_d = None
            N = Bs * int(rois_per_image)
            oc = torch.empty((max(Bs, 1), 2), dtype=torch.int32, device=dev)
        out_rois = torch.empty((N, 5), dtype=torch.float32, device=dev)
        labels = torch.empty((N, 1), dtype=torch.float32, device=dev)
        targets = torch.empty((N, 4 * K), dtype=torch.float32, device=dev)
        inside = torch.empty_like(targets)
        outside = torch.empty_like(targets)
        if N == 0 or Bs == 0:
            return out_rois, labels, targets, inside, outside, (oc[:Bs] if oc is not None else None)
        rc = _lib.lib().wssdl_roi_targets(
            _ptr(match.rois), int(match.rois.shape[0]), _ptr(match.gt_boxes),
            max(int(match.gt_boxes.shape[1]), 1), Bs, K, _ptr(match.workspace), _ptr(match.counts), mode,
            _ptr(sel_d), _ptr(sel_off_d), _ptr(row_off_d),
            int(fg_rois_per_image), int(rois_per_image), int(seed or 0) & (2 ** 64 - 1),
            means.ctypes.data_as(_vp) if means is not None else None,
            stds.ctypes.data_as(_vp) if stds is not None else None, iw.ctypes.data_as(_vp),
            _ptr(out_rois), _ptr(labels), _ptr(targets), _ptr(inside), _ptr(outside), _ptr(oc),
            _stream(dev))
    _lib.check(rc, "wssdl_roi_targets")
    return out_rois, labels, targets, inside, outside, (oc[:Bs] if oc is not None else None)


# ------------------------------------------------------------------ detection post-processing
def detect_postprocess(rois, scores, bbox_pred, im_meta, roi_counts=None, roi_stride=None,
                       score_thresh=0.05, nms_thresh=0.3, max_per_image=300, cls_agnostic=False,
                       want_pred_boxes=False):
    """Batched tail of im_detect + per-image body of test_net (fast_rcnn/test_bus.py:207-223,
    :360-401).  rois [B*S,5] (scaled frame, image b = rows [b*S, b*S+roi_counts[b])), scores
    [B*S,K], bbox_pred [B*S,4K], im_meta [B,3] = (im_h, im_w of the unscaled image, im_scale).
    Returns dict of device tensors: dets [B,K,S,5] (x1,y1,x2,y2,score; class-j rows in
    descending-score order), counts [B,K] i32, status i32 [1] (+ pred_boxes [B*S,4K])."""
    scores = _cuda(scores, torch.float32)
    dev = scores.device
    rois = _cuda(rois, torch.float32, dev)
    bbox_pred = _cuda(bbox_pred, torch.float32, dev)
    im_meta = _cuda(im_meta, torch.float32, dev)
    if im_meta.dim() == 1:
        im_meta = im_meta.reshape(1, -1)
    B = im_meta.shape[0]
    if rois.dim() != 2 or rois.shape[1] != 5 or scores.dim() != 2:
        raise ValueError("rois [B*S,5], scores [B*S,K]")
    K = scores.shape[1]
    S = int(roi_stride) if roi_stride is not None else (rois.shape[0] // max(B, 1))
    if (rois.shape[0] != B * S or scores.shape[0] != B * S or im_meta.shape[1] != 3
            or tuple(bbox_pred.shape) != (B * S, 4 * K)):
        raise ValueError("shape mismatch: rois [B*S,5], scores [B*S,K], bbox_pred [B*S,4K], "
                         "im_meta [B,3]")
    counts_in = _cuda(roi_counts, torch.int32, dev) if roi_counts is not None else None
    with torch.cuda.device(dev):
        dets = torch.empty((B, K, S, 5), dtype=torch.float32, device=dev)
        counts = torch.empty((B, K), dtype=torch.int32, device=dev)
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
        pred = (torch.empty((B * S, 4 * K), dtype=torch.float32, device=dev)
                if want_pred_boxes else None)
        rc = _lib.lib().wssdl_detect_postprocess(
            _ptr(rois), _ptr(counts_in), S, _ptr(scores), _ptr(bbox_pred), _ptr(im_meta), B, K,
            float(score_thresh), float(nms_thresh), int(max_per_image), int(bool(cls_agnostic)),
            _ptr(dets), _vp(counts.data_ptr()), _ptr(pred), _vp(status.data_ptr()), _stream(dev))
    _lib.check(rc, "wssdl_detect_postprocess")
    out = dict(dets=dets, counts=counts, status=status)
    if want_pred_boxes:
        out["pred_boxes"] = pred
    return out


# ------------------------------------------------------------------ evaluation: detection matching
def eval_match(dets, det_counts, gt_boxes, num_gt, difficult=None, ovthresh=0.5, score_thresh=0.5):
    """Per-detection TP / FP / FROC flags and per-image CorLoc flags (voc_eval_bus.py:161-247)
    for the blob of detect_postprocess.  dets [B,K,S,5], det_counts [B,K], gt_boxes [B,G,5]
    (x1,y1,x2,y2,cls), num_gt [B], difficult [B,G] bool/u8 or None.  Returns dict of device
    tensors: tp, fp, fp_froc [B,K,S] u8, img_stats [B,K,2] i32, npos [K] i32."""
    dets = _cuda(dets, torch.float32)
    dev = dets.device
    det_counts = _cuda(det_counts, torch.int32, dev)
    gt_boxes = _cuda(gt_boxes, torch.float32, dev)
    num_gt = _cuda(num_gt, torch.int32, dev)
    if dets.dim() != 4 or dets.shape[3] != 5 or gt_boxes.dim() != 3 or gt_boxes.shape[2] != 5:
        raise ValueError("dets [B,K,S,5], gt_boxes [B,G,5]")
    B, K, S, _ = dets.shape
    G = gt_boxes.shape[1]
    if tuple(det_counts.shape) != (B, K) or gt_boxes.shape[0] != B or num_gt.shape[0] != B:
        raise ValueError("det_counts [B,K], gt_boxes [B,G,5], num_gt [B]")
    diff = _cuda(np.asarray(difficult, dtype=np.uint8) if isinstance(difficult, (list, np.ndarray))
                 else difficult, torch.uint8, dev) if difficult is not None else None
    if diff is not None and tuple(diff.shape) != (B, G):
        raise ValueError("difficult [B,G]")
    with torch.cuda.device(dev):
        tp = torch.zeros((B, K, S), dtype=torch.uint8, device=dev)
        fp = torch.zeros_like(tp)
        ff = torch.zeros_like(tp)
        stats = torch.zeros((B, K, 2), dtype=torch.int32, device=dev)
        npos = torch.zeros((K,), dtype=torch.int32, device=dev)
        rc = _lib.lib().wssdl_eval_match(_ptr(dets), _vp(det_counts.data_ptr()), B, K, S,
                                         _ptr(gt_boxes), _vp(num_gt.data_ptr()), _ptr(diff), G,
                                         float(ovthresh), float(score_thresh), _vp(tp.data_ptr()),
                                         _vp(fp.data_ptr()), _vp(ff.data_ptr()),
                                         _vp(stats.data_ptr()), _vp(npos.data_ptr()), _stream(dev))
    _lib.check(rc, "wssdl_eval_match")
    return dict(tp=tp, fp=fp, fp_froc=ff, img_stats=stats, npos=npos)
