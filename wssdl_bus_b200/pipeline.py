"""Batched detector hot path: RPN outputs + feature maps -> proposals -> RoI pooling.

This is the composition the reference executes per image inside one `sess.run`
(VGGnet_test_bus.py:57-62: proposal_layer -> roi_pool), batched over images and sharded by
image across GPUs.  Images are independent on this path (proposal_layer_tf_bus.py:75 loops
them, RoIs carry their batch index), so ranks never exchange data while computing; the only
collective is one all-gather of the fixed-stride per-image detections for evaluation.
"""
import numpy as np
import torch

from . import ops
from .fast_rcnn.config import cfg
from .rpn_msr.generate_anchors import generate_anchors


def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPUs of the NUMA node its GPU hangs off, BEFORE it pins
    host memory: pinned pages are placed on the node of the allocating thread, and with one
    process per GPU an unbound launcher puts every rank's staging buffers on one socket (the
    8-GPU e2e leg then crosses the inter-socket link for most of its 16 GB per step).
    Returns the node, or None when the topology is not visible (containers may hide sysfs)."""
    import os
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def shard_images(n_images, rank, world_size):
    """Image i belongs to rank i mod world_size (round robin)."""
    return np.arange(rank, n_images, world_size, dtype=np.int64)


class HotPath:
    """proposal_layer -> roi_pool for a batch of images resident on one GPU."""

    def __init__(self, is_training=False, pooled_h=7, pooled_w=7, spatial_scale=1.0 / 16,
                 feat_stride=16, anchor_scales=(8, 16, 32), pre_nms_topN=None, post_nms_topN=None,
                 nms_thresh=None, min_size=None, bin_mode="cpu"):
        c = cfg['TRAIN' if is_training else 'TEST']
        self.pre = c.RPN_PRE_NMS_TOP_N if pre_nms_topN is None else pre_nms_topN
        self.post = c.RPN_POST_NMS_TOP_N if post_nms_topN is None else post_nms_topN
        self.thresh = c.RPN_NMS_THRESH if nms_thresh is None else nms_thresh
        self.min_size = c.RPN_MIN_SIZE if min_size is None else min_size
        self.pooled_h, self.pooled_w, self.scale = pooled_h, pooled_w, spatial_scale
        self.feat_stride = feat_stride
        self.base = generate_anchors(scales=np.array(anchor_scales))
        self.bin_mode = bin_mode
        # host calls of one run(): the fused wssdl_hot_path_fwd entry (proposals kernel + bin-sort
        # pre-pass + pooling kernel on the C4 shapes)
        self.launches_per_run = 3

    def run(self, feat, cls_prob, bbox_pred, im_info, need_argmax=True, blob=None, fused=True,
            rois_ready=None):
        """Device tensors in, device tensors out, no synchronisation.
        feat [B,H,W,C], cls_prob [B,H,W,2A], bbox_pred [B,H,W,4A], im_info [B,3+].
        blob: a DetectionBlob the proposals are written into (for the all-gather).
        fused (default): one wssdl_hot_path_fwd call; rows >= counts[b] of an image's `post` RoI
        slots carry batch index -1 and their pooled rows are zeros / argmax -1.
        rois_ready: torch.cuda.Event recorded when rois / scores / counts are final (before the
        pooling), so that their all-gather can run on another stream meanwhile.
        fused=False: wssdl_proposals then wssdl_roi_pool_fwd, the two public ops back to back (the
        unused rows are then zero RoIs of image 0 and pool its cell (0,0))."""
        outv = None if blob is None else blob.views()
        if fused and self.post > 0:
            return ops.hot_path_forward(feat, cls_prob, bbox_pred, im_info, self.base,
                                        self.feat_stride, self.pre, self.post, self.thresh,
                                        self.min_size, self.pooled_h, self.pooled_w, self.scale,
                                        self.bin_mode, need_argmax, out=outv, rois_ready=rois_ready)
        p = ops.proposals(cls_prob, bbox_pred, im_info, self.base, self.feat_stride, self.pre,
                          self.post, self.thresh, self.min_size, out=outv)
        if rois_ready is not None:
            rois_ready.record()
        top, argmax = ops.roi_pool_forward(feat, p["rois"], self.pooled_h, self.pooled_w,
                                           self.scale, self.bin_mode, need_argmax)
        p["top"], p["argmax"] = top, argmax
        return p

    def detections(self, p):
        """Fixed-stride per-image detections: boxes [B, post, 5] (batch, x1,y1,x2,y2), scores
        [B, post], counts [B] -- views of the kernel outputs (no copy, no extra launch); these
        are the tensors that are all-gathered."""
        B = p["counts"].shape[0]
        return p["rois"].view(B, self.post, 5), p["scores"].view(B, self.post), p["counts"]


class DetectionBlob:
    """The per-image detections of one rank as ONE contiguous f32 buffer, so that a single
    all-gather moves them (boxes + scores + counts in one blob):

        [ rois   n_local*post*5 f32 | scores n_local*post f32 | counts n_local i32 ]

    (sections padded to 16 bytes).  `rois`, `scores`, `counts` are views the proposals kernel
    writes directly (ops.proposals(out=blob.views())): no packing kernel, no concatenation."""

    def __init__(self, n_local, post, device=None, buffer=None):
        self.n_local, self.post = int(n_local), int(post)
        r4 = lambda v: (v + 3) // 4 * 4  # noqa: E731
        self.o_scores = r4(self.n_local * self.post * 5)
        self.o_counts = self.o_scores + r4(self.n_local * self.post)
        self.numel = self.o_counts + r4(self.n_local)
        self.buffer = (torch.zeros((self.numel,), dtype=torch.float32, device=device)
                       if buffer is None else buffer)
        assert self.buffer.numel() == self.numel and self.buffer.dtype == torch.float32

    @staticmethod
    def _split(buf, n_local, post, o_scores, o_counts):
        lead = tuple(buf.shape[:-1])
        rois = buf[..., :n_local * post * 5].reshape(lead + (n_local, post, 5))
        scores = buf[..., o_scores:o_scores + n_local * post].reshape(lead + (n_local, post))
        counts = buf[..., o_counts:o_counts + n_local].view(torch.int32)
        return rois, scores, counts

    def views(self):
        """(rois [n_local*post,5], scores [n_local*post], counts [n_local] i32): views of the buffer."""
        rois, scores, counts = self._split(self.buffer, self.n_local, self.post, self.o_scores,
                                           self.o_counts)
        return rois.view(self.n_local * self.post, 5), scores.view(-1), counts

    def all_gather(self, group=None, async_op=False):
        """ONE all_gather_into_tensor of the buffer (NCCL over NVLink on GPUs, gloo in the CPU
        tests).  Returns ((rois [world,n_local,post,5], scores [world,n_local,post], counts
        [world,n_local] i32), work-or-None); image of slot (r, j) is j*world + r under
        shard_images().  With async_op the collective runs on the communicator's stream behind
        the work already enqueued on the current one; wait() before reading."""
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            g, work = self.buffer[None], None
        else:
            g = torch.empty((world * self.numel,), dtype=torch.float32, device=self.buffer.device)
            work = dist.all_gather_into_tensor(g, self.buffer, group=group, async_op=async_op)
            g = g.view(world, self.numel)
            if not async_op:
                work = None
        return self._split(g, self.n_local, self.post, self.o_scores, self.o_counts), work


class PipelinedHotPath:
    """The hot path software-pipelined over CONSECUTIVE batches on two streams.

    One image's proposals are a latency-bound kernel (one CTA, or one cluster, per image) that
    leaves most of the GPU idle when the batch is small, and the RoI pooling of a batch cannot
    start before its proposals are done.  Consecutive batches are independent, so batch k+1's
    proposals run on a high-priority stream WHILE batch k is pooled on the other stream: their
    CTAs take SMs as pooling CTAs retire, and the pooling kernel never waits for them.  The RoI
    blobs are double buffered (`depth` slots): batch k+depth's proposals wait for batch k's pooling
    (and for its all-gather) before they overwrite slot k % depth.

    submit() enqueues one batch and returns its outputs (device tensors; complete once `done`
    has fired, or after drain()).  Inputs must be ready on the submitting stream and stay
    untouched until the batch is done.  world_size > 1: the packed detections of every batch are
    all-gathered behind its proposals (one collective, on the communicator's stream)."""

    def __init__(self, hot, n_images, device=None, depth=2, gather=False, group=None,
                 inputs_static=False):
        """inputs_static: the inputs of every submit() are already complete when it is called (a
        benchmark's resident tensors), so the proposals stream need not wait for the submitting
        stream each time (one stream wait less per batch on the host)."""
        self.hot, self.n, self.depth = hot, int(n_images), int(depth)
        self.inputs_static = bool(inputs_static)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.s_prop = torch.cuda.Stream(device=self.dev, priority=-1)
        self.s_pool = torch.cuda.Stream(device=self.dev)
        self.blobs = [DetectionBlob(self.n, hot.post, device=self.dev) for _ in range(self.depth)]
        self.views = [b.views() for b in self.blobs]       # (computed once: six view ops each)
        self.pool_done = [None] * self.depth
        self.gather_work = [None] * self.depth
        self.gather, self.group = bool(gather), group
        self.k = 0

    def submit(self, feat, cls_prob, bbox_pred, im_info, need_argmax=True, marks=None):
        """marks: optional (e0, e1, e2, e3) timing events recorded around the proposals (on the
        proposals stream) and around the RoI-pool stage (on the pooling stream)."""
        hot, b = self.hot, self.k % self.depth
        blob = self.blobs[b]
        if self.k == 0 or not self.inputs_static:
            self.s_prop.wait_stream(torch.cuda.current_stream(self.dev))   # inputs are ready
        with torch.cuda.stream(self.s_prop):
            if self.pool_done[b] is not None:
                self.s_prop.wait_event(self.pool_done[b])  # slot b's RoIs have been pooled
            if self.gather_work[b] is not None:
                self.gather_work[b].wait()                 # ... and gathered
                self.gather_work[b] = None
            if marks:
                marks[0].record()
            p = ops.proposals(cls_prob, bbox_pred, im_info, hot.base, hot.feat_stride, hot.pre,
                              hot.post, hot.thresh, hot.min_size, out=self.views[b],
                              pad_rows_invalid=True)
            if marks:
                marks[1].record()
            ready = torch.cuda.Event()
            ready.record()
            if self.gather:
                p["gathered"], self.gather_work[b] = blob.all_gather(self.group, async_op=True)
        with torch.cuda.stream(self.s_pool):
            self.s_pool.wait_event(ready)
            if marks:
                marks[2].record()
            top, argmax = ops.roi_pool_forward_grouped(feat, p["rois"], hot.post, hot.pooled_h,
                                                       hot.pooled_w, hot.scale, hot.bin_mode,
                                                       need_argmax)
            if marks:
                marks[3].record()
            done = torch.cuda.Event()
            done.record()
        self.pool_done[b] = done
        p["top"], p["argmax"], p["done"] = top, argmax, done
        self.k += 1
        return p

    def drain(self):
        """The submitting stream waits for everything submitted so far (incl. the gathers)."""
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_stream(self.s_pool)
        for b, w in enumerate(self.gather_work):
            if w is not None:
                w.wait()
                self.gather_work[b] = None
        cur.wait_stream(self.s_prop)


def agree_on_faster_mode(t_pipelined_ms, t_fused_ms, device=None, group=None):
    """Every rank passes ITS measured step times of the two ways to run the hot path
    (PipelinedHotPath / one fused call per step); all ranks get the same answer, decided on the
    slowest rank's times (one all-reduce): (pipelined_is_faster, t_pipelined_ms, t_fused_ms).
    A collective every rank has to call."""
    import torch.distributed as dist
    tt = torch.tensor([float(t_pipelined_ms), float(t_fused_ms)], dtype=torch.float64,
                      device=device if device is not None else "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX, group=group)
    a, b = (float(v) for v in tt.tolist())
    return a < b, a, b


def all_gather_blobs(tensors, group=None, async_op=False):
    """all_gather_into_tensor of several [n_local, ...] tensors (NCCL over NVLink on GPUs,
    gloo in the CPU tests): -> list of [world, n_local, ...].

    async_op=True returns (list, works): the collectives are enqueued behind the work already
    on the current stream and run on the communicator's own stream, so kernels launched
    afterwards on the current stream overlap them (the detections depend on the proposals
    only, so a step gathers them WHILE its RoI pooling runs); call wait() on every work before
    the gathered tensors are read."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out, works = [], []
    for t in tensors:
        if world == 1:
            out.append(t[None])
            continue
        g = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        w = dist.all_gather_into_tensor(g, t.contiguous(), group=group, async_op=async_op)
        if async_op:
            works.append(w)
        out.append(g.reshape((world, t.shape[0]) + tuple(t.shape[1:])))
    return (out, works) if async_op else out


def all_gather_detections(det, counts, group=None):
    """One all-gather of [n_local, post, 5] f32 + [n_local] i32 (NCCL over NVLink on GPUs,
    gloo in the CPU tests).  Every rank must hold the same n_local (pad the last shard).
    Returns ([world, n_local, post, 5], [world, n_local]); image of slot (r, j) is
    j * world + r under shard_images()."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return det[None], counts[None]
    n_local = det.shape[0]
    # concatenated-along-dim-0 output: the form both NCCL and gloo accept
    det_all = torch.empty((world * n_local,) + tuple(det.shape[1:]), dtype=det.dtype,
                          device=det.device)
    cnt_all = torch.empty((world * n_local,), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(det_all, det.contiguous(), group=group)
    dist.all_gather_into_tensor(cnt_all, counts.contiguous(), group=group)
    return (det_all.reshape((world, n_local) + tuple(det.shape[1:])),
            cnt_all.reshape(world, n_local))


def unshard_detections(det_all, cnt_all, n_images):
    """[world, n_local, ...] gathered slots -> per-image order [n_images, ...]."""
    world, n_local = cnt_all.shape
    det = det_all.transpose(0, 1).reshape((n_local * world,) + tuple(det_all.shape[2:]))
    cnt = cnt_all.transpose(0, 1).reshape(n_local * world)
    return det[:n_images], cnt[:n_images]


class HostPipeline:
    """End-to-end form of HotPath.run for HOST buffers (the shape of the reference's py_func
    boundary, network.py:216): pinned host inputs are copied to the device in chunks of
    images, computed, and every output (rois, scores, counts, pooled features, argmax) is
    copied back to pinned host memory.  Two CUDA streams double-buffer the chunks so the
    H2D copy of chunk i+1 and the D2H copy of chunk i-1 overlap the kernels of chunk i."""

    def __init__(self, hot, B, H, W, C, A, info_cols=3, chunk=32, device=None, need_argmax=True,
                 features_to_host=True):
        """features_to_host=False keeps the pooled features (and argmax) on the device -- the
        reference's own arrangement, where fc6 consumes them on the GPU and only the RoIs
        cross the py_func boundary (network.py:216); the kernels still run in full."""
        self.hot, self.B, self.chunk = hot, B, min(chunk, B)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.need_argmax = need_argmax
        self.features_to_host = features_to_host
        post, ph, pw = hot.post, hot.pooled_h, hot.pooled_w
        pin = dict(pin_memory=True)
        self.h_out = dict(
            rois=torch.empty((B * post, 5), dtype=torch.float32, **pin),
            scores=torch.empty((B * post,), dtype=torch.float32, **pin),
            counts=torch.empty((B,), dtype=torch.int32, **pin),
        )
        if features_to_host:
            self.h_out["top"] = torch.empty((B * post, ph, pw, C), dtype=torch.float32, **pin)
            if need_argmax:
                self.h_out["argmax"] = torch.empty((B * post, ph, pw, C), dtype=torch.int32, **pin)
        self.d_last = None      # device outputs of the last chunk (features_to_host=False)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(2)]
        n = self.chunk
        dev = self.device
        self.d_in = [dict(feat=torch.empty((n, H, W, C), device=dev),
                          cls=torch.empty((n, H, W, 2 * A), device=dev),
                          reg=torch.empty((n, H, W, 4 * A), device=dev),
                          info=torch.empty((n, info_cols), device=dev)) for _ in range(2)]
        self._slot = torch.arange(post, device=dev, dtype=torch.int32)
        self.h2d_bytes = 4 * B * (H * W * (C + 6 * A) + info_cols)
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.h_out.values())

    def run(self, h_feat, h_cls, h_reg, h_info):
        """Pinned host tensors in -> dict of pinned host tensors out.  Synchronises."""
        post = self.hot.post
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        for ci, b0 in enumerate(range(0, self.B, self.chunk)):
            b1 = min(b0 + self.chunk, self.B)
            n = b1 - b0
            s = self.streams[ci % 2]
            d = self.d_in[ci % 2]
            with torch.cuda.stream(s):
                d["feat"][:n].copy_(h_feat[b0:b1], non_blocking=True)
                d["cls"][:n].copy_(h_cls[b0:b1], non_blocking=True)
                d["reg"][:n].copy_(h_reg[b0:b1], non_blocking=True)
                d["info"][:n].copy_(h_info[b0:b1], non_blocking=True)
                p = self.hot.run(d["feat"][:n], d["cls"][:n], d["reg"][:n], d["info"][:n],
                                 self.need_argmax)
                # batch indices are chunk-local on the device; make them global for the host
                # (valid rows only: the slots behind counts[b] keep batch index -1)
                rois = p["rois"]
                valid = (self._slot[None, :] < p["counts"][:, None]).reshape(-1)
                rois[:, 0] += valid.to(rois.dtype) * float(b0)
                self.h_out["rois"][b0 * post:b1 * post].copy_(rois, non_blocking=True)
                self.h_out["scores"][b0 * post:b1 * post].copy_(p["scores"], non_blocking=True)
                self.h_out["counts"][b0:b1].copy_(p["counts"], non_blocking=True)
                if self.features_to_host:
                    self.h_out["top"][b0 * post:b1 * post].copy_(p["top"], non_blocking=True)
                    if self.need_argmax:
                        self.h_out["argmax"][b0 * post:b1 * post].copy_(p["argmax"], non_blocking=True)
                else:
                    self.d_last = p
        for s in self.streams:
            cur.wait_stream(s)
        cur.synchronize()
        return self.h_out
