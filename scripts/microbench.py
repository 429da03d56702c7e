#!/usr/bin/env python
"""Per-kernel timings for the hot path on one GPU (CUDA events, warm-up, L2-cold inputs where
the working set allows).  Writes JSON lines to stdout / --out.  Used to fill BASELINE.md and
to compare kernel variants; bench.py stays the contract benchmark."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wssdl_bus_b200 import _lib, ops, synthetic as syn  # noqa: E402
from wssdl_bus_b200.pipeline import HotPath  # noqa: E402

PEAK = 6553.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, iters=20, warmup=5, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def emit(out, **kw):
    line = json.dumps(kw)
    print(line, flush=True)
    if out:
        out.write(line + "\n")
        out.flush()


def realistic_rois(B, post=300, seed=0):
    """RoIs as the proposal layer emits them for the synthetic RPN outputs."""
    cls, reg, info = syn.rpn_outputs(seed, B, 38, 50, 9)
    hot = HotPath(post_nms_topN=post)
    p = ops.proposals(cls, reg, info, hot.base, 16, hot.pre, post, hot.thresh, hot.min_size)
    return p["rois"]


def bench_roi(out, tag, B, C, PH, PW, rois, mode="cpu", bwd=True):
    H, W = 38, 50
    x = torch.from_numpy(syn.feature_map(1, B, H, W, C)).cuda()
    R = rois.shape[0]
    fwd_bytes = B * H * W * C * 4 + R * 20 + R * PH * PW * C * 8
    top, arg = ops.roi_pool_forward(x, rois, PH, PW, 1 / 16., mode)
    med, best = timeit(lambda: ops.roi_pool_forward(x, rois, PH, PW, 1 / 16., mode),
                       flush=fwd_bytes < (1 << 30))
    emit(out, op="roi_pool_fwd", tag=tag, B=B, C=C, R=R, PH=PH, bin_mode=mode, ms=med, ms_min=best,
         alg_bytes=fwd_bytes, gbs=fwd_bytes / med / 1e6, frac_measured=fwd_bytes / med / 1e6 / PEAK,
         frac_8000=fwd_bytes / med / 1e6 / 8000)
    if not bwd:
        return
    g = torch.randn_like(top)
    bwd_bytes = R * PH * PW * C * 8 + B * H * W * C * 4
    for det in (False, True):
        med, best = timeit(lambda: ops.roi_pool_backward((B, H, W, C), rois, arg, g, PH, PW, 1 / 16.,
                                                         deterministic=det),
                           iters=10, flush=bwd_bytes < (1 << 30))
        emit(out, op="roi_pool_bwd_" + ("gather" if det else "atomic"), tag=tag, B=B, C=C, R=R, PH=PH,
             ms=med, ms_min=best, alg_bytes=bwd_bytes, gbs=bwd_bytes / med / 1e6,
             frac_measured=bwd_bytes / med / 1e6 / PEAK)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="")
    ap.add_argument("--kernels", default="", help="roicmp: comma list of forward kernels")
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle (1 core)")
    args = ap.parse_args()
    out = open(args.out, "a") if args.out else None
    only = set(args.only.split(",")) if args.only else None

    def want(name):
        return (only is None and name not in ("roi4", "roicmp")) or (only is not None and name in only)

    emit(out, op="env", gpu=torch.cuda.get_device_name(0), peak_gbs=PEAK)
    if want("roi4"):
        r256 = realistic_rois(256)
        bench_roi(out, "C4 256 images x 300 proposal RoIs", 256, 512, 7, 7, r256, bwd=False)
    if want("roicmp"):
        # every forward kernel on every BASELINE shape
        r1 = realistic_rois(1)
        r256 = realistic_rois(256)
        ru = torch.from_numpy(syn.rois_for_pool(5, 16 * 300, 16)).cuda()
        ru = ru[torch.argsort(ru[:, 0], stable=True)].contiguous()
        for kern in (args.kernels.split(",") if args.kernels else ("direct", "tiled", "band", "sorted")):
            _lib.set_tuning("roi_fwd_kernel", kern)
            tag = "[%s] " % kern
            bench_roi(out, tag + "C4 256x300", 256, 512, 7, 7, r256, bwd=False)
            bench_roi(out, tag + "C4 256x300 GPU_CEIL", 256, 512, 7, 7, r256, mode="gpu", bwd=False)
            bench_roi(out, tag + "C1 1x300", 1, 512, 7, 7, r1, bwd=False)
            bench_roi(out, tag + "C2 1x128", 1, 512, 7, 7, r1[:128].contiguous(), bwd=False)
            bench_roi(out, tag + "16x300 7x7 C512", 16, 512, 7, 7, realistic_rois(16), bwd=False)
            bench_roi(out, tag + "C3 16x1024 4800 RoIs 14x14", 16, 1024, 14, 14, ru, bwd=False)
        _lib.set_tuning("roi_fwd_kernel", "auto")
    if want("roi"):
        r1 = realistic_rois(1)
        bench_roi(out, "C1 single image (300 proposal RoIs)", 1, 512, 7, 7, r1)
        bench_roi(out, "C2 train step (128 sampled RoIs)", 1, 512, 7, 7, r1[:128].contiguous())
        r256 = realistic_rois(256)
        bench_roi(out, "C4 256 images x 300 proposal RoIs", 256, 512, 7, 7, r256, bwd=False)
        bench_roi(out, "C4 256 images, GPU_CEIL bins", 256, 512, 7, 7, r256, mode="gpu", bwd=False)
        ru = torch.from_numpy(syn.rois_for_pool(5, 16 * 300, 16)).cuda()
        ru = ru[torch.argsort(ru[:, 0], stable=True)].contiguous()
        bench_roi(out, "C3 ResNet C4 16x1024, 4800 random RoIs 14x14", 16, 1024, 14, 14, ru)
    if want("cpu") and args.cpu:
        import oracle
        c = syn.C1
        feat = syn.feature_map(1, 1, 38, 50, 512)
        cls, reg, info = syn.rpn_outputs(7, 1, 38, 50, 9)

        def best_of(fn, n=3):
            ts = []
            for _ in range(n):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            return min(ts) * 1e3

        blob = oracle.layers.proposal_layer(cls, reg, info)
        for th in (1, oracle.clib.default_threads()):
            emit(out, op="cpu_roi_pool_fwd", tag="C1 1x300", threads=th,
                 ms=best_of(lambda: oracle.clib.roi_pool_fwd(feat, blob, 7, 7, 1 / 16., threads=th)))
        top, arg = oracle.clib.roi_pool_fwd(feat, blob[:128], 7, 7, 1 / 16.)
        g = np.ones_like(top)
        emit(out, op="cpu_roi_pool_bwd", tag="C2 1x128", threads=oracle.clib.default_threads(),
             ms=best_of(lambda: oracle.clib.roi_pool_bwd(g, arg, blob[:128], feat.shape, 1 / 16.)))
        emit(out, op="cpu_proposal_layer", tag="TEST 6000->300", ms=best_of(
            lambda: oracle.layers.proposal_layer(cls, reg, info), n=2))
        emit(out, op="cpu_proposal_layer", tag="TRAIN 12000->2000", ms=best_of(
            lambda: oracle.layers.proposal_layer(cls, reg, info, is_training=True), n=1))
        gt, num = syn.gt_boxes(9, 1)
        emit(out, op="cpu_anchor_labels", B=1, ms=best_of(
            lambda: oracle.layers.anchor_labels(38, 50, gt[0, :num[0]], info[0])))
    if want("proposals"):
        for B in (1, 16, 32, 64, 128, 256):
            cls, reg, info = syn.rpn_outputs(7, B, 38, 50, 9)
            cls, reg, info = [torch.from_numpy(v).cuda() for v in (cls, reg, info)]
            hot = HotPath()
            for pre, post, tag in ((6000, 300, "TEST 6000->300"), (12000, 2000, "TRAIN 12000->2000"),
                                   (2000, 2000, "C2 2000->2000")):
                if B > 16 and post == 2000:
                    continue
                med, best = timeit(lambda: ops.proposals(cls, reg, info, hot.base, 16, pre, post, 0.7, 16),
                                   iters=10, flush=False)
                emit(out, op="proposals", tag=tag, B=B, ms=med, ms_min=best, images_per_s=B / med * 1e3)
    if want("nms"):
        for n in (1000, 2000, 6000, 12000, 20000, 50000, 100000):
            for clustered in (False, True):
                d = torch.from_numpy(syn.dets(11 + n, n, clustered=clustered)).cuda()
                for t in (0.3, 0.5, 0.7):
                    keep, num, _ = ops.nms_device(d, t)
                    med, best = timeit(lambda: ops.nms_device(d, t), iters=5, warmup=2, flush=False)
                    emit(out, op="nms", N=n, clustered=clustered, thresh=t, kept=int(num.item()), ms=med,
                         ms_min=best, pairs=n * (n - 1) // 2, gpairs_per_s=n * (n - 1) / 2 / med / 1e6)
                    if args.cpu and n <= 12000 and t == 0.7:
                        import oracle
                        dn = d.cpu().numpy()
                        t0 = time.perf_counter()
                        fn = oracle.ref.cpu_nms if oracle.ref.available() else oracle.clib.nms
                        fn(dn, t)
                        emit(out, op="cpu_nms", N=n, clustered=clustered, thresh=t,
                             ms=(time.perf_counter() - t0) * 1e3,
                             kind="reference" if oracle.ref.available() else "port")
    if want("iou"):
        for n, k in ((5944, 20), (17100, 20), (100000, 20), (100000, 128), (100000, 1024), (20000, 20000)):
            b = torch.from_numpy(syn.random_boxes(3, n).astype(np.float64)).cuda()
            q = torch.from_numpy(syn.random_boxes(4, k).astype(np.float64)).cuda()
            for dt in (torch.float64, torch.float32):
                if n * k * (8 if dt == torch.float64 else 4) > (8 << 30):
                    continue
                bb, qq = b.to(dt), q.to(dt)
                nbytes = (n + k) * 4 * bb.element_size() + n * k * bb.element_size()
                med, best = timeit(lambda: ops.bbox_overlaps_device(bb, qq, ops.IOU, dt), iters=5,
                                   warmup=2, flush=nbytes < (1 << 30))
                emit(out, op="bbox_overlaps", N=n, K=k, dtype=str(dt), ms=med, ms_min=best, alg_bytes=nbytes,
                     gbs=nbytes / med / 1e6, frac_measured=nbytes / med / 1e6 / PEAK,
                     gpairs_per_s=n * k / med / 1e6)
                if args.cpu and dt == torch.float64 and n * k <= 2e7:
                    import oracle
                    fn = oracle.ref.bbox_overlaps if oracle.ref.available() else oracle.clib.bbox_overlaps
                    bn, qn = b.cpu().numpy(), q.cpu().numpy()
                    t0 = time.perf_counter()
                    fn(bn, qn)
                    emit(out, op="cpu_bbox_overlaps", N=n, K=k, ms=(time.perf_counter() - t0) * 1e3)
    if want("chain"):
        # BASELINE config C1: the single-image chain proposals -> RoI pooling, the latency a
        # serving loop sees per image.  Three ways to issue the same three kernels: the two
        # public ops back to back, the fused entry (one host call, PDL between the kernels), and
        # the fused entry captured once in a CUDA graph and replayed.
        for B in (1, 2, 4):
            feat = torch.from_numpy(syn.feature_map(1, B, 38, 50, 512)).cuda()
            cls, reg, info = [torch.from_numpy(v).cuda() for v in syn.rpn_outputs(7, B, 38, 50, 9)]
            hot = HotPath()
            for tag, fn in (("two ops", lambda: hot.run(feat, cls, reg, info, fused=False)),
                            ("fused entry", lambda: hot.run(feat, cls, reg, info))):
                med, best = timeit(fn, iters=30, flush=False)
                emit(out, op="c1_chain", tag=tag, B=B, ms=med, ms_min=best)
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                for _ in range(3):
                    hot.run(feat, cls, reg, info)
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                res = hot.run(feat, cls, reg, info)
            eager = hot.run(feat, cls, reg, info)
            g.replay()
            torch.cuda.synchronize()
            same = bool(torch.equal(res["top"], eager["top"]) and torch.equal(res["rois"], eager["rois"]))
            med, best = timeit(g.replay, iters=30, flush=False)
            emit(out, op="c1_chain", tag="fused entry, CUDA graph replay", B=B, ms=med, ms_min=best,
                 equals_eager=same)
    if want("train"):
        # BASELINE config C2: the device kernels of one training step, chained on one stream
        # with NO host synchronisation: anchor-target layer (labels + Philox subsampling +
        # targets) -> proposals 12000 -> 2000 -> proposal-target layer (RoI x GT matching, Philox
        # sampling of 128 RoIs per image, targets) -> roi_pool fwd + bwd on the sampled RoIs.
        from wssdl_bus_b200.rpn_msr import anchor_target_layer_tf_bus as atl
        from wssdl_bus_b200.rpn_msr import proposal_target_layer_tf_bus as ptl
        from wssdl_bus_b200.rpn_msr.generate_anchors import generate_anchors
        base = generate_anchors()
        for B in (1, 16):
            feat = torch.from_numpy(syn.feature_map(1, B, 38, 50, 512)).cuda()
            cls, reg, info = [torch.from_numpy(v).cuda() for v in syn.rpn_outputs(7, B, 38, 50, 9)]
            gt, num = [torch.from_numpy(v).cuda() for v in syn.gt_boxes(9, B)]
            num = num.int()
            score = torch.zeros((B, 38, 50, 18), device="cuda")
            gtop = torch.randn((B * 128, 7, 7, 512), device="cuda")

            def step():
                atl.anchor_target_layer(score, gt, num, info, None, [16, ], [8, 16, 32], "SNUBH",
                                        sampler="philox", seed=1, return_device=True)
                p = ops.proposals(cls, reg, info, base, 16, 12000, 2000, 0.7, 16)
                out_ = ptl.proposal_target_layer(p["rois"], gt, num, 3, True, False, sampler="philox",
                                                 seed=2, return_device=True)
                rr = out_[0]
                top, arg = ops.roi_pool_forward(feat, rr, 7, 7, 1 / 16.)
                ops.roi_pool_backward((B, 38, 50, 512), rr, arg, gtop, 7, 7, 1 / 16.)
            with torch.autograd.profiler.profile(enabled=False):
                med, best = timeit(step, iters=10, flush=False)
            emit(out, op="train_step_device_kernels",
                 tag="C2: anchor targets + proposals 12000->2000 + proposal targets (Philox, no D2H) + "
                     "roi_pool fwd+bwd on 128 RoIs/image", B=B, ms=med, ms_min=best,
                 images_per_s=B / med * 1e3)
    if want("detect"):
        for B in (1, 256):
            S, K = 300, 3
            rois = np.concatenate([syn.rois_for_pool(9 + b, S) for b in range(B)])
            scores, deltas = syn.rcnn_head_outputs(9, B * S, K)
            meta = np.tile(np.array([[437, 583, 600.0 / 437]], np.float32), (B, 1))
            a = [torch.from_numpy(v).cuda() for v in (rois, scores, deltas, meta)]
            med, best = timeit(lambda: ops.detect_postprocess(*a, roi_stride=S), iters=10, flush=False)
            emit(out, op="detect_postprocess", B=B, S=S, K=K, ms=med, ms_min=best, images_per_s=B / med * 1e3)
            if args.cpu and B == 1:
                import oracle
                t0 = time.perf_counter()
                pb = oracle.layers.im_detect_boxes(rois, deltas, (437, 583), 600.0 / 437)
                oracle.layers.detections_postprocess(scores, pb, thresh=np.float32(0.05))
                emit(out, op="cpu_detect_postprocess", B=1, ms=(time.perf_counter() - t0) * 1e3)
    if want("labels"):
        from wssdl_bus_b200.rpn_msr.generate_anchors import generate_anchors
        for B in (1, 64):
            gt, num = syn.gt_boxes(9, B)
            info = np.tile(np.array([[600, 800, 1.0]], np.float32), (B, 1))
            gt, num, info = [torch.from_numpy(v).cuda() for v in (gt, num, info)]
            base = generate_anchors()
            med, best = timeit(lambda: ops.anchor_labels(gt, num, info, 38, 50, base, 16), iters=10, flush=False)
            emit(out, op="anchor_labels", B=B, ms=med, ms_min=best)


if __name__ == "__main__":
    main()
