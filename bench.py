#!/usr/bin/env python
"""bench.py -- RoI-pool + proposal images/s on N B200s (BASELINE.json metric), C4 workload.

    python bench.py --gpus N --steps K --warmup W              # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path

One "step" = one pass of the hot path over one batch of synthetic images:
  per image: 17100 anchors -> decode/clip/filter -> top 6000 -> NMS 0.7 -> 300 RoIs ->
  RoI max-pool 7x7 (+argmax) on a 38x50x512 NHWC map          (BASELINE.json configs[3]/[0])
Images are sharded by image over ranks (weak scaling: --images-per-gpu each); the only
collective is the NCCL all-gather of the per-image detections at the end of every step.

Printed JSON (rank 0, one line): value = whole-job images/s with inputs resident in HBM;
e2e = the same through HOST buffers (pinned H2D of every input, D2H of every output inside
the timed region); roofline = RoI-pool forward kernel, algorithmic bytes / CUDA-event time
against MEASURED_PEAKS.json; cpu_baseline = the reference CPU path timed on this box.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from wssdl_bus_b200 import synthetic as syn  # noqa: E402

CFG = syn.C1
METRIC = "roi_pool+proposal images/sec"
UNIT = "images/s"
# map read once + rois + (out + argmax) written, per image (C1)
ROI_POOL_FWD_BYTES_PER_IMAGE = (CFG["H"] * CFG["W"] * CFG["C"] * 4 + CFG["post"] * 5 * 4 +
                                CFG["post"] * CFG["PH"] * CFG["PW"] * CFG["C"] * 8)


def workload_config(images_per_gpu, n_gpus):
    return {
        "workload": "C4 batched inference: per image 38x50x512 NHWC map, 17100 anchors -> "
                    "top-6000 -> NMS 0.7 -> 300 RoIs -> roi_pool 7x7 fwd+argmax; synthetic "
                    "600x800 images",
        "images_per_gpu": images_per_gpu,
        "global_images": images_per_gpu * n_gpus,
        "parallelism": "image-sharded x%d, all-gather of detections per step (overlapped with the RoI pooling)" % n_gpus,
        "l2": "per-step inputs+outputs (%.1f GB/GPU) exceed the 126 MB L2; no flush needed"
              % (images_per_gpu * (ROI_POOL_FWD_BYTES_PER_IMAGE + CFG["H"] * CFG["W"] * 54 * 4) / 1e9),
    }


def make_inputs(n_images, seed0):
    """Host (numpy) inputs for n_images images; per-image seeds so shards differ."""
    feat = syn.feature_map(seed0, n_images, CFG["H"], CFG["W"], CFG["C"])
    cls, reg, info = syn.rpn_outputs(seed0 + 1, n_images, CFG["H"], CFG["W"], CFG["A"])
    return feat, cls, reg, info


# --------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """One image through the reference's CPU path: proposal_layer (numpy glue + the
    reference's own cpu_nms from oracle/_ref when built, else its C restatement) and the
    RoiPool CPU kernel (the reference's own roi_pooling_op.cc compiled into
    oracle/_ref/ref_roi_pool.so when built, else its C restatement), single thread."""
    seed, = args
    import oracle
    feat, cls, reg, info = make_inputs(1, seed)
    t0 = time.perf_counter()
    blob = oracle.layers.proposal_layer(cls, reg, info)
    if oracle.ref.roi_pool_available():
        top, arg = oracle.ref.roi_pool_fwd(feat, blob, CFG["PH"], CFG["PW"], CFG["scale"], threads=1)
    else:
        top, arg = oracle.clib.roi_pool_fwd(feat, blob, CFG["PH"], CFG["PW"], CFG["scale"], threads=1)
    return time.perf_counter() - t0, int(blob.shape[0])


def cpu_arm(steps, warmup, sample_images=None):
    """Times the reference CPU implementation on this box: image-parallel over all host
    cores (each worker = the reference's single-threaded per-image path)."""
    import multiprocessing as mp
    import oracle
    oracle.clib.build()
    kind = "reference" if (oracle.ref.available() and oracle.ref.roi_pool_available()) else "port"
    roi_kind = ("the reference's own RoiPool CPU kernel (roi_pooling_op.cc compiled unmodified)"
                if oracle.ref.roi_pool_available() else "C++ RoiPool CPU kernel restatement")
    cores = max(1, len(os.sched_getaffinity(0)))
    n = sample_images or max(cores, 8)
    n = min(n, 256)
    ctx = mp.get_context("fork")
    with ctx.Pool(processes=min(cores, n)) as pool:
        for w in range(max(warmup, 0)):
            pool.map(_cpu_worker, [(9000 + i,) for i in range(min(cores, n))])
        times = []
        for k in range(steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(1000 * k + i,) for i in range(n)])
            times.append(time.perf_counter() - t0)
    per_image = float(np.mean([r[0] for r in res]))
    ms_step = 1e3 * float(np.mean(times))
    value = n / (ms_step / 1e3)
    sample = ("%d images/step (of the 256-image C4 batch), image-parallel over %d processes; per "
              "image: numpy decode/clip/filter/argsort + %s cpu_nms(6000 boxes, 0.7) + %s "
              "(1 thread); %.2f s/image/core"
              % (n, min(cores, n), "reference Cython" if oracle.ref.available() else "C-port",
                 roi_kind, per_image))
    return dict(value=value, unit=UNIT, cores=min(cores, n), kind=kind, sample=sample), ms_step, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms_step, n = cpu_arm(args.steps, args.warmup, sample_images=args.cpu_sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.images_per_gpu, args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active")
                                                         for r in self.rows if len(r) > 4 + i)]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel, images):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
    (profiles/roofline_traffic.json), only when it was taken on the same batch size."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))[kernel]
        if int(t["images"]) == int(images):
            return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"]), t["source"]
    except Exception:
        pass
    return None, None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from wssdl_bus_b200 import ops
    from wssdl_bus_b200.pipeline import (HostPipeline, HotPath, all_gather_blobs,
                                         bind_to_gpu_numa_node)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)      # before any pinned allocation
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.images_per_gpu
    K, Wm = args.steps, args.warmup

    feat, cls, reg, info = make_inputs(B, 100000 * rank)
    h = [torch.from_numpy(x).pin_memory() for x in (feat, cls, reg, info)]
    d = [x.to(dev) for x in h]
    hot = HotPath()
    post = hot.post

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    roi_ev = [(ev(), ev()) for _ in range(K)]
    prop_ev = [(ev(), ev()) for _ in range(K)]

    def step(k=None):
        if k is not None:
            prop_ev[k][0].record()
        p = ops.proposals(d[1], d[2], d[3], hot.base, hot.feat_stride, hot.pre, hot.post,
                          hot.thresh, hot.min_size)
        if k is not None:
            prop_ev[k][1].record()
        # the detections (boxes, scores, counts) are complete after the proposals: their
        # all-gather runs on the communicator's stream while the RoI pooling runs on this one
        works = []
        if world > 1:
            gathered, works = all_gather_blobs(hot.detections(p), async_op=True)
            p["gathered"] = gathered
        if k is not None:
            roi_ev[k][0].record()
        top, argmax = ops.roi_pool_forward(d[0], p["rois"], hot.pooled_h, hot.pooled_w, hot.scale)
        if k is not None:
            roi_ev[k][1].record()
        p["top"], p["argmax"] = top, argmax
        for w in works:
            w.wait()                                  # the step ends when the gather has landed
        return p

    for _ in range(max(Wm, 3)):
        p = step()
    counts = p["counts"].cpu().numpy()
    del p
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    t0, t1 = ev(), ev()
    t0.record()
    for k in range(K):
        step(k)
    t1.record()
    barrier()
    ms_total = t0.elapsed_time(t1)
    roi_ms = float(np.mean([a.elapsed_time(b) for a, b in roi_ev]))
    prop_ms = float(np.mean([a.elapsed_time(b) for a, b in prop_ev]))

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region
    e2e = e2e_dev = None
    if not args.no_e2e:
        hp = HostPipeline(hot, B, CFG["H"], CFG["W"], CFG["C"], CFG["A"], chunk=args.e2e_chunk,
                          device=dev)
        for _ in range(2):
            out = hp.run(*h)
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(K):
            out = hp.run(*h)
            if world > 1:
                all_gather_blobs([out["rois"].view(B, post, 5).to(dev, non_blocking=True),
                                  out["scores"].view(B, post).to(dev, non_blocking=True),
                                  out["counts"].to(dev, non_blocking=True)])
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        e2e = (e2e_ms, hp.h2d_bytes, hp.d2h_bytes)
        # same call, pooled features left on the device (the reference's arrangement: fc6
        # consumes them on the GPU, only the RoIs cross the py_func boundary)
        del hp, out
        hp2 = HostPipeline(hot, B, CFG["H"], CFG["W"], CFG["C"], CFG["A"], chunk=args.e2e_chunk,
                           device=dev, features_to_host=False)
        for _ in range(2):
            hp2.run(*h)
        barrier()
        f0, f1 = ev(), ev()
        f0.record()
        for _ in range(K):
            out2 = hp2.run(*h)
            if world > 1:
                all_gather_blobs([out2["rois"].view(B, post, 5).to(dev, non_blocking=True),
                                  out2["scores"].view(B, post).to(dev, non_blocking=True),
                                  out2["counts"].to(dev, non_blocking=True)])
        f1.record()
        barrier()
        e2e_dev = (f0.elapsed_time(f1), hp2.h2d_bytes, hp2.d2h_bytes)
    clocks = sampler.summary() if sampler else None

    # max over ranks
    t = torch.tensor([ms_total, roi_ms, prop_ms, e2e[0] if e2e else 0.0,
                      e2e_dev[0] if e2e else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, roi_ms, prop_ms, e2e_ms, e2e_dev_ms = [float(v) for v in t.tolist()]
    if rank == 0:
        ms_step = ms_total / K
        value = world * B / (ms_step / 1e3)
        peak, peak_src = measured_peak()
        achieved = B * ROI_POOL_FWD_BYTES_PER_IMAGE / (roi_ms / 1e3) / 1e9
        # wssdl_roi_pool_fwd picks the band kernel for this workload (csrc/roi_pool.cu);
        # WSSDL_ROI_FWD_KERNEL=direct|tiled forces the other two
        kenv = os.environ.get("WSSDL_ROI_FWD_KERNEL", "")
        kname, ktmpl, nlaunch = {
            "d": ("roi_pool_fwd_kernel", "<4,CPU_TRUNC,128,2>", 2),
            "t": ("roi_pool_fwd_tiled_kernel", "<CPU_TRUNC>", 4),
        }.get(kenv[:1], ("roi_pool_fwd_band_kernel", "<CPU_TRUNC,argmax,linear>", 4))
        traffic, traffic_src = measured_traffic(kname, B)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": max(Wm, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, world),
            "roofline": {"kernel": kname + ktmpl,
                         "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "frac_of_nominal_8000": achieved / 8000.0,
                         "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": B * ROI_POOL_FWD_BYTES_PER_IMAGE,
                         "ms_per_launch": roi_ms},
            "kernels_ms_per_step": {"proposals_kernel": prop_ms, "roi_pool_fwd_kernel": roi_ms},
            # proposals_kernel + roi_pool_fwd kernel (+ roi_hist_kernel and roi_scatter_kernel
            # when a shared-memory kernel groups the RoIs by image; their few us are inside
            # ms_per_launch) per step
            "gpu_launches": nlaunch * K,
            "clocks": clocks,
            "rois_per_image": [int(counts.min()), int(counts.max())],
            "numa_node_rank0": numa,
        }
        if e2e:
            ems = e2e_ms / K
            line["e2e"] = {"value": world * B / (ems / 1e3), "unit": UNIT,
                           "h2d_bytes_per_step": e2e[1], "d2h_bytes_per_step": e2e[2],
                           "ms_per_step": ems,
                           "note": "pinned host inputs -> device -> all outputs (rois, scores, "
                                   "counts, pooled features, argmax) back to pinned host, "
                                   "chunked over 2 streams; bound by the PCIe D2H copy of the "
                                   "pooled features"}
            dms = e2e_dev_ms / K
            line["e2e_features_on_device"] = {
                "value": world * B / (dms / 1e3), "unit": UNIT, "h2d_bytes_per_step": e2e_dev[1],
                "d2h_bytes_per_step": e2e_dev[2], "ms_per_step": dms,
                "note": "same call and kernels; pooled features + argmax stay in HBM for the "
                        "next layer (the reference's TF graph does the same), RoIs/scores/counts "
                        "return to pinned host"}
        if world == 1 and not args.no_cpu_baseline:
            # the CPU arm runs in a fresh process (no CUDA context in the forked workers)
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                   "--warmup", "0"]
            if args.cpu_sample:
                cmd += ["--cpu-sample", str(args.cpu_sample)]
            env = dict(os.environ, RANK="0", WORLD_SIZE="1", CUDA_VISIBLE_DEVICES="")
            r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900,
                               preexec_fn=lambda: os.sched_setaffinity(0, all_cpus))  # all host cores
            try:
                line["cpu_baseline"] = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
            except Exception:
                line["cpu_baseline"] = {"error": (r.stderr or r.stdout)[-300:]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images-per-gpu", type=int, default=256)
    ap.add_argument("--e2e-chunk", type=int, default=32)
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
