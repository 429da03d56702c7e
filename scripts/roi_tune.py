#!/usr/bin/env python
"""Sorted-bins RoI-pool forward: slices-per-CTA x RoI-chunks sweep on the bench shapes
(CUDA events, median of 15).  python scripts/roi_tune.py [--out file]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wssdl_bus_b200 import _lib, ops, synthetic as syn  # noqa: E402
from scripts.microbench import PEAK, realistic_rois, timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--images", default="256,32,16,1")
    ap.add_argument("--slices", default="0,1,2,4")
    ap.add_argument("--chunks", default="0,1,2,4")
    ap.add_argument("--kernels", default="sorted")
    ap.add_argument("--threads", default="0")
    args = ap.parse_args()
    out = open(args.out, "a") if args.out else None
    for B in [int(v) for v in args.images.split(",")]:
        rois = realistic_rois(B)
        x = torch.from_numpy(syn.feature_map(1, B, 38, 50, 512)).cuda()
        nbytes = B * 38 * 50 * 512 * 4 + rois.shape[0] * (20 + 49 * 512 * 8)
        for kern in args.kernels.split(","):
            _lib.set_tuning("roi_fwd_kernel", kern)
            for sl in [int(v) for v in args.slices.split(",")]:
                for ch in [int(v) for v in args.chunks.split(",")]:
                    if kern != "sorted" and (sl or ch):
                        continue
                    for th in [int(v) for v in args.threads.split(",")]:
                        if kern != "sorted" and th:
                            continue
                        _lib.set_tuning("roi_fwd_slices", sl)
                        _lib.set_tuning("roi_fwd_chunks", ch)
                        _lib.set_tuning("roi_fwd_threads", th)
                        med, best = timeit(lambda: ops.roi_pool_forward(x, rois, 7, 7, 1 / 16.), iters=15,
                                           flush=nbytes < (1 << 30))
                        line = json.dumps(dict(B=B, kernel=kern, slices=sl, chunks=ch, threads=th, ms=med,
                                               ms_min=best, frac_measured=nbytes / med / 1e6 / PEAK))
                        print(line, flush=True)
                        if out:
                            out.write(line + "\n")
    _lib.set_tuning("roi_fwd_kernel", "auto")
    for k in ("roi_fwd_slices", "roi_fwd_chunks", "roi_fwd_threads"):
        _lib.set_tuning(k, 0)


if __name__ == "__main__":
    main()
