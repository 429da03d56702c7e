"""Experiment: the C4 RoI-pool stage (grouped entry) in both bin modes, 256 images (run on a B200)."""
import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from wssdl_bus_b200 import ops, synthetic as syn
from wssdl_bus_b200.pipeline import HotPath
B=256
cls, reg, info = syn.rpn_outputs(0, B, 38, 50, 9)
hot = HotPath()
rois = ops.proposals(cls, reg, info, hot.base, 16, hot.pre, hot.post, hot.thresh, hot.min_size)["rois"]
x = torch.from_numpy(syn.feature_map(1, B, 38, 50, 512)).cuda()
for mode in ("cpu","gpu"):
    for _ in range(3): ops.roi_pool_forward_grouped(x, rois, 300, 7, 7, 1/16., mode)
    torch.cuda.synchronize()
    ts=[]
    for _ in range(10):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); ops.roi_pool_forward_grouped(x, rois, 300, 7, 7, 1/16., mode); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms=float(np.median(ts)); print(mode, ms, 16411750400/ms/1e6/6550.7)
