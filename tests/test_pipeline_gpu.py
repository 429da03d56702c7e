"""proposals -> roi_pool composition (C1/C4 shapes) and the host-buffer pipeline."""
import numpy as np
import pytest
import torch

from wssdl_bus_b200 import synthetic as syn
from wssdl_bus_b200.pipeline import HostPipeline, HotPath

pytestmark = pytest.mark.gpu


def _inputs(seed, B):
    c = syn.C1
    feat = syn.feature_map(seed, B, c["H"], c["W"], c["C"])
    cls, reg, info = syn.rpn_outputs(seed + 1, B, c["H"], c["W"], c["A"])
    return feat, cls, reg, info


def test_hot_path_matches_oracle_chain(oracle_mod):
    B = 3
    feat, cls, reg, info = _inputs(700, B)
    hot = HotPath()
    p = hot.run(torch.from_numpy(feat).cuda(), cls, reg, info)
    rois = p["rois"].cpu().numpy()
    # the pooled output must be the oracle's pooling of the device RoIs, bit for bit
    wt, wa = oracle_mod.clib.roi_pool_fwd(feat, rois, 7, 7, 1 / 16.)
    assert np.array_equal(p["top"].cpu().numpy(), wt)
    assert np.array_equal(p["argmax"].cpu().numpy(), wa)
    assert p["counts"].cpu().numpy().tolist() == [300] * B
    boxes, scores, cnt = hot.detections(p)
    assert boxes.shape == (B, 300, 5) and scores.shape == (B, 300) and cnt.shape == (B,)
    assert boxes.data_ptr() == p["rois"].data_ptr()                   # views, no copy kernel
    assert bool(torch.all(scores[:, :-1] > scores[:, 1:]))            # descending scores


def test_host_pipeline_equals_device_path():
    B = 5
    feat, cls, reg, info = _inputs(710, B)
    hot = HotPath()
    p = hot.run(torch.from_numpy(feat).cuda(), cls, reg, info)
    hp = HostPipeline(hot, B, 38, 50, 512, 9, chunk=2)
    pin = [torch.from_numpy(x).pin_memory() for x in (feat, cls, reg, info)]
    out = hp.run(*pin)
    assert torch.equal(out["rois"], p["rois"].cpu())
    assert torch.equal(out["top"], p["top"].cpu())
    assert torch.equal(out["argmax"], p["argmax"].cpu())
    assert torch.equal(out["counts"], p["counts"].cpu())
    assert hp.h2d_bytes == sum(x.nbytes for x in (feat, cls, reg, info))
