"""NMS parity on the GPU: keep lists bit-exact vs cpu_nms (the reference's own compiled
Cython when oracle/_ref is present, its pinned C restatement otherwise)."""
import numpy as np
import pytest
import torch

from wssdl_bus_b200 import ops, synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["cta", "cluster"])
def sweep_variant(request, tuning):
    """Every test runs twice: the single-CTA sweep and the sweep over a thread-block cluster
    of 8 CTAs (csrc/nms.cu; the default picks the cluster from N = 65536)."""
    tuning("nms_sweep_cluster", 1 if request.param == "cluster" else 0)
    return request.param


def _oracle_nms(oracle_mod, d, t, variant=0):
    if variant == 0 and oracle_mod.ref.available() and d.shape[0] <= 3000:
        return oracle_mod.ref.cpu_nms(d, t)
    return oracle_mod.clib.nms(d, t, variant=variant)


@pytest.mark.parametrize("key", ["nms_uni_257", "nms_uni_1000", "nms_clu_257", "nms_clu_1000", "nms_fork"])
def test_golden_keep_lists(golden, key):
    d = golden[key + "_dets"]
    for t in (0.3, 0.5, 0.7):
        assert ops.nms(d, t) == golden[key + "_keep_%02d" % int(t * 10)].tolist()
        assert ops.nms(torch.from_numpy(d).cuda(), t) == golden[key + "_keep_%02d" % int(t * 10)].tolist()
    if key + "_keepnew_05" in golden:
        assert ops.nms(d, 0.5, ops.NMS_GE_F64 | ops.NMS_CONTAIN) == golden[key + "_keepnew_05"].tolist()


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 4095, 4096, 4097, 6000, 12000])
@pytest.mark.parametrize("clustered", [False, True])
def test_sizes_vs_oracle(oracle_mod, n, clustered):
    d = syn.dets(100 + n, n, clustered=clustered)
    for t in (0.3, 0.5, 0.7):
        assert ops.nms(d, t) == _oracle_nms(oracle_mod, d, t)


_C5_ORACLE = {}


def _c5_oracle(oracle_mod, n, clustered, t):
    """cpu_nms semantics at C5 sizes: the pinned C restatement (seconds per call at 100 k);
    cached so that the two sweep variants share one oracle run."""
    key = (n, clustered, t)
    if key not in _C5_ORACLE:
        _C5_ORACLE[key] = oracle_mod.clib.nms(syn.dets(300 + n, n, clustered=clustered), t)
    return _C5_ORACLE[key]


@pytest.mark.parametrize("n,clustered,thresholds", [(50000, False, (0.3, 0.5, 0.7)),
                                                    (100000, True, (0.3, 0.5, 0.7)),
                                                    (100000, False, (0.7,))])
def test_c5_50k_100k_keep_lists_equal_the_oracle(oracle_mod, n, clustered, thresholds, sweep_variant):
    """BASELINE config 5 at its upper sizes, both sweeps (the cluster sweep is the default route
    from N = 65536): keep lists bit-exact against cpu_nms semantics."""
    d = syn.dets(300 + n, n, clustered=clustered)
    for t in thresholds:
        assert ops.nms(d, t) == _c5_oracle(oracle_mod, n, clustered, t), (n, clustered, t, sweep_variant)


def test_c5_default_route_is_the_cluster_sweep_from_65536(oracle_mod, tuning):
    """Without a forced variant: N = 100 k takes the cluster sweep, and gives the oracle's list."""
    tuning("nms_sweep_cluster", -1)
    d = syn.dets(300 + 100000, 100000, clustered=True)
    assert ops.nms(d, 0.5) == _c5_oracle(oracle_mod, 100000, True, 0.5)


def test_c5_sweep_large_properties(oracle_mod):
    """N = 20k / 50k / 100k: oracle at 20k (C restatement, seconds); beyond that properties:
    keep sorted by score, no kept pair above the threshold, every dropped box has a kept
    suppressor with a higher score (checked on a sample), idempotence."""
    d = syn.dets(200, 20000, clustered=True)
    assert ops.nms(d, 0.5) == oracle_mod.clib.nms(d, 0.5)
    for n in (50000, 100000):
        d = syn.dets(201 + n, n)
        keep = ops.nms(d, 0.7)
        k = np.asarray(keep)
        assert np.all(np.diff(d[k, 4]) < 0)
        assert len(set(keep)) == len(keep)
        dk = d[k]
        assert ops.nms(dk, 0.7) == list(range(len(k)))                 # idempotent
        rng = np.random.default_rng(n)
        dropped = np.setdiff1d(np.arange(n), k)
        for j in rng.choice(dropped, size=min(50, len(dropped)), replace=False):
            higher = dk[dk[:, 4] > d[j, 4]]
            iou = oracle_mod.clib.bbox_overlaps(higher[:, :4].astype(np.float64),
                                                d[j:j + 1, :4].astype(np.float64))
            assert iou.max() >= 0.7 - 1e-6


def test_modes_gt_and_contain(oracle_mod, golden):
    d = golden["nms_fork_dets"]
    # '>' in fp32 (gpu_nms / py_cpu_nms rules): iou == 0.7f is not > 0.7f, iou == 0.3f not > 0.3f
    assert ops.nms(d, 0.7, ops.NMS_GT_F32) == [0, 1, 2, 3]
    # at 0.3 box 1 (iou 0.7 with box 0) goes; box 3 (iou == 0.3f, not > 0.3f) stays -- cpu_nms
    # rules drop it (golden nms_fork_keep_03 == [0, 2])
    assert ops.nms(d, 0.3, ops.NMS_GT_F32) == [0, 2, 3]
    d = syn.dets(300, 2500, clustered=True)
    # differential twin: numpy restatement of py_cpu_nms.py:10-38 (fp32, '<=' keeps)
    def py_nms(dets, thresh):
        x1, y1, x2, y2, s = dets.T
        areas = (x2 - x1 + 1) * (y2 - y1 + 1)
        order = s.argsort()[::-1]
        keep = []
        while order.size > 0:
            i = order[0]
            keep.append(int(i))
            xx1 = np.maximum(x1[i], x1[order[1:]]); yy1 = np.maximum(y1[i], y1[order[1:]])
            xx2 = np.minimum(x2[i], x2[order[1:]]); yy2 = np.minimum(y2[i], y2[order[1:]])
            w = np.maximum(np.float32(0.0), xx2 - xx1 + 1); h = np.maximum(np.float32(0.0), yy2 - yy1 + 1)
            inter = w * h
            ovr = inter / (areas[i] + areas[order[1:]] - inter)
            order = order[np.where(ovr <= np.float32(thresh))[0] + 1]
        return keep
    assert ops.nms(d, 0.5, ops.NMS_GT_F32) == py_nms(d, 0.5)
    # the reference's own CUDA NMS (nms/nms_kernel.cu: devIoU, nms_kernel, _nms) built for the
    # host (oracle/build_ref.py:build_cuda_nms) behind the gpu_nms.pyx glue
    if oracle_mod.ref.gpu_nms_available():
        assert ops.nms(d, 0.5, ops.NMS_GT_F32) == [int(k) for k in oracle_mod.ref.gpu_nms(d, 0.5)]
        fork = golden["nms_fork_dets"]
        assert [int(k) for k in oracle_mod.ref.gpu_nms(fork, 0.3)] == ops.nms(fork, 0.3, ops.NMS_GT_F32)
    for t in (0.3, 0.7):
        assert ops.nms(d, t, ops.NMS_GE_F64 | ops.NMS_CONTAIN) == oracle_mod.clib.nms(d, t, variant=1)


def test_max_keep_truncates_like_slicing(oracle_mod):
    d = syn.dets(301, 6000)
    full = oracle_mod.clib.nms(d, 0.7)
    for m in (1, 63, 64, 300, 2000):
        assert ops.nms(d, 0.7, max_keep=m) == full[:m]
        keep, num, _ = ops.nms_device(torch.from_numpy(d).cuda(), 0.7, max_keep=m)
        assert keep[:int(num.item())].tolist() == full[:m]


def test_empty_ties_and_zero_union(oracle_mod):
    assert ops.nms(np.zeros((0, 5), np.float32), 0.5) == []
    from wssdl_bus_b200.fast_rcnn.nms_wrapper import nms as wrapped
    assert wrapped(np.zeros((0, 5), np.float32), 0.5) == []
    # ties: documented order (score desc, index desc) == argsort(kind='stable')[::-1]
    d = syn.dets(302, 500)
    d[:, 4] = np.round(d[:, 4] * 8) / 8
    order = d[:, 4].argsort(kind="stable")[::-1]
    assert ops.nms(d, 0.5) == oracle_mod.clib.nms(d, 0.5, order=order)
    z = np.array([[5, 5, 4, 4, 0.9], [7, 7, 6, 6, 0.8]], np.float32)
    with pytest.raises(ZeroDivisionError):
        ops.nms(z, 0.5)


def test_dropin_modules_and_host_abi(oracle_mod):
    from wssdl_bus_b200.nms.cpu_nms import cpu_nms
    from wssdl_bus_b200.nms.gpu_nms import gpu_nms
    from wssdl_bus_b200.utils.cython_nms import nms as cy_nms
    d = syn.dets(303, 1500)
    want = _oracle_nms(oracle_mod, d, 0.7)
    assert cpu_nms(d, 0.7) == want and cy_nms(d, 0.7) == want
    assert isinstance(gpu_nms(d, 0.7), list)
    # `_nms` twin: sorted input, positions in the sorted array (gpu_nms.pyx:25-31)
    order = d[:, 4].argsort()[::-1]
    keep_pos = ops.gpu_nms_sorted_host(d[order], 0.7)
    assert list(order[keep_pos]) == ops.nms(d, 0.7, ops.NMS_GT_F32)
    # dets with extra columns (all_dets has 6, test_bus.py:378)
    d6 = np.hstack([d, np.ones((len(d), 1), np.float32)])
    assert ops.nms(d6, 0.7) == want
