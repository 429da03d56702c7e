"""Host-side logic (no GPU): anchors, config mirror, synthetic generators, sharding."""
import numpy as np

from wssdl_bus_b200 import synthetic as syn
from wssdl_bus_b200.fast_rcnn.config import cfg
from wssdl_bus_b200.rpn_msr.generate_anchors import generate_anchors, shifted_anchors


def test_generate_anchors_matches_reference_table(golden):
    assert np.array_equal(generate_anchors(), golden["anchors_table_minus_1"])


def test_generate_anchors_matches_oracle_for_other_scales(oracle_mod):
    for scales in ([8, 16, 32], [4, 8, 16, 32], [2, 4]):
        a = generate_anchors(scales=np.array(scales))
        b = oracle_mod.layers.generate_anchors(scales=np.array(scales))
        assert np.array_equal(a, b)
    a, A = oracle_mod.layers.shifted_anchors(5, 7, 16, (8, 16, 32))
    assert np.array_equal(shifted_anchors(5, 7, 16, generate_anchors()), a) and A == 9


def test_config_constants():
    assert cfg.TEST.RPN_PRE_NMS_TOP_N == 6000 and cfg.TEST.RPN_POST_NMS_TOP_N == 300
    assert cfg.TRAIN.RPN_PRE_NMS_TOP_N == 12000 and cfg['TRAIN'].RPN_POST_NMS_TOP_N == 2000
    assert cfg.TEST.RPN_NMS_THRESH == 0.7 and cfg.TEST.NMS == 0.3 and cfg.USE_GPU_NMS is False


def test_synthetic_scores_are_unique_and_seeded():
    cls, reg, info = syn.rpn_outputs(0, 2, 38, 50, 9)
    fg = cls[0, :, :, 9:].ravel()
    assert len(np.unique(fg)) == fg.size == 17100
    cls2, _, _ = syn.rpn_outputs(0, 2, 38, 50, 9)
    assert np.array_equal(cls, cls2)
    fm = syn.feature_map(1, 1, 38, 50, 512)
    assert 0.4 < (fm == 0).mean() < 0.6          # ties exist
    d = syn.dets(3, 500)
    assert len(np.unique(d[:, 4])) == 500


def test_image_sharding_is_a_partition():
    from wssdl_bus_b200.pipeline import shard_images
    for n, world in ((256, 8), (10, 4), (3, 8), (0, 2)):
        parts = [shard_images(n, r, world) for r in range(world)]
        allidx = np.concatenate(parts) if parts else np.zeros(0, int)
        assert sorted(allidx.tolist()) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_numa_binding_is_a_noop_without_a_visible_topology():
    """bind_to_gpu_numa_node must never raise or shrink the affinity when the GPU / sysfs
    topology is not visible (CPU-only containers, VMs with numa_node = -1)."""
    import os
    from wssdl_bus_b200.pipeline import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None or isinstance(bind_to_gpu_numa_node(0), int)
    assert os.sched_getaffinity(0) <= before and len(os.sched_getaffinity(0)) > 0


# ---------------------------------------------------------------- RoI-pool forward dispatch
# wssdl_roi_pool_fwd_plan is a host-only query (no CUDA call): the kernel a shape takes and the
# row-band geometry of the band / sorted-bins kernels (csrc/roi_pool.cu: choose_fwd / plan_band,
# csrc/roi_pool_bins.cu: plan_bins).
DIRECT, TILED, BAND, SORTED = 0, 1, 2, 3


def _plan(B, H, W, C, R, PH, PW, ws=True, force=0):
    import ctypes
    from wssdl_bus_b200 import _lib
    out = (ctypes.c_int * 10)()
    rc = _lib.lib().wssdl_roi_pool_fwd_plan(B, H, W, C, R, PH, PW, int(ws), force, out)
    assert rc == 0
    return dict(zip(("kernel", "NB", "Hb", "step", "nchunks", "RB", "smem", "scan", "slices",
                     "threads"), list(out)))


def test_fwd_kernel_choice_on_the_baseline_shapes():
    # C4 (the bench workload): sorted-bins kernel, two bands that own 19 rows each (27 resident
    # rows: equal shares), RoI lists from the workspace; the shortest bands (23 rows, owning 15 and
    # 23) through the tuning key
    p = _plan(256, 38, 50, 512, 256 * 300, 7, 7)
    assert (p["kernel"], p["NB"], p["Hb"], p["step"], p["scan"]) == (SORTED, 2, 27, 19, 0)
    from wssdl_bus_b200 import _lib
    prev = _lib.set_tuning("roi_fwd_balanced", 0)
    try:
        p = _plan(256, 38, 50, 512, 256 * 300, 7, 7)
        assert (p["kernel"], p["NB"], p["Hb"], p["step"]) == (SORTED, 2, 23, 15)
    finally:
        _lib.set_tuning("roi_fwd_balanced", prev)
    # short grids (a rank's 32 images at N = 8) keep the shortest bands
    p = _plan(32, 38, 50, 512, 32 * 300, 7, 7)
    assert (p["NB"], p["Hb"], p["step"]) == (2, 23, 15)
    # C1 / C2 (one image): one wave of the band kernel (a single launch), lists built in-kernel
    for R in (300, 128):
        p = _plan(1, 38, 50, 512, R, 7, 7)
        assert p["kernel"] == BAND and p["scan"] == 1
        assert (512 // 32) * p["NB"] * p["nchunks"] <= 148
    # 16 / 32 images (the per-rank batches of the strong-scaling run): sorted bins as well
    assert _plan(16, 38, 50, 512, 16 * 300, 7, 7)["kernel"] == SORTED
    assert _plan(32, 38, 50, 512, 32 * 300, 7, 7)["kernel"] == SORTED
    # C3 (14x14 bins on 1024 channels): direct kernel
    assert _plan(16, 38, 50, 1024, 4800, 14, 14)["kernel"] == DIRECT
    # C % 32 != 0: the 16-channel tiled kernel takes single images; C % 16 != 0: direct
    assert _plan(1, 38, 50, 48, 300, 7, 7)["kernel"] == TILED
    assert _plan(1, 38, 50, 20, 300, 7, 7)["kernel"] == DIRECT
    # without a workspace: no room for the bin records (band kernel while the lists can be
    # built in-kernel), > 4096 RoIs: nothing to group them with either
    assert _plan(1, 38, 50, 512, 300, 7, 7, ws=False)["kernel"] == BAND
    assert _plan(16, 38, 50, 512, 4000, 7, 7, ws=False)["kernel"] == DIRECT
    assert _plan(16, 38, 50, 512, 4000, 7, 7, ws=False, force=3)["kernel"] == BAND
    assert _plan(256, 38, 50, 512, 256 * 300, 7, 7, ws=False)["kernel"] == DIRECT
    # forcing
    assert _plan(256, 38, 50, 512, 256 * 300, 7, 7, force=1)["kernel"] == DIRECT
    assert _plan(256, 38, 50, 512, 256 * 300, 7, 7, force=2)["kernel"] == TILED
    assert _plan(256, 38, 50, 512, 256 * 300, 7, 7, force=3)["kernel"] == BAND
    assert _plan(16, 38, 50, 1024, 4800, 14, 14, force=3)["kernel"] == BAND
    assert _plan(16, 38, 50, 1024, 4800, 7, 7, force=4)["kernel"] == SORTED


def test_sorted_bins_plan_invariants():
    """Seeded sweep of shapes: whenever the sorted-bins kernel is available its bands cover the
    map and overlap by at least the tallest bin of a RoI inside the map, a band's cells fit the
    11-bit first-cell field of a bin record, and the band fits the shared memory of a CTA (one or
    two CTAs per SM)."""
    from wssdl_bus_b200 import _lib
    rng = np.random.default_rng(11)
    seen = seen_multi = 0
    try:
        for threads in (0, 512):
            _lib.set_tuning("roi_fwd_threads", threads)
            for _ in range(400):
                B = int(rng.integers(1, 40))
                H, W = int(rng.integers(1, 120)), int(rng.integers(1, 120))
                C = int(rng.choice([32, 64, 128, 256, 512, 1024]))
                R = int(rng.choice([1, 17, 300, 4096, 4097, 20000]))
                PH, PW = int(rng.integers(1, 10)), int(rng.integers(1, 10))
                p = _plan(B, H, W, C, R, PH, PW, force=4)
                if p["kernel"] != SORTED:
                    continue
                seen += 1
                NB, Hb, step = p["NB"], p["Hb"], p["step"]
                assert 1 <= NB <= 4 and step >= 1 and 1 <= Hb <= H
                assert (NB - 1) * step + Hb >= H
                assert Hb * W <= 2047 and PH * PW <= 64
                if NB > 1:
                    seen_multi += 1
                    assert Hb - step >= -(-(H + 1) // PH) + 2
                assert p["threads"] == (512 if threads == 512 else 1024)
                assert 1 <= p["RB"] <= 1024 and p["nchunks"] >= 1 and 1 <= p["slices"] <= C // 32
                cap = 113 * 1024 if threads == 512 else 227 * 1024 - 1024
                assert Hb * W * 128 + 128 <= p["smem"] <= cap
                assert p["scan"] == (1 if R <= 4096 else 0)
    finally:
        _lib.set_tuning("roi_fwd_threads", 0)
    assert seen > 100 and seen_multi > 10


def test_band_geometry_invariants():
    """Seeded sweep of shapes: whenever the band kernel is available its bands cover the map,
    overlap by at least the tallest bin of a RoI inside the map (+1 for GPU_CEIL edges), and
    the CTA's shared memory (map + bin-edge tables + lists) fits the 227 KB carve-out."""
    rng = np.random.default_rng(7)
    seen_multi = 0
    for _ in range(400):
        B = int(rng.integers(1, 40))
        H, W = int(rng.integers(1, 120)), int(rng.integers(1, 120))
        C = int(rng.choice([32, 64, 128, 256, 512, 1024]))
        R = int(rng.choice([1, 17, 300, 4096, 4097, 20000]))
        PH, PW = int(rng.integers(1, 15)), int(rng.integers(1, 15))
        p = _plan(B, H, W, C, R, PH, PW, force=3)
        if p["kernel"] != BAND:
            continue
        NB, Hb, step = p["NB"], p["Hb"], p["step"]
        assert NB >= 1 and step >= 1 and 1 <= Hb <= H
        assert (NB - 1) * step + Hb >= H                      # the last band reaches the last row
        if NB > 1:
            seen_multi += 1
            ov = Hb - step
            assert ov >= -(-(H + 1) // PH) + 2                # ceil((H+1)/PH) + 2
            assert (NB - 2) * step + Hb < H + step            # no band is superfluous
        assert Hb * W * 128 + 128 <= p["smem"] <= 227 * 1024 - 1024
        assert p["RB"] >= min(R, 16) and p["nchunks"] >= 1
        assert p["scan"] == (1 if R <= 4096 else 0)
    assert seen_multi > 20


def test_generate_anchors_equals_the_reference_function_output(oracle_mod):
    """Host generate_anchors (product) and the oracle restatement against what the reference's
    own rpn_msr/generate_anchors.py returns when executed (tests/golden/make_layers_golden.py),
    for the default, the 4-scale layer default and a non-standard base size / ratio set."""
    import os
    from wssdl_bus_b200.rpn_msr.generate_anchors import generate_anchors
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_layers_golden.npz"))
    cases = (("anchors_default", {}),
             ("anchors_4_8_16_32", dict(scales=np.array([4, 8, 16, 32]))),
             ("anchors_ratios_1_2_base8", dict(base_size=8, ratios=[1, 2], scales=np.array([2, 16]))))
    for key, kw in cases:
        assert np.array_equal(generate_anchors(**kw), g[key]), key
        assert np.array_equal(oracle_mod.layers.generate_anchors(**kw), g[key]), key
