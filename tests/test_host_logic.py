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
