"""RoI pooling parity on the GPU, through the C ABI (ctypes) behind wssdl_bus_b200.ops.
Forward out+argmax: bit-exact vs the CPU restatement of roi_pooling_op.cc (both bin modes).
Backward: deterministic gather bit-exact; atomic scatter within 1e-5 relative."""
import numpy as np
import pytest
import torch

from wssdl_bus_b200 import ops, synthetic as syn

pytestmark = pytest.mark.gpu

# tolerance for the atomics-accumulated gradient (north star: 1e-5 relative)
RTOL, ATOL = 1e-5, 1e-5


@pytest.fixture(params=["direct", "tiled", "band", "sorted"])
def fwd_kernel(request, tuning):
    """Forces one of the four forward kernels (roi_pool.cu: direct = one CTA per output row
    reading L2; tiled = shared-memory resident 16-channel slice of the whole map; band =
    32-channel slice of overlapping row bands, one thread per column of bins; roi_pool_bins.cu:
    sorted = same staging by TMA, bins counting-sorted by size class; C % 128 == 0 takes the
    linear-index variant of band / sorted).
    Shapes a shared-memory kernel does not take (C % 16 / 32 != 0, misaligned pointers) fall
    through to the direct one."""
    tuning("roi_fwd_kernel", request.param)
    return request.param


def _fwd_both(oracle_mod, bottom, rois, PH, PW, scale, mode):
    want_top, want_arg = oracle_mod.clib.roi_pool_fwd(bottom, rois, PH, PW, scale,
                                                      bin_mode=0 if mode == "cpu" else 1)
    top, arg = ops.roi_pool_forward(bottom, rois, PH, PW, scale, bin_mode=mode)
    return want_top, want_arg, top.cpu().numpy(), arg.cpu().numpy()


@pytest.mark.parametrize("mode", ["cpu", "gpu"])
def test_fwd_c1_shapes_bit_exact(oracle_mod, mode, fwd_kernel):
    c = syn.C1
    bottom = syn.feature_map(0, c["B"], c["H"], c["W"], c["C"])
    rois = syn.rois_for_pool(1, 300)
    wt, wa, t, a = _fwd_both(oracle_mod, bottom, rois, 7, 7, c["scale"], mode)
    assert np.array_equal(t, wt) and np.array_equal(a, wa)


@pytest.mark.parametrize("C", [64, 128])
@pytest.mark.parametrize("mode", ["cpu", "gpu"])
def test_fwd_adversarial_rois_bit_exact(oracle_mod, mode, C, fwd_kernel):
    B, H, W = 2, 38, 50
    bottom = syn.feature_map(2, B, H, W, C)
    bottom[0, 3:9, 4:11] = -np.inf                     # cells that can never win
    bottom[1, 0, 0, :] = np.nan                        # NaN never wins either
    rois = np.concatenate([syn.adversarial_rois(B, W, H), syn.rois_for_pool(3, 64, B)])
    for PH, PW in ((7, 7), (14, 14), (6, 6), (1, 1), (3, 5)):
        wt, wa, t, a = _fwd_both(oracle_mod, bottom, rois, PH, PW, 1 / 16., mode)
        assert np.array_equal(a, wa)
        assert np.array_equal(t, wt, equal_nan=True)


@pytest.mark.parametrize("C", [3, 5, 32, 100, 1024])
def test_fwd_channel_counts_and_scalar_path(oracle_mod, C, fwd_kernel):
    # C=3 is the reference's own smoke shape (roi_pooling_op_test.py: 32x100x100x3, scale 1/3)
    B, H, W = 2, 20, 24
    bottom = syn.feature_map(4, B, H, W, C)
    rois = syn.rois_for_pool(5, 17, B, im_w=W * 3, im_h=H * 3)
    wt, wa, t, a = _fwd_both(oracle_mod, bottom, rois, 6, 6, 1 / 3., "cpu")
    assert np.array_equal(t, wt) and np.array_equal(a, wa)


def test_fwd_misaligned_pointer_uses_scalar_kernel(oracle_mod):
    B, H, W, C = 1, 10, 12, 8
    bottom = syn.feature_map(6, B, H, W, C)
    rois = syn.rois_for_pool(7, 9, B, im_w=W * 16, im_h=H * 16)
    buf = torch.zeros(bottom.size + 1, dtype=torch.float32, device="cuda")
    view = buf[1:].view(B, H, W, C)                    # 4-byte aligned only
    view.copy_(torch.from_numpy(bottom))
    top, arg = ops.roi_pool_forward(view, rois, 7, 7, 1 / 16.)
    wt, wa = oracle_mod.clib.roi_pool_fwd(bottom, rois, 7, 7, 1 / 16.)
    assert np.array_equal(top.cpu().numpy(), wt) and np.array_equal(arg.cpu().numpy(), wa)


@pytest.mark.parametrize("kern,C", [("tiled", 32), ("band", 32), ("band", 128), ("sorted", 32), ("sorted", 128)])
@pytest.mark.parametrize("mode", ["cpu", "gpu"])
def test_fwd_tiled_sorted_lists_many_images(oracle_mod, mode, kern, C, tuning):
    """R > 4096: the shared-memory kernels take their per-image RoI lists from the counting-sort
    pre-pass in the workspace.  Batch indices are shuffled and some are out of range."""
    tuning("roi_fwd_kernel", kern)
    B, H, W = 24, 38, 50
    bottom = syn.feature_map(30, B, H, W, C)
    rois = np.concatenate([syn.rois_for_pool(31, 5000, B), syn.adversarial_rois(B, W, H)])
    rng = np.random.default_rng(32)
    rois = rois[rng.permutation(rois.shape[0])]
    rois[::97, 0] = B + 3            # no such image: (0, -1)
    rois[5::211, 0] = -2
    bad = (rois[:, 0] < 0) | (rois[:, 0] >= B)
    safe = rois.copy()
    safe[bad, 0] = 0                 # the oracle would read out of bounds for those rows
    wt, wa = oracle_mod.clib.roi_pool_fwd(bottom, safe, 7, 7, 1 / 16., bin_mode=0 if mode == "cpu" else 1)
    top, arg = ops.roi_pool_forward(bottom, rois, 7, 7, 1 / 16., bin_mode=mode)
    t, a = top.cpu().numpy(), arg.cpu().numpy()
    assert np.array_equal(a[~bad], wa[~bad]) and np.array_equal(t[~bad], wt[~bad])
    assert not t[bad].any() and (a[bad] == -1).all()
    # same answer from the direct kernel and without a workspace
    tuning("roi_fwd_kernel", "direct")
    t2, a2 = ops.roi_pool_forward(bottom, rois, 7, 7, 1 / 16., bin_mode=mode)
    assert np.array_equal(t2.cpu().numpy(), t) and np.array_equal(a2.cpu().numpy(), a)


@pytest.mark.parametrize("kern,C", [("tiled", 48), ("band", 64), ("band", 128), ("sorted", 64), ("sorted", 128)])
@pytest.mark.parametrize("B,R", [(1, 300), (2, 700), (3, 40), (5, 4096), (1, 1)])
def test_fwd_tiled_chunked_scan_lists(oracle_mod, B, R, kern, C, tuning):
    """R <= 4096: RoI lists are built inside each CTA; few images split their RoIs over
    several CTAs (chunk = RoI index mod nchunks).  Also without argmax."""
    tuning("roi_fwd_kernel", kern)
    H, W = 38, 50
    bottom = syn.feature_map(33, B, H, W, C)
    rois = syn.rois_for_pool(34, R, B)
    wt, wa, t, a = _fwd_both(oracle_mod, bottom, rois, 7, 7, 1 / 16., "cpu")
    assert np.array_equal(t, wt) and np.array_equal(a, wa)
    t3, a3 = ops.roi_pool_forward(bottom, rois, 7, 7, 1 / 16., need_argmax=False)
    assert a3 is None and np.array_equal(t3.cpu().numpy(), wt)


def test_fwd_workspace_query_and_null_workspace(oracle_mod):
    """The C ABI accepts a NULL workspace (direct kernel for big batches) and reports the
    scratch size it wants."""
    import ctypes
    from wssdl_bus_b200 import _lib
    L = _lib.lib()
    assert L.wssdl_roi_pool_fwd_workspace_bytes(256, 76800, 7, 7) >= 4 * (256 + 2 + 76800) + 8 * 76800 * 49
    B, H, W, C, R = 4, 20, 24, 16, 5000
    bottom = syn.feature_map(35, B, H, W, C)
    rois = syn.rois_for_pool(36, R, B, im_w=W * 16, im_h=H * 16)
    x = torch.from_numpy(bottom).cuda()
    r = torch.from_numpy(rois).cuda()
    top = torch.empty((R, 7, 7, C), device="cuda")
    arg = torch.empty((R, 7, 7, C), device="cuda", dtype=torch.int32)
    vp = ctypes.c_void_p
    rc = L.wssdl_roi_pool_fwd(vp(x.data_ptr()), vp(r.data_ptr()), B, H, W, C, R, 7, 7, 1 / 16., 0,
                              vp(top.data_ptr()), vp(arg.data_ptr()), None, 0,
                              vp(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    wt, wa = oracle_mod.clib.roi_pool_fwd(bottom, rois, 7, 7, 1 / 16.)
    assert np.array_equal(top.cpu().numpy(), wt) and np.array_equal(arg.cpu().numpy(), wa)


def test_fwd_empty_and_errors():
    bottom = torch.zeros((1, 4, 4, 8), device="cuda")
    top, arg = ops.roi_pool_forward(bottom, np.zeros((0, 5), np.float32), 7, 7, 1 / 16.)
    assert top.shape == (0, 7, 7, 8) and arg.shape == (0, 7, 7, 8)
    with pytest.raises(ValueError, match="4-dimensional"):
        ops.roi_pool_forward(bottom[0], np.zeros((1, 5), np.float32), 7, 7, 1.0)
    with pytest.raises(ValueError, match="2-dimensional"):
        ops.roi_pool_forward(bottom, np.zeros((5,), np.float32), 7, 7, 1.0)
    with pytest.raises(ValueError, match="pooled_height"):
        ops.roi_pool_forward(bottom, np.zeros((1, 5), np.float32), -1, 7, 1.0)
    # batch index outside [0,B): documented (0, -1) instead of the reference's OOB read
    top, arg = ops.roi_pool_forward(bottom + 1, np.array([[3, 0, 0, 40, 40]], np.float32), 2, 2, 1 / 16.)
    assert float(top.abs().sum()) == 0 and int((arg != -1).sum()) == 0


@pytest.mark.parametrize("mode", ["cpu", "gpu"])
def test_bwd_gather_bit_exact_and_atomic_close(oracle_mod, mode):
    B, H, W, C, PH, PW = 2, 38, 50, 64, 7, 7
    bottom = syn.feature_map(8, B, H, W, C)
    rois = np.concatenate([syn.rois_for_pool(9, 128, B), syn.adversarial_rois(B, W, H)])
    rng = np.random.default_rng(10)
    top, arg = ops.roi_pool_forward(bottom, rois, PH, PW, 1 / 16., bin_mode=mode)
    g = rng.standard_normal(tuple(top.shape)).astype(np.float32)
    want = oracle_mod.clib.roi_pool_bwd(g, arg.cpu().numpy(), rois, bottom.shape, 1 / 16.)
    det = ops.roi_pool_backward(bottom.shape, rois, arg, g, PH, PW, 1 / 16., deterministic=True)
    assert np.array_equal(det.cpu().numpy(), want)
    atm = ops.roi_pool_backward(bottom.shape, rois, arg, g, PH, PW, 1 / 16., deterministic=False)
    np.testing.assert_allclose(atm.cpu().numpy(), want, rtol=RTOL, atol=ATOL)


def test_bwd_arbitrary_argmax_equals_reference_gather(oracle_mod):
    """The reference's gradient is a gather with in-RoI and feasible-bin tests; a scatter
    must reproduce it even for argmax tensors the forward would never produce."""
    B, H, W, C, PH, PW = 1, 12, 14, 8, 4, 4
    rois = syn.rois_for_pool(11, 20, B, im_w=W * 16, im_h=H * 16)
    rng = np.random.default_rng(12)
    arg = rng.integers(-1, H * W * C, size=(20, PH, PW, C)).astype(np.int32)
    g = rng.standard_normal(arg.shape).astype(np.float32)
    want = oracle_mod.clib.roi_pool_bwd(g, arg, rois, (B, H, W, C), 1 / 16., literal=True)
    for det in (True, False):
        got = ops.roi_pool_backward((B, H, W, C), rois, arg, g, PH, PW, 1 / 16., deterministic=det)
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=RTOL, atol=ATOL)


def test_c2_train_shapes_fwd_bwd(oracle_mod):
    c = syn.C2
    bottom = syn.feature_map(13, 1, c["H"], c["W"], c["C"])
    rois = syn.rois_for_pool(14, c["sampled"])
    rng = np.random.default_rng(15)
    x = torch.from_numpy(bottom).cuda().requires_grad_(True)
    top, arg = ops.roi_pool(x, rois, 7, 7, c["scale"])
    wt, wa = oracle_mod.clib.roi_pool_fwd(bottom, rois, 7, 7, c["scale"])
    assert np.array_equal(top.detach().cpu().numpy(), wt) and np.array_equal(arg.cpu().numpy(), wa)
    g = rng.standard_normal(wt.shape).astype(np.float32)
    top.backward(torch.from_numpy(g).cuda())
    want = oracle_mod.clib.roi_pool_bwd(g, wa, rois, bottom.shape, c["scale"])
    np.testing.assert_allclose(x.grad.cpu().numpy(), want, rtol=RTOL, atol=ATOL)
    # drop-in names
    from wssdl_bus_b200.roi_pooling_layer import roi_pooling_op
    gg = roi_pooling_op.roi_pool_grad(x.detach(), rois, arg, g, 7, 7, c["scale"], deterministic=True)
    assert np.array_equal(gg.cpu().numpy(), want)


def test_c3_resnet_shapes_properties(oracle_mod):
    """C3 at full size (16x38x50x1024, 4800 RoIs, 14x14): oracle on a slice, size-independent
    properties on the whole: argmax inside its bin's image, top == bottom[argmax], gradient
    mass conservation."""
    c = syn.C3
    B, H, W, C = c["B"], c["H"], c["W"], c["C"]
    bottom = syn.feature_map(16, B, H, W, C)
    rois = syn.rois_for_pool(17, B * c["rois_per_image"], B)
    x = torch.from_numpy(bottom).cuda()
    top, arg = ops.roi_pool_forward(x, rois, 14, 14, c["scale"])
    sl = slice(0, 64)
    wt, wa = oracle_mod.clib.roi_pool_fwd(bottom, rois[sl], 14, 14, c["scale"])
    assert np.array_equal(top[sl].cpu().numpy(), wt) and np.array_equal(arg[sl].cpu().numpy(), wa)
    # top == bottom[b, argmax] wherever argmax >= 0, 0 elsewhere
    bidx = torch.from_numpy(rois[:, 0]).cuda().long()
    flat = x.reshape(B, -1)
    a = arg.reshape(arg.shape[0], -1).long()
    gathered = flat[bidx[:, None].expand_as(a), a.clamp(min=0)]
    t = top.reshape(top.shape[0], -1)
    assert bool(torch.all(torch.where(a >= 0, gathered, torch.zeros_like(t)) == t))
    # channel of the argmax is the output channel
    cc = torch.arange(C, device="cuda").repeat(14 * 14)[None].expand_as(a)
    assert bool(torch.all((a % C == cc) | (a < 0)))
    # backward: total gradient mass is conserved (all RoIs well formed here)
    g = torch.ones_like(top)
    gb = ops.roi_pool_backward((B, H, W, C), rois, arg, g, 14, 14, c["scale"])
    assert abs(float(gb.double().sum()) - float((arg >= 0).sum())) < 1e-3 * float((arg >= 0).sum())
    gd = ops.roi_pool_backward((B, H, W, C), rois, arg, g, 14, 14, c["scale"], deterministic=True)
    np.testing.assert_allclose(gb.cpu().numpy(), gd.cpu().numpy(), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("kern,C", [("band", 96), ("band", 128), ("sorted", 96), ("sorted", 128)])
@pytest.mark.parametrize("mode", ["cpu", "gpu"])
def test_fwd_band_bins_taller_than_the_overlap(oracle_mod, mode, kern, C, tuning):
    """Band kernel: a 38x50 map is held as two overlapping row bands.  RoIs several times
    taller than the map have bins that no band holds completely: those take the kernel's
    global-memory path.  Mixed with ordinary RoIs, RoIs whose bins all belong to one band,
    and RoIs that start far above / end far below the map."""
    tuning("roi_fwd_kernel", kern)
    B, H, W = 3, 38, 50
    bottom = syn.feature_map(40, B, H, W, C)
    rng = np.random.default_rng(41)
    tall = []
    for i in range(60):
        x1 = rng.integers(-200, 700)
        y1 = rng.integers(-2500, 500)
        tall.append([i % B, x1, y1, x1 + rng.integers(16, 900), y1 + rng.integers(700, 4000)])
    top_only = [[b, 10, 0, 400, 150] for b in range(B)]        # rows 0..9: band 0 only
    bot_only = [[b, 300, 400, 790, 599] for b in range(B)]     # rows 25..37: band 1 only
    rois = np.concatenate([np.array(tall + top_only + bot_only, np.float32),
                           syn.rois_for_pool(42, 200, B), syn.adversarial_rois(B, W, H)])
    for PH, PW in ((7, 7), (2, 3), (14, 14)):
        wt, wa, t, a = _fwd_both(oracle_mod, bottom, rois, PH, PW, 1 / 16., mode)
        assert np.array_equal(a, wa) and np.array_equal(t, wt), (PH, PW)


def test_fwd_random_shape_sweep_both_kernels(oracle_mod, tuning):
    """Seeded sweep over map sizes, channel counts, pooled sizes, scales and RoI counts (incl.
    PH != PW, maps that barely fit / do not fit the shared-memory slice, C = 16, single RoIs):
    direct, tiled and band kernels against the oracle, both bin modes."""
    rng = np.random.default_rng(2024)
    for trial in range(24):
        B = int(rng.integers(1, 5))
        H, W = int(rng.integers(3, 61)), int(rng.integers(3, 61))
        C = int(rng.choice([16, 32, 48, 64, 80, 128, 20, 7, 256]))
        PH, PW = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        stride = int(rng.choice([4, 8, 16, 32]))
        R = int(rng.choice([1, 2, 17, 100, 333]))
        bottom = syn.feature_map(500 + trial, B, H, W, C)
        rois = syn.rois_for_pool(600 + trial, R, B, im_w=W * stride, im_h=H * stride)
        if trial % 5 == 0:                       # a few boxes hanging over the border
            rois[: max(1, R // 4), 3:5] += 3 * stride
        mode = "cpu" if trial % 2 == 0 else "gpu"
        want_top, want_arg = oracle_mod.clib.roi_pool_fwd(bottom, rois, PH, PW, 1.0 / stride,
                                                          bin_mode=0 if mode == "cpu" else 1)
        for kern in ("direct", "tiled", "band", "sorted"):
            tuning("roi_fwd_kernel", kern)
            top, arg = ops.roi_pool_forward(bottom, rois, PH, PW, 1.0 / stride, bin_mode=mode)
            ctx = (trial, kern, B, H, W, C, PH, PW, stride, R, mode)
            assert np.array_equal(arg.cpu().numpy(), want_arg), ctx
            assert np.array_equal(top.cpu().numpy(), want_top), ctx


def test_fwd_bwd_against_the_reference_kernel_itself(oracle_mod, fwd_kernel):
    """Every forward kernel and both backward modes against the object code of the
    reference's own RoiPoolOp / RoiPoolGradOp (roi_pooling_op.cc compiled unmodified into
    oracle/_ref/ref_roi_pool.so, see oracle/build_ref.py): BASELINE shapes C1 (300 proposal-like
    RoIs, 512 channels) forward, C2 (128 RoIs) forward + backward, plus adversarial RoIs."""
    if not oracle_mod.ref.roi_pool_available():
        pytest.skip("oracle/_ref/ref_roi_pool.so not built (reference absent)")
    ref = oracle_mod.ref
    c = syn.C1
    bottom = syn.feature_map(50, 1, c["H"], c["W"], c["C"])
    rois = np.concatenate([syn.rois_for_pool(51, 300), syn.adversarial_rois(1, c["W"], c["H"])])
    want_t, want_a = ref.roi_pool_fwd(bottom, rois, 7, 7, c["scale"], threads=16)
    top, arg = ops.roi_pool_forward(bottom, rois, 7, 7, c["scale"])
    assert np.array_equal(arg.cpu().numpy(), want_a) and np.array_equal(top.cpu().numpy(), want_t)
    # C2: 128 sampled RoIs, forward + backward on a 128-channel slice of the map (the
    # reference's gather is O(cells x RoIs))
    b2 = np.ascontiguousarray(bottom[..., :128])
    r2 = rois[:128]
    t2, a2 = ops.roi_pool_forward(b2, r2, 7, 7, c["scale"])
    wt2, wa2 = ref.roi_pool_fwd(b2, r2, 7, 7, c["scale"], threads=16)
    assert np.array_equal(a2.cpu().numpy(), wa2) and np.array_equal(t2.cpu().numpy(), wt2)
    g = np.random.default_rng(52).standard_normal(wt2.shape).astype(np.float32)
    want_g = ref.roi_pool_bwd(g, wa2, r2, b2.shape, c["scale"], threads=16)
    det = ops.roi_pool_backward(b2.shape, r2, a2, g, 7, 7, c["scale"], deterministic=True)
    assert np.array_equal(det.cpu().numpy(), want_g)
    atm = ops.roi_pool_backward(b2.shape, r2, a2, g, 7, 7, c["scale"], deterministic=False)
    np.testing.assert_allclose(atm.cpu().numpy(), want_g, rtol=RTOL, atol=ATOL)
