"""Seeded synthetic inputs for the hot path; numpy only.

Shared by tests/ and bench.py so that the CUDA path and the CPU oracle always see the
same bytes.  All generators take a seed and use np.random.default_rng(seed).
"""
import numpy as np

# BASELINE.json configs
C1 = dict(name="C1 VGG-16 test", B=1, H=38, W=50, C=512, A=9, PH=7, PW=7, scale=1.0 / 16,
          im_h=600, im_w=800, pre=6000, post=300, thresh=0.7)
C2 = dict(name="C2 VGG-16 train", B=1, H=38, W=50, C=512, A=9, PH=7, PW=7, scale=1.0 / 16,
          im_h=600, im_w=800, pre=2000, post=2000, thresh=0.7, sampled=128)
C3 = dict(name="C3 ResNet-101 C4", B=16, H=38, W=50, C=1024, A=9, PH=14, PW=14, scale=1.0 / 16,
          im_h=600, im_w=800, rois_per_image=300)


def feature_map(seed, B, H, W, C):
    """relu(N(0,1)) f32 NHWC: about half the entries are exact zeros (ties)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, H, W, C), dtype=np.float32)
    return np.maximum(x, 0, out=x)


def rpn_outputs(seed, B, H, W, A, im_h=600, im_w=800, im_scale=1.0, info_cols=3):
    """rpn_cls_prob_reshape [B,H,W,2A], rpn_bbox_pred [B,H,W,4A], im_info [B,info_cols].
    fg scores are a random permutation of (i+0.5)/(H*W*A) -> unique fp32 values."""
    rng = np.random.default_rng(seed)
    n = H * W * A
    cls = np.empty((B, H, W, 2 * A), np.float32)
    for b in range(B):
        fg = ((rng.permutation(n) + 0.5) / n).astype(np.float32).reshape(H, W, A)
        cls[b, :, :, A:] = fg
        cls[b, :, :, :A] = 1.0 - fg
    reg = (rng.standard_normal((B, H, W, 4 * A)) * 0.5).astype(np.float32)
    reg4 = reg.reshape(B, H, W, A, 4)
    np.clip(reg4[..., 2:], -2.0, 2.0, out=reg4[..., 2:])     # dw, dh
    info = np.zeros((B, info_cols), np.float32)
    info[:, 0], info[:, 1], info[:, 2] = im_h, im_w, im_scale
    return cls, reg, info


def random_boxes(seed, n, im_w=800, im_h=600, lo=16.0, hi=400.0, clustered=False):
    """[n,4] f32 boxes: centres uniform (or jittered around 50 seeds), sides log-uniform."""
    rng = np.random.default_rng(seed)
    if clustered:
        k = 50
        cx0, cy0 = rng.uniform(0, im_w, k), rng.uniform(0, im_h, k)
        w0 = np.exp(rng.uniform(np.log(lo), np.log(hi), k))
        h0 = np.exp(rng.uniform(np.log(lo), np.log(hi), k))
        a = rng.integers(0, k, n)
        cx = cx0[a] + rng.normal(0, 8, n)
        cy = cy0[a] + rng.normal(0, 8, n)
        w = w0[a] * np.exp(rng.normal(0, 0.1, n))
        h = h0[a] * np.exp(rng.normal(0, 0.1, n))
    else:
        cx, cy = rng.uniform(0, im_w, n), rng.uniform(0, im_h, n)
        w = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
        h = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, im_w - 1)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, im_h - 1)
    return b.astype(np.float32)


def dets(seed, n, **kw):
    """[n,5] f32 (x1,y1,x2,y2,score) with unique scores."""
    rng = np.random.default_rng(seed + 7919)
    b = random_boxes(seed, n, **kw)
    s = ((rng.permutation(n) + 0.5) / n).astype(np.float32)
    return np.hstack([b, s[:, None]])


def rois_for_pool(seed, R, B=1, im_w=800, im_h=600):
    """[R,5] f32 well-formed RoIs (batch idx, x1,y1,x2,y2), batch indices shuffled."""
    rng = np.random.default_rng(seed + 104729)
    b = random_boxes(seed, R, im_w, im_h)
    bi = rng.integers(0, B, R).astype(np.float32)
    return np.hstack([bi[:, None], b]).astype(np.float32)


def adversarial_rois(B=1, W=50, H=38, stride=16):
    """RoIs that hit the semantic forks of SURVEY.md Appendix A.1 / section 7:
    tiny RoIs (empty CPU bins), border-touching and out-of-image RoIs, rounded sides of
    31 / 57 / 62 cells (fp32 edge products), half-way roundings (x=8 -> 0.5 -> 1), malformed
    x2<x1 / y2<y1 boxes."""
    s = float(stride)
    r = [
        [0, 0, 0, 3 * s - 1, 3 * s - 1],              # 3x3 cells < 7 bins -> empty bins
        [0, 16, 16, 16, 16],                          # single cell
        [0, 0, 0, W * s - 1, H * s - 1],              # whole map
        [0, -40, -40, 100, 90],                       # starts outside
        [0, (W - 3) * s, (H - 2) * s, (W + 4) * s, (H + 5) * s],   # ends outside
        [0, 0, 0, 30 * s, 30 * s],                    # 31 cells: 7*fl(31/7) < 31
        [0, 8, 8, 8 + 30 * s, 8 + 30 * s],            # half-way rounding of the corners
        [0, 24, 40, 24 + 36 * s, 40 + 30 * s],
        [0, 0, 0, 56 * s, 20 * s],                    # 57 cells wide (wider than the map)
        [0, 0, 0, 61 * s, 36 * s],                    # 62 cells
        [0, 300, 200, 100, 400],                      # malformed: x2 < x1
        [0, 100, 400, 300, 200],                      # malformed: y2 < y1
        [0, 500, 500, 100, 100],                      # malformed both
        [0, 7.9, 8.1, 135.5, 120.49],                 # fractional coordinates
        [0, 799, 599, 799, 599],                      # last pixel
    ]
    r = np.asarray(r, np.float32)
    if B > 1:
        r = np.concatenate([np.concatenate([np.full((len(r), 1), b, np.float32), r[:, 1:]], 1)
                            for b in range(B)])
    return r


def gt_boxes(seed, B, max_gt=20, im_w=800, im_h=600, n_fg=(1, 3), n_bg=(0, 2)):
    """gt_boxes [B,max_gt,5] f32 (fg rows, class 1/2, first; then class-0 background boxes),
    num_gt [B] i32 -- the layout contract of anchor_target_layer_tf_bus.py:434-436."""
    rng = np.random.default_rng(seed + 15485863)
    gt = np.zeros((B, max_gt, 5), np.float32)
    num = np.zeros((B,), np.int32)
    for b in range(B):
        nf = int(rng.integers(n_fg[0], n_fg[1] + 1))
        nb = int(rng.integers(n_bg[0], n_bg[1] + 1))
        boxes = random_boxes(seed * 131 + b, nf + nb, im_w, im_h, lo=40.0, hi=300.0)
        gt[b, :nf + nb, :4] = boxes
        gt[b, :nf, 4] = rng.integers(1, 3, nf)
        num[b] = nf + nb
    return gt, num


def rcnn_head_outputs(seed, n_rows, K=3, delta_std=0.1):
    """Synthetic RCNN head outputs for n_rows RoIs: cls_prob [n_rows,K] f32 (rows sum to ~1,
    every entry a distinct fp32 value so NMS order is well defined) and bbox_pred [n_rows,4K]
    f32 ~ N(0, delta_std) with dw, dh clipped to +-1."""
    rng = np.random.default_rng(seed + 32452843)
    n = n_rows * K
    u = ((rng.permutation(n) + 0.5) / n).astype(np.float64).reshape(n_rows, K)
    u[:, 0] *= 3.0                                      # background usually wins
    p = (u / u.sum(1, keepdims=True)).astype(np.float32)
    # normalisation can merge neighbours: nudge duplicates apart (deterministically)
    flat = p.reshape(-1)
    order = np.argsort(flat, kind="stable")
    for a, b in zip(order[:-1], order[1:]):
        if flat[b] <= flat[a]:
            flat[b] = np.nextafter(flat[a], np.float32(2.0))
    d = (rng.standard_normal((n_rows, 4 * K)) * delta_std).astype(np.float32)
    d4 = d.reshape(n_rows, K, 4)
    np.clip(d4[..., 2:], -1.0, 1.0, out=d4[..., 2:])
    return p, d
