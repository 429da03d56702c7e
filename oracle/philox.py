"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the device-side subsampling stream of the two
target layers (include/wssdl_b200.h, WSSDL_SAMPLE_PHILOX; csrc/targets.cu): Philox4x32-10
(Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
generator of Random123 / cuRAND) and the two selection rules built on it.  The generator is
pinned by Random123's published known-answer vectors (tests/test_oracle.py); the GPU tests
compare the device selections with these functions bit for bit.  Not a reference-derived
algorithm: the reference draws from numpy.random on the host (the layers' default mode)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Counters c0..c3 (arrays or scalars, uint32 values), key (k0, k1) -> four uint32 arrays."""
    c = [np.asarray(v, dtype=np.uint64) & MASK for v in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [(hi1 ^ c[1] ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c[3] ^ np.uint64(k1)) & MASK, lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [v.astype(np.uint32) for v in c]


def _keys(index, image, which, seed):
    seed = int(seed) & (2 ** 64 - 1)
    return philox4x32_10(index, image, which, 0, seed & 0xFFFFFFFF, seed >> 32)[0].astype(np.uint64)


def anchor_subsample(labels, image, seed, num_fg, batchsize):
    """labels [NA] (-1 / 0 / 1, anchors in (h, w, a) order) -> labels after the fg / bg subsampling
    of anchor_target_layer (:512-528) with the device stream: anchor i gets the key
    philox(counter (i, image, which, 0), key seed).x, which = 0 fg / 1 bg, and the surplus anchors
    with the smallest (key, i) are disabled."""
    out = np.array(labels, copy=True)
    idx = np.arange(out.size, dtype=np.uint64)
    fg = np.where(out == 1)[0]
    kill = max(len(fg) - num_fg, 0)
    if kill:
        comp = (_keys(idx[fg], image, 0, seed) << np.uint64(15)) | idx[fg]
        out[fg[np.argsort(comp, kind="stable")[:kill]]] = -1
    num_bg = batchsize - int((out == 1).sum())
    bg = np.where(out == 0)[0]
    kill = max(len(bg) - max(num_bg, 0), 0)
    if kill:
        comp = (_keys(idx[bg], image, 1, seed) << np.uint64(15)) | idx[bg]
        out[bg[np.argsort(comp, kind="stable")[:kill]]] = -1
    return out


def roi_select(n_fg, n_bg, image, seed, fg_quota, rois_per_image):
    """Ranks (among the fg / bg candidates of an image, in candidate order) that
    proposal_target_layer's device sampler selects, in output order: the fg_this (bg_this)
    candidates with the smallest (philox(counter (rank, image, 2 fg / 3 bg, 0), key seed).x, rank)."""
    fg_this = min(fg_quota, n_fg)
    bg_this = min(rois_per_image - fg_this, n_bg)
    out = []
    for which, n, k in ((2, n_fg, fg_this), (3, n_bg, bg_this)):
        r = np.arange(n, dtype=np.uint64)
        comp = (_keys(r, image, which, seed) << np.uint64(32)) | r
        out.append(np.argsort(comp, kind="stable")[:k].astype(np.int64))
    return out[0], out[1]
