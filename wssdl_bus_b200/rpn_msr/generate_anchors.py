"""rpn_msr/generate_anchors.py:37-97 twin (host side, 9-12 boxes; runs once per call).

Closed form of the reference's ratio/scale enumeration around the (0,0,15,15) base
window: for every ratio r, w_r = round(sqrt(size / r)), h_r = round(w_r * r); for every
scale s the anchor is the (w_r*s) x (h_r*s) window centred on the base centre.  Ratio-major
order, as np.vstack over the ratio anchors gives (:46-47).
"""
import numpy as np


def generate_anchors(base_size=16, ratios=[0.5, 1, 2], scales=2 ** np.arange(3, 6)):
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    ctr = 0.5 * (base_size - 1)                       # centre of (0, 0, base-1, base-1)
    size = float(base_size) * float(base_size)
    ws = np.round(np.sqrt(size / ratios))             # :83
    hs = np.round(ws * ratios)                        # :84
    W = (ws[:, None] * scales[None, :]).reshape(-1)   # :94-95, ratio-major
    Hh = (hs[:, None] * scales[None, :]).reshape(-1)
    return np.stack([ctr - 0.5 * (W - 1), ctr - 0.5 * (Hh - 1),
                     ctr + 0.5 * (W - 1), ctr + 0.5 * (Hh - 1)], axis=1)


def shifted_anchors(height, width, feat_stride, base_anchors):
    """(K*A, 4) float64 anchors in (h, w, a) order (proposal_layer_tf_bus.py:55-71)."""
    sx = np.arange(width, dtype=np.float64) * feat_stride
    sy = np.arange(height, dtype=np.float64) * feat_stride
    shifts = np.stack([np.tile(sx, height), np.repeat(sy, width),
                       np.tile(sx, height), np.repeat(sy, width)], axis=1)
    return (shifts[:, None, :] + np.asarray(base_anchors, np.float64)[None, :, :]).reshape(-1, 4)
