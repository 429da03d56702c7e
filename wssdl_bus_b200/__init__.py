"""wssdl_bus_b200 -- the B200 (sm_100a) implementation of the detector hot path of
syshin1014/wssdl_bus: NHWC RoI max pooling (fwd+argmax, bwd), RPN proposal generation
(decode, clip, min-size filter, top-k, NMS), IoU matrices and anchor labelling.

Layout mirrors the reference's ``code/lib`` so its call sites stay unchanged:

    reference import                                   this package
    -------------------------------------------------  -----------------------------------
    roi_pooling_layer.roi_pooling_op.roi_pool          wssdl_bus_b200.roi_pooling_layer...
    nms.cpu_nms.cpu_nms / nms.gpu_nms.gpu_nms          wssdl_bus_b200.nms...
    utils.cython_bbox.bbox_overlaps (+ _ui, _nms)      wssdl_bus_b200.utils...
    fast_rcnn.nms_wrapper.nms / bbox_transform.*       wssdl_bus_b200.fast_rcnn...
    rpn_msr.proposal_layer_tf_bus.proposal_layer ...   wssdl_bus_b200.rpn_msr...

``install_dropin()`` puts this directory at the front of ``sys.path`` so that the
reference's own top-level imports (``from nms.cpu_nms import cpu_nms``) resolve here.
All compute happens in libwssdl_b200.so (hand-written CUDA behind a C ABI, see
include/wssdl_b200.h); there is no CPU fallback.
"""
import os
import sys

from . import _lib, ops  # noqa: F401
from .ops import (anchor_labels, bbox_overlaps, bbox_overlaps_ui, bbox_transform,  # noqa: F401
                  bbox_transform_inv, clip_boxes, nms, nms_device, proposals, roi_pool,
                  roi_pool_grad)

__version__ = "1.0"


def install_dropin():
    """Make `import nms.cpu_nms`, `import utils.cython_bbox`, ... resolve to this package."""
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    return here
