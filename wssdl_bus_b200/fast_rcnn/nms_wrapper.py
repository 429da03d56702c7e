"""fast_rcnn/nms_wrapper.py:13-21 twin: same dispatch, both branches run on the GPU.

cfg.USE_GPU_NMS only selects the *semantics*: False (the reference default, config.py:321)
-> cpu_nms rules (suppress iff (double)iou >= thresh); True -> gpu_nms rules (iou > thresh
in fp32).
"""
from wssdl_bus_b200.fast_rcnn.config import cfg
from wssdl_bus_b200.nms.cpu_nms import cpu_nms
from wssdl_bus_b200.nms.gpu_nms import gpu_nms


def nms(dets, thresh, force_cpu=False):
    """Dispatch to either CPU-semantics or GPU-semantics NMS (both on the device)."""
    if dets.shape[0] == 0:
        return []
    if cfg.USE_GPU_NMS and not force_cpu:
        return gpu_nms(dets, thresh, device_id=cfg.GPU_ID)
    return cpu_nms(dets, thresh)
