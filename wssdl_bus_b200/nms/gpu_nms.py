"""nms/gpu_nms.pyx:16-31 twin: '>' predicate in fp32 (nms_kernel.cu:71)."""
from wssdl_bus_b200 import ops


def gpu_nms(dets, thresh, device_id=0):
    del device_id  # the current torch device is used
    return ops.nms(dets, thresh, ops.NMS_GT_F32)
