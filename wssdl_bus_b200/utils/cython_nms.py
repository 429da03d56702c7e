"""utils/nms.pyx twin: nms (:17-68, == cpu_nms) and nms_new (:70-123, + containment)."""
from wssdl_bus_b200 import ops


def nms(dets, thresh):
    return ops.nms(dets, thresh, ops.NMS_GE_F64)


def nms_new(dets, thresh):
    return ops.nms(dets, thresh, ops.NMS_GE_F64 | ops.NMS_CONTAIN)
