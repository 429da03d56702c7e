"""nms/py_cpu_nms.py:10-38 twin: keeps iou <= thresh, i.e. the '>' predicate in fp32."""
from wssdl_bus_b200 import ops


def py_cpu_nms(dets, thresh):
    return ops.nms(dets, thresh, ops.NMS_GT_F32)
