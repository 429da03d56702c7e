"""nms/cpu_nms.pyx:17-68 twin, on the GPU with cpu_nms's exact predicate."""
from wssdl_bus_b200 import ops


def cpu_nms(dets, thresh):
    """dets [N,5] f32 (x1,y1,x2,y2,score), thresh Python float -> list of kept indices in
    descending-score order.  Suppress iff (double)iou_f32 >= thresh (cpu_nms.pyx:65)."""
    return ops.nms(dets, thresh, ops.NMS_GE_F64)
