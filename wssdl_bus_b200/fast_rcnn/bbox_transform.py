"""fast_rcnn/bbox_transform.py twin: bbox_transform (:10-28), bbox_transform_inv (:30-61),
clip_boxes (:63-77) on the device.  numpy in -> numpy out; CUDA tensors stay on the device."""
from wssdl_bus_b200.ops import bbox_transform, bbox_transform_inv, clip_boxes  # noqa: F401
