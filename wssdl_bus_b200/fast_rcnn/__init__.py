"""Mirror of the reference's ``fast_rcnn`` package for the hot path (config constants,
nms_wrapper, bbox_transform)."""
