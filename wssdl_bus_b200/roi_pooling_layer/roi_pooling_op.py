"""roi_pooling_layer/roi_pooling_op.py:4-7 twin: `roi_pool` and `roi_pool_grad` backed by
libwssdl_b200.so instead of a TensorFlow custom-op library."""
from wssdl_bus_b200.ops import roi_pool, roi_pool_grad  # noqa: F401
