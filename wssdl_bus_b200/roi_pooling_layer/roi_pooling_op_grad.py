"""roi_pooling_layer/roi_pooling_op_grad.py:24-44 twin.  The reference registers the
gradient with TensorFlow; here `roi_pool` is a torch.autograd.Function whose backward is
`roi_pool_grad`, so importing this module is all a caller has to do (as in the reference,
where the import performs the registration)."""
from wssdl_bus_b200.ops import _RoiPoolFn as RoiPoolGradient  # noqa: F401
