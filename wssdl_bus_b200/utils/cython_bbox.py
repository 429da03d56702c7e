"""utils/bbox.pyx:15-55 twin (module name as built by the reference's setup.py:112-118)."""
from wssdl_bus_b200.ops import bbox_overlaps  # noqa: F401
