"""utils/bbox_ui.pyx:12-46 twin."""
from wssdl_bus_b200.ops import bbox_overlaps_ui  # noqa: F401
