"""CPU parity oracle for the detector hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in ``wssdl_bus_b200`` imports this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py``.

Pieces
  oracle.clib      ctypes view of liboracle.so (hotpath_ref.c): RoiPool fwd/bwd
                   restatement (parity UNPINNED by reference tests, see the C header),
                   plus C restatements of cpu_nms / bbox_overlaps[_ui] (pinned against
                   oracle/_ref and tests/golden).
  oracle.ref       the reference's own Cython modules compiled into oracle/_ref
                   (cpu_nms, cython_nms.nms/nms_new, bbox_overlaps, bbox_overlaps_ui).
  oracle.layers    numpy restatements of the python glue the reference cannot import
                   under py3 (generate_anchors, bbox_transform*, proposal_layer,
                   anchor/proposal target layers).
"""
from . import clib, layers, ref  # noqa: F401
