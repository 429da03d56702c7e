"""CPU parity oracle for the detector hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in ``wssdl_bus_b200`` imports this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py``.

Pieces
  oracle.clib      ctypes view of liboracle.so (hotpath_ref.c): C restatements of RoiPool
                   fwd/bwd (both bin modes), cpu_nms / nms_new, bbox_overlaps[_ui] -- every one
                   pinned bit for bit to oracle.ref (tests/test_oracle.py) and tests/golden.
  oracle.ref       the reference's OWN code compiled into oracle/_ref by oracle/build_ref.py:
                   the Cython modules (cpu_nms, cython_nms.nms/nms_new, bbox_overlaps,
                   bbox_overlaps_ui), the RoiPool / RoiPoolGrad C++ op (unmodified source
                   against oracle/tf_stub), and host builds of its two CUDA sources
                   (roi_pooling_op_gpu.cu.cc, nms_kernel.cu).
  oracle.layers    numpy restatements of the python glue the reference cannot import
                   under py3 (generate_anchors, bbox_transform*, proposal_layer,
                   anchor/proposal target layers, detection post-processing, VOC eval),
                   pinned by tests/golden/reference_layers_golden.npz: outputs of the
                   reference's own layer code run under a mechanical py2->py3 shim
                   (tests/golden/make_layers_golden.py).
  oracle.philox    numpy Philox4x32-10 + the two selection rules of the device-side subsampling
                   mode of the target layers (pinned by Random123's known-answer vectors).
"""
from . import clib, layers, philox, ref  # noqa: F401
