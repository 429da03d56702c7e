"""Build the reference's own Cython hot-path modules into ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it.

What is built (sources are read where they lie under /root/reference, never copied
into the repository; every output goes to ``oracle/_ref/`` which is git-ignored):

  ref_cpu_nms        <- code/lib/nms/cpu_nms.pyx        (cpu_nms, the live NMS path)
  ref_cython_nms     <- code/lib/utils/nms.pyx          (nms, nms_new: test-loop twins)
  ref_cython_bbox    <- code/lib/utils/bbox.pyx         (bbox_overlaps, fp64)
  ref_cython_bbox_ui <- code/lib/utils/bbox_ui.pyx      (bbox_overlaps_ui, fp64)

Build-time patches (token level, semantics unchanged):
  * ``np.int_t`` -> ``np.intp_t``: Cython 3 / numpy 2 no longer export ``np.int_t``
    (cpu_nms.pyx:25,28; utils/nms.pyx same lines).  The buffer dtype of
    ``argsort()`` is intp either way.
  * the module name (so the four modules can coexist in one directory).
At import time ``oracle.ref`` sets ``np.float = float`` and ``np.int = int`` before
importing, because the modules evaluate ``np.float`` / ``np.int`` at run time
(cpu_nms.pyx:17,29; bbox.pyx:12).  ``thresh`` stays a Python float, so the
comparison stays ``(double)ovr >= thresh`` exactly as in the reference.

The RoiPool/RoiPoolGrad CPU kernels (roi_pooling_op.cc) need TensorFlow 1.x headers
and are NOT buildable here; they are restated in ``oracle/roi_pool_ref.c``.

Flags: ``-O2`` only; no ``-march=native``, no ``-ffast-math`` and
``-ffp-contract=off`` so that x86 FMA contraction can never change a result.
"""
import os
import re
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("WSSDL_REFERENCE", "/root/reference")

MODULES = {
    "ref_cpu_nms": "code/lib/nms/cpu_nms.pyx",
    "ref_cython_nms": "code/lib/utils/nms.pyx",
    "ref_cython_bbox": "code/lib/utils/bbox.pyx",
    "ref_cython_bbox_ui": "code/lib/utils/bbox_ui.pyx",
}


def have_reference():
    return all(os.path.isfile(os.path.join(REF, p)) for p in MODULES.values())


def built():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.isfile(os.path.join(OUT, m + suffix)) for m in MODULES)


def build(force=False, verbose=False):
    """Cythonize + compile the four reference modules.  Returns True when built."""
    if built() and not force:
        return True
    if not have_reference():
        return False
    import numpy as np

    os.makedirs(OUT, exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    py_inc = sysconfig.get_paths()["include"]
    np_inc = np.get_include()
    for mod, rel in MODULES.items():
        with open(os.path.join(REF, rel)) as f:
            src = f.read()
        src = re.sub(r"\bnp\.int_t\b", "np.intp_t", src)
        pyx = os.path.join(OUT, mod + ".pyx")
        with open(pyx, "w") as f:
            f.write(src)
        c = os.path.join(OUT, mod + ".c")
        cmd = [sys.executable, "-m", "cython", "-3", pyx, "-o", c]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("cython failed for %s:\n%s\n%s" % (rel, r.stdout, r.stderr))
        so = os.path.join(OUT, mod + suffix)
        cmd = ["gcc", "-shared", "-fPIC", "-O2", "-ffp-contract=off", "-w",
               "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
               "-I", py_inc, "-I", np_inc, c, "-o", so]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc failed for %s:\n%s\n%s" % (rel, r.stdout, r.stderr))
        # keep only the binary: no reference-derived source text stays in the tree
        os.remove(pyx)
        os.remove(c)
        if verbose:
            print("built", so)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref:", "built" if ok else "reference not present; nothing built")
