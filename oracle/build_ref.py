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

  ref_roi_pool.so    <- code/lib/roi_pooling_layer/roi_pooling_op.cc, UNMODIFIED: the RoiPool /
                        RoiPoolGrad CPU kernels.  The file needs the TensorFlow 1.x op framework,
                        which is not installed; it is compiled against ``oracle/tf_stub`` (a
                        stand-in that provides containers, status / attribute plumbing and the
                        registration macros and contains no arithmetic of the op) and linked
                        with ``oracle/ref_roi_pool_driver.cc`` (binds the op to numpy buffers,
                        defines Shard() as an even multi-threaded range split).  The kernels
                        are ALSO restated in ``oracle/hotpath_ref.c`` (faster, plus the CUDA
                        twin's bins); tests pin the restatement to this build.
  ref_gpu_nms_hostbuild.so <- code/lib/nms/nms_kernel.cu (devIoU, nms_kernel, _nms: the '>' NMS
                        behind gpu_nms) built FOR THE HOST the same way (cuda_emu.h).
  ref_roi_pool_cudatwin.so <- code/lib/roi_pooling_layer/roi_pooling_op_gpu.cu.cc built FOR THE
                        HOST (``oracle/tf_stub/cuda_emu.h``): the CUDA kernels' bodies (the
                        GPU_CEIL bin arithmetic) run once per emulated thread; pins bin_mode
                        GPU_CEIL of the restatement.

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


ROI_POOL_SRC = "code/lib/roi_pooling_layer/roi_pooling_op.cc"
ROI_POOL_SO = os.path.join(OUT, "ref_roi_pool.so")


def built():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.isfile(os.path.join(OUT, m + suffix)) for m in MODULES)


def roi_pool_built():
    return os.path.isfile(ROI_POOL_SO)


def build_roi_pool(force=False, verbose=False):
    """g++ the reference's roi_pooling_op.cc (where it lies) against oracle/tf_stub and link it
    with the driver.  Returns True when oracle/_ref/ref_roi_pool.so exists afterwards."""
    if roi_pool_built() and not force:
        return True
    src = os.path.join(REF, ROI_POOL_SRC)
    if not os.path.isfile(src):
        return False
    os.makedirs(OUT, exist_ok=True)
    flags = ["-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w",
             "-I", os.path.join(HERE, "tf_stub"), "-I", os.path.dirname(src)]
    objs = []
    for path, obj in ((src, "ref_roi_pooling_op.o"),
                      (os.path.join(HERE, "ref_roi_pool_driver.cc"), "ref_roi_pool_driver.o")):
        o = os.path.join(OUT, obj)
        r = subprocess.run(["g++"] + flags + ["-c", path, "-o", o], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed for %s:\n%s\n%s" % (path, r.stdout, r.stderr))
        objs.append(o)
    r = subprocess.run(["g++", "-shared"] + objs + ["-o", ROI_POOL_SO, "-lpthread"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    for o in objs:
        os.remove(o)
    if verbose:
        print("built", ROI_POOL_SO)
    return True


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


CUDA_TWIN_SRC = "code/lib/roi_pooling_layer/roi_pooling_op_gpu.cu.cc"
CUDA_TWIN_SO = os.path.join(OUT, "ref_roi_pool_cudatwin.so")


def cuda_twin_built():
    return os.path.isfile(CUDA_TWIN_SO)


def build_cuda_twin(force=False, verbose=False):
    """The reference's CUDA kernels (roi_pooling_op_gpu.cu.cc) compiled FOR THE HOST: g++ with
    oracle/tf_stub/cuda_emu.h.  One build-time source edit: the two triple-chevron launch
    statements become CUDA_EMU_LAUNCH(grid, block, kernel(args)) (a launch has no C++ spelling);
    the kernel bodies -- the arithmetic -- are compiled as they are."""
    if cuda_twin_built() and not force:
        return True
    src = os.path.join(REF, CUDA_TWIN_SRC)
    if not os.path.isfile(src):
        return False
    os.makedirs(OUT, exist_ok=True)
    text = open(src).read()
    text, n = re.subn(r"(\w+)<<<(.*?),\s*(\w+),\s*0,\s*d\.stream\(\)>>>\((.*?)\);",
                      r"CUDA_EMU_LAUNCH((\2), (\3), \1(\4));", text, flags=re.S)
    if n != 2:
        raise RuntimeError("expected 2 kernel launches in %s, found %d" % (CUDA_TWIN_SRC, n))
    tmp = os.path.join(OUT, "ref_cuda_twin_tmp.cc")
    with open(tmp, "w") as f:
        f.write(text)
    flags = ["-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w",
             "-DGOOGLE_CUDA=1", "-include", os.path.join(HERE, "tf_stub", "cuda_emu.h"),
             "-I", os.path.join(HERE, "tf_stub"), "-I", os.path.dirname(src)]
    try:
        r = subprocess.run(["g++"] + flags + ["-shared", tmp,
                                              os.path.join(HERE, "ref_roi_pool_cudatwin_driver.cc"),
                                              "-o", CUDA_TWIN_SO], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed for the CUDA twin:\n%s\n%s" % (r.stdout, r.stderr))
    finally:
        os.remove(tmp)                   # no reference-derived source text stays in the tree
    if verbose:
        print("built", CUDA_TWIN_SO)
    return True


CUDA_NMS_SRC = "code/lib/nms/nms_kernel.cu"
CUDA_NMS_SO = os.path.join(OUT, "ref_gpu_nms_hostbuild.so")


def cuda_nms_built():
    return os.path.isfile(CUDA_NMS_SO)


def build_cuda_nms(force=False, verbose=False):
    """The reference's CUDA NMS (nms/nms_kernel.cu: devIoU, nms_kernel, the host sweep _nms)
    compiled FOR THE HOST with oracle/tf_stub/cuda_emu.h.  One build-time source edit: the
    launch statement `nms_kernel<<<blocks, threads>>>(args);` becomes
    CUDA_EMU_LAUNCH_SHARED(blocks, threads, nms_kernel(args));"""
    if cuda_nms_built() and not force:
        return True
    src = os.path.join(REF, CUDA_NMS_SRC)
    if not os.path.isfile(src):
        return False
    os.makedirs(OUT, exist_ok=True)
    text = open(src).read()
    text, n = re.subn(r"(\w+)<<<(\w+),\s*(\w+)>>>\((.*?)\);",
                      r"CUDA_EMU_LAUNCH_SHARED(\2, \3, \1(\4));", text, flags=re.S)
    if n != 1:
        raise RuntimeError("expected 1 kernel launch in %s, found %d" % (CUDA_NMS_SRC, n))
    tmp = os.path.join(OUT, "ref_cuda_nms_tmp.cc")
    with open(tmp, "w") as f:
        f.write(text)
    flags = ["-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w",
             "-include", os.path.join(HERE, "tf_stub", "cuda_emu.h"), "-I", os.path.dirname(src)]
    try:
        r = subprocess.run(["g++"] + flags + ["-shared", tmp,
                                              os.path.join(HERE, "ref_gpu_nms_driver.cc"),
                                              "-o", CUDA_NMS_SO], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed for nms_kernel.cu:\n%s\n%s" % (r.stdout, r.stderr))
    finally:
        os.remove(tmp)
    if verbose:
        print("built", CUDA_NMS_SO)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref:", "built" if ok else "reference not present; nothing built")
    ok = build_roi_pool(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref/ref_roi_pool.so:", "built" if ok else "reference not present; not built")
    ok = build_cuda_twin(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref/ref_roi_pool_cudatwin.so:", "built" if ok else "reference not present; not built")
    ok = build_cuda_nms(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref/ref_gpu_nms_hostbuild.so:", "built" if ok else "reference not present; not built")


# ---------------------------------------------------------------------------------------------
# The reference's CALL SITES of the hot path, for the drop-in test (tests/test_dropin_gpu.py).
# Its layer modules are Python 2 sources; what travels to the GPU box (where /root/reference does
# not exist) is their compiled form: each file is read where it lies, passed through the
# mechanical py2 -> py3 shim (oracle/py2shim.py), compiled, and the CODE OBJECT is marshalled into
# oracle/_ref/callsites/<name>.bin -- a build output like the .so files beside it, git-ignored.
CALLSITES = {
    # module name the reference imports it under -> (file, lines whose int '/' is '//' in py2)
    "fast_rcnn.config": ("code/lib/fast_rcnn/config.py", ()),
    "fast_rcnn.nms_wrapper": ("code/lib/fast_rcnn/nms_wrapper.py", ()),
    "generate_anchors": ("code/lib/rpn_msr/generate_anchors.py", ()),
    "rpn_msr.proposal_layer_tf_bus": ("code/lib/rpn_msr/proposal_layer_tf_bus.py", ()),
    "rpn_msr.anchor_target_layer_tf_bus": ("code/lib/rpn_msr/anchor_target_layer_tf_bus.py", ()),
    "rpn_msr.proposal_target_layer_tf_bus": ("code/lib/rpn_msr/proposal_target_layer_tf_bus.py", (57, 135)),
}
CALLSITES_DIR = os.path.join(OUT, "callsites")


def callsites_built():
    return all(os.path.isfile(os.path.join(CALLSITES_DIR, n + ".bin")) for n in CALLSITES)


def build_callsites(force=False):
    """Compile the reference's call-site modules into oracle/_ref/callsites/*.bin."""
    import marshal
    from .py2shim import py2_to_py3
    if callsites_built() and not force:
        return True
    if not all(os.path.isfile(os.path.join(REF, f)) for f, _ in CALLSITES.values()):
        return False
    os.makedirs(CALLSITES_DIR, exist_ok=True)
    for name, (rel, int_div) in CALLSITES.items():
        src = py2_to_py3(open(os.path.join(REF, rel)).read(), int_div)
        code = compile(src, "<reference>/" + rel, "exec")
        with open(os.path.join(CALLSITES_DIR, name + ".bin"), "wb") as f:
            marshal.dump((tuple(sys.version_info[:2]), rel, code), f)
    return True


def load_callsite(name):
    """-> (relative path of the reference file, code object), or None when it was built by another
    Python version / not built."""
    import marshal
    path = os.path.join(CALLSITES_DIR, name + ".bin")
    if not os.path.isfile(path):
        return None
    with open(path, "rb") as f:
        ver, rel, code = marshal.load(f)
    if tuple(ver) != tuple(sys.version_info[:2]):
        return None
    return rel, code
