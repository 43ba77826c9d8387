"""TEST INFRASTRUCTURE: g++ build of the SGN_HD row rules the refinement kernels are made of (refine_host.cpp), loaded
through ctypes behind the two C-ABI signatures.  Used by CPU tests only (no GPU in the build container)."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def load_refine_harness():
    from street_gaussians_ns_b200 import _lib
    src = os.path.join(HERE, "refine_host.cpp")
    out = os.path.join(HERE, "librefine_host.so")
    rules = os.path.join(ROOT, "street-gaussians-ns_b200", "csrc", "sgn_refine_rules.cuh")
    header = os.path.join(ROOT, "include", "sgn_raster.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(map(os.path.getmtime, (src, rules, header))):
        tmp = f"{out}.{os.getpid()}.tmp"
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", tmp, src], check=True)
        os.replace(tmp, out)  # atomic: several test processes may build at once
    H = C.CDLL(out)
    vp, i32 = C.c_void_p, C.c_int
    H.sgn_refine_decide.argtypes = [i32, C.POINTER(_lib.RefineConfig), vp, vp, vp, vp, vp, vp, vp, vp]
    H.sgn_refine_apply.argtypes = [i32, C.POINTER(_lib.RefineConfig), C.POINTER(_lib.RefineTensors), vp, vp, C.POINTER(C.c_int32), vp, vp]
    H.sgn_sizeof_refine_config.restype = H.sgn_sizeof_refine_tensors.restype = C.c_size_t
    assert H.sgn_sizeof_refine_config() == C.sizeof(_lib.RefineConfig)
    assert H.sgn_sizeof_refine_tensors() == C.sizeof(_lib.RefineTensors)
    return H


def use_host_backend(refine_module, harness=None):
    """Route refine.py's two library calls to the harness and accept CPU tensors (what the monkeypatch fixture does)."""
    harness = harness or load_refine_harness()
    refine_module._backend = lambda: harness
    refine_module._require_cuda = lambda t, what: None
    return harness
