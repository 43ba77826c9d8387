"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/sgn_raster.h declares, and its struct layouts match the ctypes mirror.  No compute
calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import street_gaussians_ns_b200.build as b
    from street_gaussians_ns_b200 import _lib
    b.build()
    return _lib.load(), _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sgn_raster.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sgn_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    L, mod = lib
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/sgn_raster.h but not exported by libsgn_raster.so"
    assert sorted(mod.EXPORTS) == syms


def test_abi_version_and_layouts(lib):
    L, mod = lib
    assert L.sgn_abi_version() == 2
    assert L.sgn_sizeof_segment() == ctypes.sizeof(mod.Segment) == 168
    assert L.sgn_sizeof_camera() == ctypes.sizeof(mod.CameraStruct)
    assert L.sgn_sizeof_segment_grads() == ctypes.sizeof(mod.SegmentGrads) == 48


def test_argument_validation_without_gpu(lib):
    L, mod = lib
    cs = mod.CameraStruct()
    cs.block_width, cs.width, cs.height, cs.sh_degree = 32, 64, 48, 3
    rc = L.sgn_project_fwd(ctypes.c_void_p(16), 1, 10, 1, ctypes.byref(cs), ctypes.c_void_p(16), ctypes.c_void_p(16),
                           ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), None)
    assert rc == -1 and b"block_width" in L.sgn_last_error()


def test_sass_is_sm100a(lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "street-gaussians-ns_b200", "libsgn_raster.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "street-gaussians-ns_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace(
                    "oracle/sgn_oracle.c", ""), f
