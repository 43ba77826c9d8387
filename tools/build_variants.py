"""Developer A/B: builds of libsgn_raster.so that differ in the register budget of the blend kernels (occupancy experiment).
    python tools/build_variants.py 20 24 32   ->  street-gaussians-ns_b200/libsgn_raster_b20.so ... (min resident CTAs per SM;
    the budget is 65536 / (32 * N) registers per thread: 102 / 85 / 64.  -maxrregcount is IGNORED for kernels with
    __launch_bounds__, which is why the first round of this experiment measured four identical builds.)
Select one with SGN_RASTER_LIB=<path> (street-gaussians-ns_b200/_lib.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from street_gaussians_ns_b200 import build as b  # noqa: E402

b.build()  # the default objects
objdir = os.path.join(b.HERE, "build")
for cap in sys.argv[1:]:
    obj = os.path.join(objdir, f"blend_b{cap}.o")
    cmd = [b._nvcc(), *b.ARCH, *b.COMMON, *b.SOURCES["blend.cu"], f"-DBLEND_FWD_MIN_BLOCKS={cap}", f"-DBLEND_BWD_MIN_BLOCKS={cap}",
           f"-DBLEND_ACC_MIN_BLOCKS={cap}", "-c", os.path.join(b.CSRC, "blend.cu"), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    open(obj + ".log", "w").write(r.stdout + r.stderr)
    assert r.returncode == 0, r.stderr
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in b.SOURCES if s != "blend.cu"] + [obj]
    out = os.path.join(b.HERE, f"libsgn_raster_b{cap}.so")
    r = subprocess.run([b._nvcc(), *b.ARCH, "-shared", "-o", out, *objs], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    spills = [l for l in open(obj + ".log").read().splitlines() if "spill" in l and "0 bytes spill stores, 0 bytes spill loads" not in l]
    print(out, f"({len(spills)} kernels with spills)")
