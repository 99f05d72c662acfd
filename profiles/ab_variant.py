"""Kernel A/B helper: builds ndt_2d_b200/lib/ab/libndt2d_b200_<name>.so with extra nvcc flags for ONE
translation unit (the other objects are the production ones); run a script against it with
NDT2D_B200_LIB=<that path>.

    python profiles/ab_variant.py w28 search_region.cu -DNDT2D_REGION_WARPS=28
"""
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ndt_2d_b200 import build as nb  # noqa: E402

name, unit, extra = sys.argv[1], sys.argv[2], sys.argv[3:]
nb.build()
out_dir = nb.LIBDIR / "ab"
out_dir.mkdir(parents=True, exist_ok=True)
obj = out_dir / f"{unit}.{name}.o"
log = out_dir / f"{unit}.{name}.log"
cmd = [nb._nvcc()] + nb.NVCC_FLAGS + extra + ["-c", str(nb.CSRC / unit), "-o", str(obj)]
p = subprocess.run(cmd, capture_output=True, text=True)
log.write_text(p.stdout + p.stderr)
if p.returncode:
    sys.exit(p.stdout + p.stderr)
objs = [obj if s == unit else nb.OBJDIR / (s + ".o") for s in nb.CU_SOURCES]
lib = out_dir / f"libndt2d_b200_{name}.so"
subprocess.run([nb._nvcc()] + nb.ARCH + ["-shared", "-o", str(lib)] + [str(o) for o in objs] + ["-lpthread"],
               check=True)
print(lib)
