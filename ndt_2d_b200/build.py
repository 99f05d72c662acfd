"""Build libndt2d_b200.so in-tree with nvcc for sm_100a (no JIT, no torch extension).

    python -m ndt_2d_b200.build [--force]

The library lands in ndt_2d_b200/lib/ (git-ignored, but it travels to the GPU
box with the repo snapshot).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = LIBDIR / "obj"
LIB = LIBDIR / "libndt2d_b200.so"

CU_SOURCES = ["api.cu", "build.cu", "search.cu", "search_region.cu", "search_window.cu", "filter.cu", "probe.cu", "frontend.cu", "occupancy.cu"]
CXX_SOURCES = []
SYNTH_SRC = PKG / "synth_src" / "synth.cpp"
SYNTH_LIB = LIBDIR / "libndt2d_synth.so"
HEADERS = [CSRC / "ndt2d_internal.h", CSRC / "search_common.cuh", CSRC / "search_region_body.inc",
           CSRC / "build_common.cuh",
           ROOT / "include" / "ndt2d_b200.h"]

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + [
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall",
    "-Xptxas", "-v",
    f"-I{ROOT / 'include'}", f"-I{CSRC}",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libndt2d_b200 cannot be built (there is no CPU fallback)")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def _run(cmd, log: Path | None = None):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        log.write_text(" ".join(map(str, cmd)) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError(f"command failed: {' '.join(map(str, cmd))}")
    return proc


def build_synth(force: bool = False) -> Path:
    """The synthetic-world generator (test / bench infrastructure): its own host-only library."""
    LIBDIR.mkdir(parents=True, exist_ok=True)
    deps = [SYNTH_SRC, SYNTH_SRC.parent / "ndt2d_synth.h"]
    if force or _stale(SYNTH_LIB, deps):
        _run(["g++", "-O3", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off", "-Wall",
              "-shared", f"-I{SYNTH_SRC.parent}", str(SYNTH_SRC), "-o", str(SYNTH_LIB), "-lpthread"])
    return SYNTH_LIB


def build(force: bool = False, verbose: bool = False) -> Path:
    build_synth(force)
    nvcc = _nvcc()
    # A/B experiments: NDT2D_NVCC_EXTRA="-DNDT2D_REGION_WARPS=20" python -m ndt_2d_b200.build --force
    extra = os.environ.get("NDT2D_NVCC_EXTRA", "").split()
    OBJDIR.mkdir(parents=True, exist_ok=True)
    jobs = []
    for src in CU_SOURCES + CXX_SOURCES:
        s = CSRC / src
        o = OBJDIR / (src + ".o")
        if force or _stale(o, [s] + HEADERS + [Path(__file__)]):
            jobs.append((s, o))
    if jobs:
        def compile_one(job):
            s, o = job
            cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", str(s), "-o", str(o)]
            if verbose:
                print(" ".join(cmd))
            _run(cmd, log=OBJDIR / (s.name + ".log"))
        with ThreadPoolExecutor(max_workers=4) as ex:
            list(ex.map(compile_one, jobs))
    objs = [OBJDIR / (src + ".o") for src in CU_SOURCES + CXX_SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", str(LIB)] + [str(o) for o in objs] + ["-lpthread"]
        if verbose:
            print(" ".join(cmd))
        _run(cmd)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
