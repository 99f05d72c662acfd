"""ndt_2d_b200 -- B200 (sm_100a) backend for ndt_2d's scan-matching hot path.

The product is libndt2d_b200.so (C ABI in include/ndt2d_b200.h, CUDA sources in
ndt_2d_b200/csrc/).  This package is the thin host-side mirror of the
reference's operator interface used by the tests and bench.py.
"""
from ._lib import Ndt2dError, lib, lib_path  # noqa: F401  (import fails loudly without the .so)
from .particle_filter import MotionModel, ParticleFilter  # noqa: F401
from .scan_matcher import ParameterNode, Pose2d, Scan, ScanMatcherNDT  # noqa: F401

__all__ = ["ScanMatcherNDT", "ParticleFilter", "MotionModel", "ParameterNode", "Pose2d", "Scan",
           "Ndt2dError", "lib", "lib_path"]
