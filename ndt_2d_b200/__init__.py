"""ndt_2d_b200 -- B200 (sm_100a) backend for ndt_2d's scan-matching hot path.

The product is libndt2d_b200.so (C ABI in include/ndt2d_b200.h, CUDA sources in
ndt_2d_b200/csrc/).  This package is the thin host-side mirror of the
reference's operator interface used by the tests and bench.py.

Attributes are resolved lazily so that `python -m ndt_2d_b200.build` (which creates
the shared library) can be imported before the library exists; touching anything else
loads the library and fails loudly if it is missing -- there is no pure-Python path.
"""
import importlib

_EXPORTS = {
    "Ndt2dError": "._lib", "lib": "._lib", "lib_path": "._lib",
    "MotionModel": ".particle_filter", "ParticleFilter": ".particle_filter",
    "ParameterNode": ".scan_matcher", "Pose2d": ".scan_matcher", "Scan": ".scan_matcher",
    "ScanMatcherNDT": ".scan_matcher", "laser_to_points": ".scan_matcher", "OccupancyGrid": ".scan_matcher",
    "find_nearest": ".scan_matcher", "graph_find_nearest": ".scan_matcher",
}

__all__ = sorted(_EXPORTS)


def __getattr__(name):
    if name in _EXPORTS:
        value = getattr(importlib.import_module(_EXPORTS[name], __name__), name)
        globals()[name] = value
        return value
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
