"""Host-side mirror of ndt_2d::ParticleFilter (include/ndt_2d/particle_filter.hpp:45-115,
src/particle_filter.cpp) over the C ABI.  The particle set lives on the device."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from .scan_matcher import Scan, ScanMatcherNDT


class MotionModel:
    """ndt_2d::MotionModel's parameters (motion_model.hpp:48): odom_alpha1..5."""

    def __init__(self, a1: float, a2: float, a3: float, a4: float, a5: float):
        self.alphas = np.array([a1, a2, a3, a4, a5], dtype=np.float64)


class ParticleFilter:
    def __init__(self, min_particles: int, max_particles: int, motion_model: Optional[MotionModel] = None,
                 device: int = -1, stream: int = 0, seed: int = 1):
        self._h = C.c_void_p()
        self.motion_model = motion_model or MotionModel(0.2, 0.2, 0.2, 0.2, 0.2)
        self.min_particles, self.max_particles = min_particles, max_particles
        self._seed = seed
        L.check(L.lib.ndt2d_filter_create(min_particles, max_particles, device, stream or None,
                                          C.byref(self._h)), "ndt2d_filter_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib.ndt2d_filter_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _next_seed(self) -> int:
        self._seed += 1
        return self._seed

    # particle_filter.cpp:53-69
    def init(self, x, y, theta, sigma_x, sigma_y, sigma_theta, seed: Optional[int] = None) -> None:
        L.check(L.lib.ndt2d_filter_init(self._h, x, y, theta, sigma_x, sigma_y, sigma_theta,
                                        seed if seed is not None else self._next_seed()),
                "ndt2d_filter_init")

    # particle_filter.cpp:71-76
    def update(self, dx, dy, dth, seed: Optional[int] = None) -> None:
        L.check(L.lib.ndt2d_filter_update(self._h, dx, dy, dth, L.dptr(self.motion_model.alphas),
                                          seed if seed is not None else self._next_seed()),
                "ndt2d_filter_update")

    # particle_filter.cpp:78-89
    def measure(self, matcher: ScanMatcherNDT, scan: Scan) -> None:
        pts = L.f64(scan.points).reshape(-1, 2)
        L.check(L.lib.ndt2d_filter_measure(self._h, matcher.handle, L.dptr(pts), pts.shape[0]),
                "ndt2d_filter_measure")

    # particle_filter.cpp:91-137
    def resample(self, kld_err: float, kld_z: float, uniforms: Optional[np.ndarray] = None,
                 seed: Optional[int] = None) -> None:
        u = L.f64(uniforms) if uniforms is not None else None
        L.check(L.lib.ndt2d_filter_resample(self._h, kld_err, kld_z, L.dptr(u),
                                            0 if u is None else u.shape[0],
                                            seed if seed is not None else self._next_seed()),
                "ndt2d_filter_resample")

    def getMean(self) -> np.ndarray:
        mean = np.zeros(3)
        L.check(L.lib.ndt2d_filter_stats(self._h, L.dptr(mean), None), "ndt2d_filter_stats")
        return mean

    def getCovariance(self) -> np.ndarray:
        cov = np.zeros((3, 3))
        L.check(L.lib.ndt2d_filter_stats(self._h, None, L.dptr(cov)), "ndt2d_filter_stats")
        return cov

    # direct state access (particles_, weights_)
    def size(self) -> int:
        n = C.c_size_t(0)
        L.check(L.lib.ndt2d_filter_size(self._h, C.byref(n)), "ndt2d_filter_size")
        return int(n.value)

    def set_particles(self, particles, weights) -> None:
        p = L.f64(particles).reshape(-1, 3)
        w = L.f64(weights).reshape(-1)
        assert p.shape[0] == w.shape[0]
        L.check(L.lib.ndt2d_filter_set_particles(self._h, L.dptr(p), L.dptr(w), p.shape[0]),
                "ndt2d_filter_set_particles")

    def get_particles(self):
        n = self.size()
        p, w = np.zeros((n, 3)), np.zeros(n)
        L.check(L.lib.ndt2d_filter_get_particles(self._h, L.dptr(p), L.dptr(w)), "ndt2d_filter_get_particles")
        return p, w

    def set_covariance(self, cov) -> None:
        c = L.f64(cov).reshape(3, 3)
        L.check(L.lib.ndt2d_filter_set_cov(self._h, L.dptr(c)), "ndt2d_filter_set_cov")

    def last_draws(self) -> np.ndarray:
        out = np.zeros(self.size(), dtype=np.uint64)
        L.check(L.lib.ndt2d_filter_last_draws(self._h, L.u64ptr(out)), "ndt2d_filter_last_draws")
        return out
