"""Synthetic laser world and the BASELINE.json workloads.

Test / bench infrastructure, host code only: thin wrappers over the generator in its own
library ndt_2d_b200/lib/libndt2d_synth.so (synth_src/synth.cpp, g++) so that tests, bench.py
and the oracle all see bit-identical inputs.  Importing this module loads nothing of the
product (libndt2d_b200.so): the CPU reference arm of bench.py uses it too.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from pathlib import Path

_SYNTH_SO = Path(__file__).resolve().parent / "lib" / "libndt2d_synth.so"
_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)


class _L:
    """The synthetic-world library + the few helpers this module needs."""

    def __init__(self):
        if not _SYNTH_SO.exists():
            from . import build as _build
            _build.build_synth()
        self.lib = C.CDLL(str(_SYNTH_SO))
        sig = {
            "ndt2d_synth_world": (C.c_int, [C.c_uint64, C.c_double, C.c_int, C.c_double, C.c_double, _dp]),
            "ndt2d_synth_scans": (C.c_int, [_dp, C.c_int, C.c_double, _dp, C.c_size_t, C.c_int, C.c_double,
                                            C.c_double, C.c_uint64, _u64p, _dp]),
            "ndt2d_synth_uniform": (None, [C.c_uint64, C.c_size_t, _dp]),
            "ndt2d_synth_normal": (None, [C.c_uint64, C.c_size_t, _dp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(self.lib, name)
            fn.restype, fn.argtypes = res, args

    @staticmethod
    def check(status: int, where: str):
        if status != 0:
            raise RuntimeError(f"{where}: status {status}")

    @staticmethod
    def dptr(a):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need contiguous float64"
        return a.ctypes.data_as(_dp)

    @staticmethod
    def u64ptr(a):
        assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], "need contiguous uint64"
        return a.ctypes.data_as(_u64p)

    @staticmethod
    def f64(a) -> np.ndarray:
        return np.ascontiguousarray(a, dtype=np.float64)


L = _L()

ARENA = 100.0
N_OBSTACLES = 400
SIDE_MIN, SIDE_MAX = 0.5, 4.0
WORLD_SEED = 42
NOISE_SIGMA = 0.01
# order of the matcher parameters wherever they are stored as a flat vector
PARAM_KEYS = ("ndt_resolution", "search_angular_resolution", "search_angular_size",
              "search_linear_resolution", "search_linear_size", "laser_max_beams", "range_max")


def uniform(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.float64)
    L.lib.ndt2d_synth_uniform(seed, n, L.dptr(out))
    return out


def normal(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.float64)
    L.lib.ndt2d_synth_normal(seed, n, L.dptr(out))
    return out


def world(seed: int = WORLD_SEED, arena: float = ARENA, n_obstacles: int = N_OBSTACLES,
          side_min: float = SIDE_MIN, side_max: float = SIDE_MAX) -> np.ndarray:
    rects = np.empty((n_obstacles, 4), dtype=np.float64)
    L.check(L.lib.ndt2d_synth_world(seed, arena, n_obstacles, side_min, side_max, L.dptr(rects)),
            "synth_world")
    return rects


def free_start(rects: np.ndarray, length: float = 2.2, margin: float = 0.75):
    """First point of a fixed candidate list whose +x corridor of `length` m is
    clear of every obstacle by `margin` m (deterministic)."""
    for k in range(400):
        x0 = 40.0 + 1.7 * (k % 20)
        y0 = 40.0 + 1.3 * (k // 20)
        lo_x, hi_x, lo_y, hi_y = x0 - margin, x0 + length + margin, y0 - margin, y0 + margin
        hit = np.any((rects[:, 0] < hi_x) & (rects[:, 2] > lo_x) & (rects[:, 1] < hi_y) & (rects[:, 3] > lo_y))
        if not hit:
            return x0, y0
    return 50.0, 50.0


def scans(rects: np.ndarray, poses: np.ndarray, beams: int, range_max: float, seed: int,
          noise_sigma: float = NOISE_SIGMA, arena: float = ARENA):
    """Returns (pt_offsets[n+1] uint64, pts_xy[total, 2] float64), sensor frame."""
    poses = L.f64(poses).reshape(-1, 3)
    n = poses.shape[0]
    offs = np.empty(n + 1, dtype=np.uint64)
    pts = np.empty((max(n * beams, 1), 2), dtype=np.float64)
    rects = L.f64(rects)
    L.check(L.lib.ndt2d_synth_scans(L.dptr(rects), rects.shape[0], arena, L.dptr(poses), n, beams,
                                    range_max, noise_sigma, seed, L.u64ptr(offs), L.dptr(pts)),
            "synth_scans")
    return offs, np.ascontiguousarray(pts[: int(offs[n])])


@dataclass
class MatchWorkload:
    """One matcher workload: map scans + a query scan + matcher parameters."""
    name: str
    params: dict
    map_poses: np.ndarray
    map_offsets: np.ndarray
    map_points: np.ndarray
    query_pose: np.ndarray          # initial guess handed to matchScan
    query_points: np.ndarray
    true_pose: np.ndarray
    extra: dict = field(default_factory=dict)


def _line_poses(n: int, seed: int, rects: np.ndarray) -> np.ndarray:
    th = -0.05 + 0.1 * uniform(seed, n)
    x0, y0 = free_start(rects)
    poses = np.zeros((n, 3))
    poses[:, 0] = x0 + 0.2 * np.arange(n)
    poses[:, 1] = y0
    poses[:, 2] = th
    return poses


def config1(laser_max_beams: int = 360, beams: int = 360, n_map: int = 10) -> MatchWorkload:
    """Local match: 360-beam scan vs rolling NDT of 10 scans, 0.25 m cells,
    +-0.25 m @0.05, +-0.25 rad @0.0025 (BASELINE.json configs[0])."""
    rects = world()
    poses = _line_poses(n_map + 1, 11, rects)
    offs, pts = scans(rects, poses, beams, 10.0, seed=43)
    true_pose = poses[n_map].copy()
    guess = true_pose - np.array([0.12, -0.07, 0.06])
    params = dict(ndt_resolution=0.25, search_angular_resolution=0.0025, search_angular_size=0.25,
                  search_linear_resolution=0.05, search_linear_size=0.25,
                  laser_max_beams=laser_max_beams, range_max=10.0)
    q0, q1 = int(offs[n_map]), int(offs[n_map + 1])
    return MatchWorkload("config1_local_match", params, poses[:n_map].copy(), offs[: n_map + 1].copy(),
                         pts[: int(offs[n_map])].copy(), guess, pts[q0:q1].copy(), true_pose)


def config4(scale: float = 1.0) -> MatchWorkload:
    """Large correlative search: 1080-beam scan, +-2 m @0.01 m, +-pi @0.002 rad
    (BASELINE.json configs[3]).  scale < 1 shrinks the window for tests."""
    rects = world()
    poses = _line_poses(11, 11, rects)
    offs, pts = scans(rects, poses, 1080, 30.0, seed=143)
    true_pose = poses[10].copy()
    guess = true_pose - np.array([0.5, -0.3, 0.4])
    params = dict(ndt_resolution=0.25, search_angular_resolution=0.002,
                  search_angular_size=math.pi * scale, search_linear_resolution=0.01,
                  search_linear_size=2.0 * scale, laser_max_beams=1080, range_max=30.0)
    q0, q1 = int(offs[10]), int(offs[11])
    return MatchWorkload("config4_large_search", params, poses[:10].copy(), offs[:11].copy(),
                         pts[: int(offs[10])].copy(), guess, pts[q0:q1].copy(), true_pose)


def config4_dense(scale: float = 1.0) -> MatchWorkload:
    """The large search of config4 in a CLUTTERED, short-range setting: 10,000 small obstacles in
    the arena, range_max 5 m, 0.5 m NDT cells.  About 36 % of the (candidate, scan point) pairs land
    in an occupied cell (config4: 3.4 %), so the search kernel's sparsity shortcuts buy little here:
    the floor of the design's candidates/s (bench.py other_workloads)."""
    rects = world(seed=7, n_obstacles=10000, side_min=0.2, side_max=0.8)
    x0, y0 = free_start(rects, margin=0.3)
    poses = np.zeros((11, 3))
    poses[:, 0] = x0 + 0.2 * np.arange(11)
    poses[:, 1] = y0
    poses[:, 2] = -0.05 + 0.1 * uniform(11, 11)
    offs, pts = scans(rects, poses, 1080, 5.0, seed=143)
    true_pose = poses[10].copy()
    guess = true_pose - np.array([0.5, -0.3, 0.4])
    params = dict(ndt_resolution=0.5, search_angular_resolution=0.002,
                  search_angular_size=math.pi * scale, search_linear_resolution=0.01,
                  search_linear_size=2.0 * scale, laser_max_beams=1080, range_max=5.0)
    q0, q1 = int(offs[10]), int(offs[11])
    return MatchWorkload("config4_dense_clutter", params, poses[:10].copy(), offs[:11].copy(),
                         pts[: int(offs[10])].copy(), guess, pts[q0:q1].copy(), true_pose)


@dataclass
class FilterWorkload:
    name: str
    params: dict
    map_poses: np.ndarray
    map_offsets: np.ndarray
    map_points: np.ndarray
    scan_points: np.ndarray
    true_pose: np.ndarray
    particles: np.ndarray
    min_particles: int
    max_particles: int
    kld_err: float = 0.01
    kld_z: float = 2.3


def config2(n_side: int = 50, n_particles: int = 5000) -> FilterWorkload:
    """Particle-filter localisation: 5,000 particles x 360 beams against a global
    NDT of the 100x100 m map built from a 50x50 lattice of scans
    (BASELINE.json configs[1])."""
    rects = world()
    ij = np.stack(np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij"), -1).reshape(-1, 2)
    poses = np.zeros((ij.shape[0], 3))
    step = 100.0 / n_side
    poses[:, 0] = 0.5 * step + step * ij[:, 0]
    poses[:, 1] = 0.5 * step + step * ij[:, 1]
    poses[:, 2] = -math.pi + 2.0 * math.pi * uniform(5, ij.shape[0])
    offs, pts = scans(rects, poses, 360, 10.0, seed=1043)
    true_pose = np.array([42.3, 57.1, 0.7])
    _, scan_pts = scans(rects, true_pose[None, :], 360, 10.0, seed=77)
    z = normal(7, 3 * n_particles).reshape(-1, 3)
    particles = true_pose[None, :] + z * np.array([0.5, 0.5, 0.2])
    params = dict(ndt_resolution=0.25, search_angular_resolution=0.0025, search_angular_size=0.1,
                  search_linear_resolution=0.005, search_linear_size=0.05, laser_max_beams=360,
                  range_max=10.0)
    return FilterWorkload("config2_particle_filter", params, poses, offs, pts, scan_pts, true_pose,
                          np.ascontiguousarray(particles), 500, n_particles)


@dataclass
class BatchWorkload:
    name: str
    params: dict
    job_scan_offsets: np.ndarray
    map_poses: np.ndarray
    map_offsets: np.ndarray
    map_points: np.ndarray
    query_poses: np.ndarray
    query_offsets: np.ndarray
    query_points: np.ndarray


def config3(n_jobs: int = 50) -> BatchWorkload:
    """Global loop closure: one new scan matched against n_jobs candidate
    windows of 2 scans each (ndt_mapper.cpp:628-635), config-1 search window
    (BASELINE.json configs[2])."""
    rects = world()
    cx = 10.0 + 80.0 * uniform(301, n_jobs)
    cy = 10.0 + 80.0 * uniform(302, n_jobs)
    cth = -math.pi + 2.0 * math.pi * uniform(303, n_jobs)
    map_poses = np.zeros((2 * n_jobs, 3))
    map_poses[0::2] = np.stack([cx, cy, cth], 1)
    map_poses[1::2] = np.stack([cx + 0.2 * np.cos(cth), cy + 0.2 * np.sin(cth), cth + 0.02], 1)
    map_offsets, map_points = scans(rects, map_poses, 360, 10.0, seed=2043)
    true_q = np.stack([cx + 0.1 * np.cos(cth), cy + 0.1 * np.sin(cth), cth + 0.01], 1)
    q_offsets, q_points = scans(rects, true_q, 360, 10.0, seed=3043)
    err = np.stack([-0.2 + 0.4 * uniform(304, n_jobs), -0.2 + 0.4 * uniform(305, n_jobs),
                    -0.2 + 0.4 * uniform(306, n_jobs)], 1)
    params = dict(ndt_resolution=0.25, search_angular_resolution=0.0025, search_angular_size=0.25,
                  search_linear_resolution=0.05, search_linear_size=0.25, laser_max_beams=360,
                  range_max=10.0)
    return BatchWorkload("config3_loop_closure", params,
                         np.arange(0, 2 * n_jobs + 1, 2, dtype=np.uint64), map_poses, map_offsets,
                         map_points, np.ascontiguousarray(true_q + err), q_offsets, q_points)


@dataclass
class BuildWorkload:
    name: str
    params: dict
    poses: np.ndarray
    offsets: np.ndarray
    points: np.ndarray


def config5(n_scans: int = 20000, beams: int = 360) -> BuildWorkload:
    """NdtModel build throughput: 20,000 scans binned into a 0.1 m grid
    (BASELINE.json configs[4])."""
    rects = world()
    poses = np.stack([100.0 * uniform(501, n_scans), 100.0 * uniform(502, n_scans),
                      -math.pi + 2.0 * math.pi * uniform(503, n_scans)], 1)
    offs, pts = scans(rects, poses, beams, 10.0, seed=5043)
    params = dict(ndt_resolution=0.1, search_angular_resolution=0.0025, search_angular_size=0.1,
                  search_linear_resolution=0.005, search_linear_size=0.05, laser_max_beams=100,
                  range_max=10.0)
    return BuildWorkload("config5_build", params, np.ascontiguousarray(poses), offs, pts)
