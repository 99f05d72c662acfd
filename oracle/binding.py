"""ctypes access to the parity oracle -- TEST INFRASTRUCTURE ONLY.

Two libraries with mirrored entry points:
  * oracle/_build/libndt2d_oracle.so  (prefix orc_)  our C restatement
  * oracle/_ref/libndt2d_ref.so       (prefix ref_)  the reference's own sources
    compiled in place (only buildable where /root/reference exists; the built
    .so travels to the GPU box)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "_build" / "libndt2d_oracle.so"
REF_SO = HERE / "_ref" / "libndt2d_ref.so"

_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p


def build(quiet: bool = True) -> None:
    """make -C oracle : the restatement always, oracle/_ref when the reference tree exists."""
    subprocess.run(["make", "-C", str(HERE)] + (["-s"] if quiet else []), check=True,
                   capture_output=quiet)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _d(a):
    return a.ctypes.data_as(_dp)


class OracleLib:
    """Uniform wrapper over either library."""

    def __init__(self, path: Path, prefix: str):
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(str(path))
        p = prefix
        sig = {
            "cell_new": (_vp, []), "cell_free": (None, [_vp]),
            "cell_add_point": (None, [_vp, C.c_double, C.c_double]),
            "cell_compute": (None, [_vp]), "cell_score": (C.c_double, [_vp, C.c_double, C.c_double]),
            "cell_get": (None, [_vp, _dp]),
            "ndt_create": (_vp, [C.c_double] * 5), "ndt_destroy": (None, [_vp]),
            "ndt_get_index": (C.c_int, [_vp, C.c_double, C.c_double]),
            "ndt_add_scan": (None, [_vp, _dp, _dp, C.c_size_t]), "ndt_compute": (None, [_vp]),
            "ndt_likelihood_point": (C.c_double, [_vp, C.c_double, C.c_double]),
            "ndt_likelihood_points": (C.c_double, [_vp, _dp, C.c_size_t]),
            "ndt_likelihood_scan": (C.c_double, [_vp, _dp, _dp, C.c_size_t]),
            "ndt_grid": (None, [_vp, _dp]), "ndt_dump_cells": (None, [_vp, _dp]),
            "matcher_create": (_vp, [C.c_double] * 5 + [C.c_int, C.c_double]),
            "matcher_destroy": (None, [_vp]), "matcher_reset": (None, [_vp]),
            "matcher_ndt": (_vp, [_vp]),
            "matcher_add_scans": (None, [_vp, C.c_size_t, _dp, _u64p, _dp]),
            "matcher_match_scan": (C.c_double, [_vp, _dp, _dp, C.c_size_t, _dp,
                                                C.POINTER(C.c_int), _dp, _dp]),
            "matcher_score_points": (C.c_double, [_vp, _dp, C.c_size_t, _dp]),
            "normalize_angle": (C.c_double, [C.c_double]),
            "shortest_angular_distance": (C.c_double, [C.c_double, C.c_double]),
            "kd_leaf_counts": (None, [_dp, C.c_size_t, _dp, _u64p]),
            "scan_barycenter": (None, [_dp, _dp, C.c_size_t, _dp]),
            "occ_create": (_vp, [C.c_double, C.c_double]), "occ_destroy": (None, [_vp]),
            "occ_render": (C.c_size_t, [_vp, C.c_size_t, _dp, _u64p, _dp, _dp, C.POINTER(C.c_int8),
                                        C.c_size_t]),
        }
        if prefix == "orc_":
            sig.update({
                "loop_values": (C.c_size_t, [C.c_double, C.c_double, _dp, C.c_size_t]),
                "matcher_match_scan_window": (
                    C.c_double, [_vp, _dp, _dp, C.c_size_t, _dp, C.POINTER(C.c_int), _dp, _dp,
                                 C.c_size_t, C.c_size_t, _u64p]),
                "find_nearest": (C.c_size_t, [_dp, C.c_size_t, C.c_longlong, _dp, C.c_double, _u64p, _dp]),
                "laser_to_points": (C.c_size_t, [C.POINTER(C.c_float), C.c_size_t, C.c_float, C.c_float,
                                                 C.c_double, _dp, _dp, C.c_int, _dp]),
                "matcher_partial": (None, [_vp, _dp, _dp, C.c_size_t, C.c_size_t, C.c_size_t, _dp]),
                "pf_measure": (None, [_vp, _dp, C.c_size_t, _dp, C.c_size_t, _dp]),
                "pf_update_statistics": (None, [_dp, _dp, C.c_size_t, _dp, _dp]),
                "pf_resample": (C.c_size_t, [_dp, _dp, C.c_size_t, C.c_size_t, C.c_size_t,
                                             C.c_double, C.c_double, _dp, _dp, _dp, _u64p]),
            })
        else:
            sig.update({
                "matcher_defaults": (None, [_dp]),
                "matcher_score_scan": (C.c_double, [_vp, _dp, _dp, C.c_size_t]),
                "matcher_candidate_count": (C.c_uint64, [_vp]),
                "pf_create": (_vp, [C.c_size_t, C.c_size_t, _dp]), "pf_destroy": (None, [_vp]),
                "pf_seed": (None, [_vp, C.c_uint32, C.c_uint32]),
                "pf_set": (None, [_vp, _dp, _dp, C.c_size_t]), "pf_size": (C.c_size_t, [_vp]),
                "pf_get": (None, [_vp, _dp, _dp]), "pf_stats": (None, [_vp, _dp, _dp]),
                "pf_set_cov": (None, [_vp, _dp]), "pf_update_statistics": (None, [_vp]),
                "pf_init": (None, [_vp] + [C.c_double] * 6),
                "pf_update": (None, [_vp] + [C.c_double] * 3),
                "pf_measure": (None, [_vp, _vp, _dp, C.c_size_t]),
                "pf_resample": (None, [_vp, C.c_double, C.c_double]),
                "canonical_uniforms": (None, [C.c_uint32, C.c_size_t, _dp]),
            })
        for name, (res, args) in sig.items():
            fn = getattr(self.lib, p + name)
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    # ---- conveniences shared by both libraries --------------------------
    def new_matcher(self, params: dict) -> "Matcher":
        return Matcher(self, params)

    def cell_state(self, cell) -> np.ndarray:
        out = np.zeros(16)
        self.cell_get(cell, _d(out))
        return out

    def kd_counts(self, poses, sizes=(0.5, 0.5, 0.2671)) -> np.ndarray:
        poses = _f64(poses).reshape(-1, 3)
        sizes = _f64(sizes)
        out = np.zeros(poses.shape[0], dtype=np.uint64)
        self.kd_leaf_counts(_d(poses), poses.shape[0], _d(sizes), out.ctypes.data_as(_u64p))
        return out


class Matcher:
    """ScanMatcherNDT of either oracle library."""

    def __init__(self, olib: OracleLib, params: dict):
        self.o = olib
        self.params = dict(params)
        self.h = olib.matcher_create(
            params["ndt_resolution"], params["search_angular_resolution"],
            params["search_angular_size"], params["search_linear_resolution"],
            params["search_linear_size"], int(params["laser_max_beams"]), params["range_max"])

    def close(self):
        if self.h:
            self.o.matcher_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self.o.matcher_reset(self.h)

    def add_scans(self, poses, offsets, points):
        poses = _f64(poses).reshape(-1, 3)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        points = _f64(points).reshape(-1, 2)
        self.o.matcher_add_scans(self.h, poses.shape[0], _d(poses), offsets.ctypes.data_as(_u64p),
                                 _d(points))

    def grid(self):
        info = np.zeros(5)
        self.o.ndt_grid(self.o.matcher_ndt(self.h), _d(info))
        return int(info[0]), int(info[1]), info[2], info[3], info[4]

    def dump_cells(self) -> np.ndarray:
        sx, sy, *_ = self.grid()
        out = np.zeros((sx * sy, 16))
        self.o.ndt_dump_cells(self.o.matcher_ndt(self.h), _d(out))
        return out

    def get_index(self, x, y) -> int:
        return self.o.ndt_get_index(self.o.matcher_ndt(self.h), x, y)

    def match_scan(self, pose, points, want_scores: bool = False):
        """-> (score, delta[3], written, cov[3,3], scores or None)"""
        pose = _f64(pose).reshape(3)
        points = _f64(points).reshape(-1, 2)
        delta, cov = np.zeros(3), np.full((3, 3), np.nan)
        written = C.c_int(0)
        scores = None
        if want_scores:
            assert self.o.prefix == "orc_", "per-candidate scores only from the restatement"
            na = self.o.loop_values(self.params["search_angular_size"],
                                    self.params["search_angular_resolution"], None, 0)
            nl = self.o.loop_values(self.params["search_linear_size"],
                                    self.params["search_linear_resolution"], None, 0)
            scores = np.zeros(na * nl * nl)
        s = self.o.matcher_match_scan(self.h, _d(pose), _d(points), points.shape[0], _d(delta),
                                      C.byref(written), _d(cov), _d(scores) if want_scores else None)
        if want_scores:
            scores = scores.reshape(na, nl, nl)
        return s, delta, bool(written.value), cov, scores

    def match_scan_window(self, pose, points, theta_lo: int, theta_hi: int):
        """Bounded sample of a large search (oracle only): -> (score, n_candidates)."""
        pose = _f64(pose).reshape(3)
        points = _f64(points).reshape(-1, 2)
        delta, cov = np.zeros(3), np.zeros((3, 3))
        written, ncand = C.c_int(0), C.c_uint64(0)
        s = self.o.matcher_match_scan_window(self.h, _d(pose), _d(points), points.shape[0],
                                             _d(delta), C.byref(written), _d(cov), None,
                                             theta_lo, theta_hi, C.byref(ncand))
        return s, int(ncand.value), delta, bool(written.value), cov

    def partial(self, pose, points, theta_lo: int, theta_hi: int) -> np.ndarray:
        """16-double partial record of the theta range (oracle only)."""
        pose = _f64(pose).reshape(3)
        points = _f64(points).reshape(-1, 2)
        out = np.zeros(16)
        self.o.matcher_partial(self.h, _d(pose), _d(points), points.shape[0], theta_lo, theta_hi, _d(out))
        return out

    def score_points(self, points, pose) -> float:
        pose = _f64(pose).reshape(3)
        points = _f64(points).reshape(-1, 2)
        return self.o.matcher_score_points(self.h, _d(points), points.shape[0], _d(pose))


class OccupancyGrid:
    """ndt_2d::OccupancyGrid of either oracle library (state persists across getMsg calls)."""

    def __init__(self, olib: OracleLib, resolution: float, occ_thresh: float):
        self.o = olib
        self.h = olib.occ_create(resolution, occ_thresh)

    def __del__(self):
        try:
            if self.h:
                self.o.occ_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def get_msg(self, poses, offsets, points):
        """-> (info dict, data int8[height, width])"""
        poses = _f64(poses).reshape(-1, 3)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        points = _f64(points).reshape(-1, 2)
        info = np.zeros(5)
        cap = 1 << 16
        while True:
            data = np.zeros(cap, dtype=np.int8)
            # NB: a second call with the same number of scans does not update the bounds, like the
            # reference; calling again only to size the buffer is therefore harmless
            n = self.o.occ_render(self.h, poses.shape[0], _d(poses), offsets.ctypes.data_as(_u64p),
                                  _d(points), _d(info), data.ctypes.data_as(C.POINTER(C.c_int8)), cap)
            if n <= cap:
                break
            cap = int(n)
        w, h = int(info[0]), int(info[1])
        return dict(width=w, height=h, origin_x=info[2], origin_y=info[3], resolution=info[4]), \
            data[:w * h].reshape(h, w).copy()


def load_oracle() -> OracleLib:
    if not ORACLE_SO.exists():
        build()
    return OracleLib(ORACLE_SO, "orc_")


def load_ref():
    """The compiled reference, or None when it was never built."""
    if not REF_SO.exists():
        return None
    return OracleLib(REF_SO, "ref_")


def laser_to_points(o: OracleLib, ranges, angle_min, angle_increment, range_max, laser_tf, translation,
                    inverted: bool) -> np.ndarray:
    """Mapper::laserCallback's LaserScan -> points conversion (oracle restatement only)."""
    r = np.ascontiguousarray(ranges, dtype=np.float32)
    out = np.zeros((max(r.shape[0], 1), 2))
    lt, tr = _f64(laser_tf).reshape(3), _f64(translation).reshape(3)
    n = o.laser_to_points(r.ctypes.data_as(C.POINTER(C.c_float)), r.shape[0], angle_min, angle_increment,
                          range_max, _d(lt), _d(tr), int(bool(inverted)), _d(out))
    return out[:n].copy()


def scan_barycenter(o: OracleLib, pose3, points) -> np.ndarray:
    """Scan::getBarycenterPose (scan.cpp:55-59, 72-91)."""
    pose3, pts = _f64(pose3).reshape(3), _f64(points).reshape(-1, 2)
    out = np.zeros(3)
    o.scan_barycenter(_d(pose3), _d(pts), pts.shape[0], _d(out))
    return out


def find_nearest(o: OracleLib, scan_xy, query_xy, dist: float, limit_scan_index: int = -1):
    """Graph::findNearest (graph.cpp:167-189), oracle restatement only -> (indices, squared distances)."""
    xy, q = _f64(scan_xy).reshape(-1, 2), _f64(query_xy).reshape(2)
    idx = np.zeros(max(xy.shape[0], 1), dtype=np.uint64)
    d2 = np.zeros(max(xy.shape[0], 1))
    n = o.find_nearest(_d(xy), xy.shape[0], int(limit_scan_index), _d(q), float(dist),
                       idx.ctypes.data_as(_u64p), _d(d2))
    return idx[:n].copy(), d2[:n].copy()


# ---- oracle-only particle-filter helpers ---------------------------------
def pf_measure(o: OracleLib, matcher: Matcher, particles, points) -> np.ndarray:
    particles = _f64(particles).reshape(-1, 3)
    points = _f64(points).reshape(-1, 2)
    w = np.zeros(particles.shape[0])
    o.pf_measure(matcher.h, _d(particles), particles.shape[0], _d(points), points.shape[0], _d(w))
    return w


def pf_update_statistics(o: OracleLib, particles, weights, cov_prev):
    """-> (normalised weights, mean[3], cov[3,3])"""
    particles = _f64(particles).reshape(-1, 3)
    w = _f64(weights).copy()
    mean = np.zeros(3)
    cov = _f64(cov_prev).reshape(3, 3).copy()
    o.pf_update_statistics(_d(particles), _d(w), particles.shape[0], _d(mean), _d(cov))
    return w, mean, cov


def pf_resample(o: OracleLib, particles, weights, min_p, max_p, kld_err, kld_z, uniforms):
    """-> (particles[N,3], weights[N], indices[N])"""
    particles = _f64(particles).reshape(-1, 3)
    weights = _f64(weights)
    uniforms = _f64(uniforms)
    assert uniforms.shape[0] >= max_p
    outp, outw = np.zeros((max_p, 3)), np.zeros(max_p)
    idx = np.zeros(max_p, dtype=np.uint64)
    n = o.pf_resample(_d(particles), _d(weights), particles.shape[0], min_p, max_p, kld_err, kld_z,
                      _d(uniforms), _d(outp), _d(outw), idx.ctypes.data_as(_u64p))
    return outp[:n].copy(), outw[:n].copy(), idx[:n].copy()
