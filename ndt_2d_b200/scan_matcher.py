"""Host-side mirror of the reference's scan-matcher plugin interface over the C ABI.

Same names, argument meaning and error behaviour as ndt_2d::ScanMatcher /
ndt_2d::ScanMatcherNDT (include/ndt_2d/scan_matcher.hpp:42-91,
src/scan_matcher_ndt.cpp:35-183) so that parity tests read like calls into the
reference.  The C++ drop-in class lives in ndt_2d_b200/plugin/; this module exists
for tests and bench.py.  All computation happens in libndt2d_b200.so.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _lib as L


@dataclass
class Pose2d:
    """include/ndt_2d/pose_2d.hpp:35-55"""
    x: float = 0.0
    y: float = 0.0
    theta: float = 0.0

    def as_array(self) -> np.ndarray:
        return np.array([self.x, self.y, self.theta], dtype=np.float64)


@dataclass
class Scan:
    """include/ndt_2d/scan.hpp:40-90: id, pose (map frame), points (sensor frame)."""
    id: int = 0
    pose: Pose2d = field(default_factory=Pose2d)
    points: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))

    def getPose(self) -> Pose2d:
        return self.pose

    def setPose(self, pose: Pose2d) -> None:
        self.pose = pose

    def getPoints(self) -> np.ndarray:
        return self.points

    def setPoints(self, points) -> None:
        self.points = L.f64(points).reshape(-1, 2)

    def getBarycenterPose(self) -> Pose2d:
        """scan.cpp:55-59, 72-91: pose moved to the mean of the points (sums in point order)."""
        p = self.pose
        x, y = p.x, p.y
        pts = L.f64(self.points).reshape(-1, 2)
        if pts.shape[0]:
            c, s = math.cos(p.theta), math.sin(p.theta)
            cx = np.add.accumulate(c * pts[:, 0] - s * pts[:, 1])[-1]     # sequential, like the loop
            cy = np.add.accumulate(s * pts[:, 0] + c * pts[:, 1])[-1]
            x += cx / pts.shape[0]
            y += cy / pts.shape[0]
        return Pose2d(float(x), float(y), p.theta)


class ParameterNode:
    """Stand-in for the rclcpp::Node the plugin reads its parameters from:
    declare_parameter(name, default) returns the override if one was given."""

    def __init__(self, overrides: Optional[dict] = None):
        self.overrides = dict(overrides or {})
        self.declared = {}

    def declare_parameter(self, name: str, default):
        value = self.overrides.get(name, default)
        self.declared[name] = value
        return type(default)(value)


def _pose3(pose) -> np.ndarray:
    if isinstance(pose, Pose2d):
        return pose.as_array()
    return L.f64(pose).reshape(3)


class OccupancyGrid:
    """Mirror of ndt_2d::OccupancyGrid (occupancy_grid.hpp:40-68) on the device."""

    def __init__(self, resolution: float, occ_thresh: float, device: int = -1):
        self._h = C.c_void_p()
        L.check(L.lib.ndt2d_occupancy_create(resolution, occ_thresh, device, C.byref(self._h)),
                "ndt2d_occupancy_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib.ndt2d_occupancy_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def getMsg(self, poses, offsets, points, fetch: bool = True):
        """-> (info dict, data int8[height, width])  (occupancy_grid.cpp:47-152)"""
        p = L.f64(poses).reshape(-1, 3)
        o = np.ascontiguousarray(offsets, dtype=np.uint64)
        pts = L.f64(points).reshape(-1, 2)
        info = np.zeros(5)
        L.check(L.lib.ndt2d_occupancy_render(self._h, p.shape[0], L.dptr(p), L.u64ptr(o), L.dptr(pts),
                                             L.dptr(info)), "ndt2d_occupancy_render")
        w, h = int(info[0]), int(info[1])
        meta = dict(width=w, height=h, origin_x=info[2], origin_y=info[3], resolution=info[4])
        if not fetch:
            return meta, None
        data = np.zeros(max(w * h, 1), dtype=np.int8)
        L.check(L.lib.ndt2d_occupancy_fetch(self._h, data.ctypes.data_as(C.POINTER(C.c_int8)), data.shape[0]),
                "ndt2d_occupancy_fetch")
        return meta, data[:w * h].reshape(h, w)


def laser_to_points(ranges, angle_min: float, angle_increment: float, range_max: float, laser_tf,
                    translation, inverted: bool = False, device: int = -1) -> np.ndarray:
    """LaserScan -> ndt_2d::Scan points on the device (Mapper::laserCallback, ndt_mapper.cpp:385-453)."""
    r = np.ascontiguousarray(ranges, dtype=np.float32)
    out = np.zeros((max(r.shape[0], 1), 2))
    lt, tr = _pose3(laser_tf), _pose3(translation)
    n = C.c_size_t(0)
    L.check(L.lib.ndt2d_laser_to_points(device, r.ctypes.data_as(C.POINTER(C.c_float)), r.shape[0],
                                        angle_min, angle_increment, range_max, L.dptr(lt), L.dptr(tr),
                                        int(bool(inverted)), L.dptr(out), C.byref(n)), "ndt2d_laser_to_points")
    return out[:n.value].copy()


def find_nearest(scan_xy, query_xy, dist: float, limit_scan_index: int = -1, device: int = -1,
                 return_distances: bool = False):
    """Graph::findNearest (graph.cpp:167-189) on the device: indices of the scans whose position
    (pose or barycenter x, y) is closer than `dist` -- a SQUARED radius, as nanoflann's
    radiusSearch takes it -- to the query, nearest first; only scans [0, limit) when limit > 0."""
    xy = np.ascontiguousarray(scan_xy, dtype=np.float64).reshape(-1, 2)
    q = np.ascontiguousarray(query_xy, dtype=np.float64).reshape(2)
    cap = max(xy.shape[0], 1)
    idx = np.zeros(cap, dtype=np.uint64)
    d2 = np.zeros(cap)
    n = C.c_size_t(0)
    L.check(L.lib.ndt2d_find_nearest(device, L.dptr(xy), xy.shape[0], int(limit_scan_index), L.dptr(q),
                                     float(dist), idx.ctypes.data_as(C.POINTER(C.c_uint64)), L.dptr(d2),
                                     cap, C.byref(n)), "ndt2d_find_nearest")
    k = min(n.value, cap)
    return (idx[:k].copy(), d2[:k].copy()) if return_distances else idx[:k].copy()


def graph_find_nearest(scans, scan: Scan, dist: float, limit_scan_index: int = -1,
                       use_barycenter: bool = False, device: int = -1) -> np.ndarray:
    """Graph::findNearest(scan, dist, limit_scan_index) over a list of Scan (graph.cpp:167-189)."""
    pick = (lambda s: s.getBarycenterPose()) if use_barycenter else (lambda s: s.getPose())
    xy = np.array([[pick(s).x, pick(s).y] for s in scans], dtype=np.float64).reshape(-1, 2)
    q = pick(scan)
    return find_nearest(xy, [q.x, q.y], dist, limit_scan_index, device)


class ScanMatcherNDT:
    """B200 backend behind the reference's ScanMatcher interface."""

    def __init__(self, device: int = -1, stream: int = 0, kernel_variant: int = 0,
                 devices: Optional[Sequence[int]] = None):
        """devices: several GPUs of this process behind ONE handle (ndt2d_params.n_devices): large
        searches are theta-sliced over them, everything else runs on devices[0]."""
        self._h = C.c_void_p()
        self._device = device
        self._stream = stream
        self._variant = kernel_variant
        self._devices = list(devices) if devices else []
        self.params: Optional[L.Params] = None

    # ---- ScanMatcher::initialize (scan_matcher_ndt.cpp:35-47)
    def initialize(self, name: str, node: ParameterNode, range_max: float) -> None:
        p = L.Params()
        L.lib.ndt2d_default_params(C.byref(p))
        p.ndt_resolution = node.declare_parameter(name + ".ndt_resolution", 0.25)
        p.search_angular_resolution = node.declare_parameter(name + ".search_angular_resolution", 0.0025)
        p.search_angular_size = node.declare_parameter(name + ".search_angular_size", 0.1)
        p.search_linear_resolution = node.declare_parameter(name + ".search_linear_resolution", 0.005)
        p.search_linear_size = node.declare_parameter(name + ".search_linear_size", 0.05)
        p.laser_max_beams = node.declare_parameter(name + ".laser_max_beams", 100)
        p.range_max = float(range_max)
        p.device = self._device
        p.stream = self._stream or None
        p.kernel_variant = self._variant
        p.n_devices = len(self._devices)
        for k, d in enumerate(self._devices):
            p.devices[k] = int(d)
        self._destroy()
        L.check(L.lib.ndt2d_matcher_create(C.byref(p), C.byref(self._h)), "ndt2d_matcher_create")
        self.params = p

    @classmethod
    def from_params(cls, params: dict, name: str = "m", **kw) -> "ScanMatcherNDT":
        node = ParameterNode({f"{name}.{k}": v for k, v in params.items() if k != "range_max"})
        m = cls(**kw)
        m.initialize(name, node, params.get("range_max", 0.0))
        return m

    def _destroy(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib.ndt2d_matcher_destroy(self._h)
            self._h = C.c_void_p()

    def close(self):
        self._destroy()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h.value:
            raise RuntimeError("ScanMatcherNDT.initialize() has not been called")
        return self._h

    # ---- ScanMatcher::addScans (scan_matcher_ndt.cpp:49-74)
    def addScans(self, scans: Sequence[Scan]) -> None:
        scans = list(scans)
        poses = np.array([[s.pose.x, s.pose.y, s.pose.theta] for s in scans], dtype=np.float64).reshape(-1, 3)
        counts = [np.asarray(s.points).reshape(-1, 2).shape[0] for s in scans]
        offs = np.zeros(len(scans) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(counts)
        pts = (np.concatenate([L.f64(s.points).reshape(-1, 2) for s in scans], 0)
               if scans else np.zeros((0, 2)))
        self.add_scans_raw(poses, offs, pts)

    def add_scans_raw(self, poses: np.ndarray, offsets: np.ndarray, points: np.ndarray) -> None:
        poses = L.f64(poses).reshape(-1, 3)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        points = L.f64(points).reshape(-1, 2)
        L.check(L.lib.ndt2d_matcher_add_scans(self.handle, poses.shape[0], L.dptr(poses),
                                              L.u64ptr(offsets), L.dptr(points)),
                "ndt2d_matcher_add_scans")

    # ---- ScanMatcher::matchScan (scan_matcher_ndt.cpp:76-149)
    def matchScan(self, scan: Scan, pose: Optional[Pose2d] = None,
                  covariance: Optional[np.ndarray] = None):
        """Returns (score, pose, covariance).  `pose` (the correction delta) is
        only modified when a candidate scored below zero and `covariance` only
        when a map exists -- exactly the reference's in/out semantics."""
        pose = pose if pose is not None else Pose2d()
        cov = covariance if covariance is not None else np.zeros((3, 3))
        score, delta, written, cov_out, status = self.match_scan_raw(scan.pose, scan.points)
        if status == L.ERR_NO_MAP:
            return 0.0, pose, cov
        if written:
            pose.x, pose.y, pose.theta = (float(v) for v in delta)
        cov[...] = cov_out
        return score, pose, cov

    def match_scan_raw(self, pose, points):
        pose3 = _pose3(pose)
        pts = L.f64(points).reshape(-1, 2)
        delta = np.zeros(3)
        cov = np.zeros((3, 3))
        written = C.c_int(0)
        score = C.c_double(0.0)
        st = L.check(L.lib.ndt2d_matcher_match_scan(self.handle, L.dptr(pose3), L.dptr(pts),
                                                    pts.shape[0], L.dptr(delta), C.byref(written),
                                                    L.dptr(cov), C.byref(score)),
                     "ndt2d_matcher_match_scan", allow=(L.ERR_NO_MAP,))
        return score.value, delta, bool(written.value), cov, st

    # ---- ScanMatcher::scoreScan / scorePoints (scan_matcher_ndt.cpp:151-178)
    def scoreScan(self, scan: Scan) -> float:
        return self.scorePoints(scan.points, scan.pose)

    def scorePoints(self, points, pose) -> float:
        pts = L.f64(points).reshape(-1, 2)
        pose3 = _pose3(pose)
        out = C.c_double(0.0)
        L.check(L.lib.ndt2d_matcher_score_points(self.handle, L.dptr(pts), pts.shape[0],
                                                 L.dptr(pose3), C.byref(out)),
                "ndt2d_matcher_score_points", allow=(L.ERR_NO_MAP,))
        return out.value

    def scorePoses(self, points, poses) -> np.ndarray:
        pts = L.f64(points).reshape(-1, 2)
        poses = L.f64(poses).reshape(-1, 3)
        out = np.zeros(poses.shape[0])
        L.check(L.lib.ndt2d_matcher_score_poses(self.handle, L.dptr(pts), pts.shape[0],
                                                L.dptr(poses), poses.shape[0], L.dptr(out)),
                "ndt2d_matcher_score_poses", allow=(L.ERR_NO_MAP,))
        return out

    def likelihoodScan(self, scan: Scan) -> float:
        """NDT::likelihood(const ScanPtr&) (ndt_model.cpp:189-201)."""
        pts = L.f64(scan.points).reshape(-1, 2)
        pose3 = _pose3(scan.pose)
        out = C.c_double(0.0)
        L.check(L.lib.ndt2d_matcher_likelihood_scan(self.handle, L.dptr(pose3), L.dptr(pts),
                                                    pts.shape[0], C.byref(out)),
                "ndt2d_matcher_likelihood_scan", allow=(L.ERR_NO_MAP,))
        return out.value

    # ---- ScanMatcher::reset (scan_matcher_ndt.cpp:180-183)
    def reset(self) -> None:
        L.check(L.lib.ndt2d_matcher_reset(self.handle), "ndt2d_matcher_reset")

    # ---- batch (loop closure, ndt_mapper.cpp:619-671)
    def match_scan_batch(self, job_scan_offsets, map_poses, map_offsets, map_points, query_poses,
                         query_offsets, query_points):
        jso = np.ascontiguousarray(job_scan_offsets, dtype=np.uint64)
        n_jobs = jso.shape[0] - 1
        mp, mo, mpts = L.f64(map_poses).reshape(-1, 3), np.ascontiguousarray(map_offsets, dtype=np.uint64), L.f64(map_points).reshape(-1, 2)
        qp, qo, qpts = L.f64(query_poses).reshape(-1, 3), np.ascontiguousarray(query_offsets, dtype=np.uint64), L.f64(query_points).reshape(-1, 2)
        delta = np.zeros((n_jobs, 3))
        written = np.zeros(n_jobs, dtype=np.int32)
        cov = np.zeros((n_jobs, 3, 3))
        score = np.zeros(n_jobs)
        L.check(L.lib.ndt2d_matcher_match_scan_batch(
            self.handle, n_jobs, L.u64ptr(jso), L.dptr(mp), L.u64ptr(mo), L.dptr(mpts), L.dptr(qp),
            L.u64ptr(qo), L.dptr(qpts), L.dptr(delta), written.ctypes.data_as(C.POINTER(C.c_int)),
            L.dptr(cov), L.dptr(score)), "ndt2d_matcher_match_scan_batch")
        return score, delta, written.astype(bool), cov

    def close_loop(self, scan_poses, scan_offsets, scan_points, candidates, rolling: int,
                   search_limit: int, typical_response: float, query_pose, query_points):
        """Mapper::loopClosureThread's inner loop for one new scan (ndt_mapper.cpp:619-671).
        -> (final query pose[3], list of dicts per processed candidate, n_batches)"""
        sp = L.f64(scan_poses).reshape(-1, 3)
        so = np.ascontiguousarray(scan_offsets, dtype=np.uint64)
        spts = L.f64(scan_points).reshape(-1, 2)
        cand = np.ascontiguousarray(candidates, dtype=np.uint64)
        qp = _pose3(query_pose).copy()
        qpts = L.f64(query_points).reshape(-1, 2)
        # (a limit of 0 means no limit: the reference's size_t countdown wraps, ndt_mapper.cpp:619,671)
        cap = max(1, min(int(search_limit), cand.shape[0]) if int(search_limit) > 0 else cand.shape[0])
        oc = np.zeros(cap, dtype=np.uint64)
        osc, oacc = np.zeros(cap), np.zeros(cap, dtype=np.int32)
        opose, ocov = np.zeros((cap, 3)), np.zeros((cap, 3, 3))
        n, nb = C.c_size_t(0), C.c_size_t(0)
        L.check(L.lib.ndt2d_matcher_close_loop(
            self.handle, sp.shape[0], L.dptr(sp), L.u64ptr(so), L.dptr(spts), L.u64ptr(cand), cand.shape[0],
            rolling, search_limit, typical_response, L.dptr(qp), L.dptr(qpts), qpts.shape[0], L.u64ptr(oc),
            L.dptr(osc), oacc.ctypes.data_as(C.POINTER(C.c_int)), L.dptr(opose), L.dptr(ocov),
            C.byref(n), C.byref(nb)), "ndt2d_matcher_close_loop")
        out = [dict(candidate=int(oc[k]), score=float(osc[k]), accepted=bool(oacc[k]), pose=opose[k].copy(),
                    covariance=ocov[k].copy()) for k in range(n.value)]
        return qp, out, int(nb.value)

    # ---- staged / partial search
    def search_shape(self):
        na, nl = C.c_uint64(0), C.c_uint64(0)
        L.check(L.lib.ndt2d_matcher_search_shape(self.handle, C.byref(na), C.byref(nl)), "search_shape")
        return int(na.value), int(nl.value)

    def search_values(self):
        na, nl = self.search_shape()
        dth, dlin = np.zeros(na), np.zeros(nl)
        L.check(L.lib.ndt2d_matcher_search_values(self.handle, L.dptr(dth), L.dptr(dlin)), "search_values")
        return dth, dlin

    def stage_scan(self, pose, points) -> None:
        pose3 = _pose3(pose)
        pts = L.f64(points).reshape(-1, 2)
        L.check(L.lib.ndt2d_matcher_stage_scan(self.handle, L.dptr(pose3), L.dptr(pts), pts.shape[0]),
                "ndt2d_matcher_stage_scan")

    def search_staged(self, theta_begin: int, theta_end: int, d_partial: int = 0, stride: int = 1) -> None:
        L.check(L.lib.ndt2d_matcher_search_staged_strided(self.handle, theta_begin, theta_end, stride,
                                                          d_partial or None), "ndt2d_matcher_search_staged")

    # ---- fused cross-GPU exchange
    def exchange_init(self, world: int, rank: int) -> bytes:
        h = (C.c_ubyte * 64)()
        L.check(L.lib.ndt2d_matcher_exchange_init(self.handle, world, rank, h), "ndt2d_matcher_exchange_init")
        return bytes(h)

    def exchange_connect(self, handles) -> None:
        blob = b"".join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        L.check(L.lib.ndt2d_matcher_exchange_connect(self.handle, buf), "ndt2d_matcher_exchange_connect")

    def search_exchange(self, theta_begin: int, theta_end: int, stride: int, seq: int) -> None:
        L.check(L.lib.ndt2d_matcher_search_exchange(self.handle, theta_begin, theta_end, stride, seq),
                "ndt2d_matcher_search_exchange")

    def fetch_result(self):
        delta, cov = np.zeros(3), np.zeros((3, 3))
        written, score = C.c_int(0), C.c_double(0.0)
        L.check(L.lib.ndt2d_matcher_fetch_result(self.handle, L.dptr(delta), C.byref(written), L.dptr(cov),
                                                 C.byref(score)), "ndt2d_matcher_fetch_result")
        return float(score.value), delta, bool(written.value), cov

    def fetch_partial(self) -> np.ndarray:
        out = np.zeros(L.PARTIAL_DOUBLES)
        L.check(L.lib.ndt2d_matcher_fetch_partial(self.handle, L.dptr(out)), "ndt2d_matcher_fetch_partial")
        return out

    def combine_partials(self, partials: np.ndarray):
        partials = L.f64(partials).reshape(-1, L.PARTIAL_DOUBLES)
        delta, cov = np.zeros(3), np.zeros((3, 3))
        written, score = C.c_int(0), C.c_double(0.0)
        L.check(L.lib.ndt2d_combine_partials(self.handle, L.dptr(partials), partials.shape[0],
                                             L.dptr(delta), C.byref(written), L.dptr(cov),
                                             C.byref(score)), "ndt2d_combine_partials")
        return score.value, delta, bool(written.value), cov

    def combine_device(self, d_partials: int, n: int):
        delta, cov = np.zeros(3), np.zeros((3, 3))
        written, score = C.c_int(0), C.c_double(0.0)
        L.check(L.lib.ndt2d_matcher_combine_device(self.handle, d_partials, n, L.dptr(delta),
                                                   C.byref(written), L.dptr(cov), C.byref(score)),
                "ndt2d_matcher_combine_device")
        return score.value, delta, bool(written.value), cov

    # ---- parity / introspection
    def grid_info(self):
        info = np.zeros(5)
        L.check(L.lib.ndt2d_matcher_grid_info(self.handle, L.dptr(info)), "grid_info")
        return int(info[0]), int(info[1]), info[2], info[3], info[4]

    def dump_cells(self) -> np.ndarray:
        sx, sy, *_ = self.grid_info()
        out = np.zeros((sx * sy, 16))
        L.check(L.lib.ndt2d_matcher_dump_cells(self.handle, L.dptr(out)), "dump_cells")
        return out

    def dump_keys(self, n_points: int) -> np.ndarray:
        out = np.zeros(n_points, dtype=np.int32)
        L.check(L.lib.ndt2d_matcher_dump_keys(self.handle, out.ctypes.data_as(C.POINTER(C.c_int32)),
                                              n_points), "dump_keys")
        return out

    def dump_scores(self, pose, points) -> np.ndarray:
        na, nl = self.search_shape()
        pose3 = _pose3(pose)
        pts = L.f64(points).reshape(-1, 2)
        out = np.zeros(na * nl * nl)
        L.check(L.lib.ndt2d_matcher_dump_scores(self.handle, L.dptr(pose3), L.dptr(pts), pts.shape[0],
                                                L.dptr(out), out.shape[0]), "dump_scores")
        return out.reshape(na, nl, nl)

    def counters(self) -> dict:
        out = np.zeros(4, dtype=np.uint64)
        L.check(L.lib.ndt2d_matcher_counters(self.handle, L.u64ptr(out)), "counters")
        return dict(launches=int(out[0]), h2d_bytes=int(out[1]), d2h_bytes=int(out[2]),
                    valid_cells=int(out[3]))

    def group_info(self) -> dict:
        """Multi-device handle: devices, fused P2P exchange in use, searches spread over all devices."""
        out = np.zeros(4, dtype=np.uint64)
        L.check(L.lib.ndt2d_matcher_group_info(self.handle, L.u64ptr(out)), "group_info")
        return dict(devices=int(out[0]), p2p=bool(out[1]), group_searches=int(out[2]), seq=int(out[3]))

    def set_group_threshold(self, min_pairs: float) -> None:
        """Smallest search (candidate x point pairs) a multi-device handle spreads over its devices."""
        L.check(L.lib.ndt2d_matcher_set_group_threshold(self.handle, float(min_pairs)), "set_group_threshold")

    def probe_call_latency(self, pose, points, calls: int = 200, map_scans=None) -> np.ndarray:
        """Microseconds per call as a C caller sees them (ndt2d_probe_call_latency): matchScan alone,
        or, with map_scans = (poses, offsets, points), reset + addScans + scoreScan + matchScan."""
        pose3 = _pose3(pose)
        pts = L.f64(points).reshape(-1, 2)
        out = np.zeros(calls)
        if map_scans is None:
            L.check(L.lib.ndt2d_probe_call_latency(self.handle, 0, 0, None, None, None, L.dptr(pose3),
                                                   L.dptr(pts), pts.shape[0], calls, L.dptr(out)),
                    "probe_call_latency")
        else:
            poses = L.f64(map_scans[0]).reshape(-1, 3)
            offs = np.ascontiguousarray(map_scans[1], dtype=np.uint64)
            mpts = L.f64(map_scans[2]).reshape(-1, 2)
            L.check(L.lib.ndt2d_probe_call_latency(self.handle, 1, poses.shape[0], L.dptr(poses),
                                                   L.u64ptr(offs), L.dptr(mpts), L.dptr(pose3),
                                                   L.dptr(pts), pts.shape[0], calls, L.dptr(out)),
                    "probe_call_latency")
        return out

    def set_tallies(self, on: bool) -> None:
        """Useful-evaluation / item tallies of the large-search kernel (ndt2d_matcher_set_tallies; default off)."""
        L.check(L.lib.ndt2d_matcher_set_tallies(self.handle, int(bool(on))), "set_tallies")

    def set_timing(self, on: bool) -> None:
        """CUDA event timing of small searches / builds too (ndt2d_matcher_set_timing; default off)."""
        L.check(L.lib.ndt2d_matcher_set_timing(self.handle, int(bool(on))), "set_timing")

    def group_search_stats(self) -> dict:
        """Per-device search-kernel durations of the last matchScan + tallies over all devices."""
        ms = np.zeros(16)
        tot = np.zeros(3, dtype=np.uint64)
        L.check(L.lib.ndt2d_matcher_group_search_stats(self.handle, L.dptr(ms), 16, L.u64ptr(tot)),
                "group_search_stats")
        n = self.group_info()["devices"]
        return dict(kernel_ms=ms[:n].tolist(), useful_evaluations=int(tot[0]), items=int(tot[1]))

    def search_stats(self) -> dict:
        """Work done by the last search launch (ndt2d_matcher_search_stats)."""
        out = np.zeros(4, dtype=np.uint64)
        L.check(L.lib.ndt2d_matcher_search_stats(self.handle, L.u64ptr(out)), "search_stats")
        return dict(useful_evaluations=int(out[0]), items=int(out[1]), jobs_drawn=int(out[2]),
                    kernel_ms=float(out[3]) * 1e-6)

    def build_stats(self) -> dict:
        """The last model build (ndt2d_matcher_build_stats)."""
        out = np.zeros(4, dtype=np.uint64)
        L.check(L.lib.ndt2d_matcher_build_stats(self.handle, L.u64ptr(out)), "build_stats")
        return dict(kernels_ms=float(out[0]) * 1e-6, points=int(out[1]), valid_cells=int(out[2]),
                    grid_cells=int(out[3]))

    def stream(self) -> int:
        return int(L.lib.ndt2d_matcher_stream(self.handle) or 0)
