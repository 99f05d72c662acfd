#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE ITSELF.

    python tests/golden/make_golden.py          (needs /root/reference -> oracle/_ref)

Every output in these files was produced by the reference's own sources
(src/ndt_model.cpp, src/scan_matcher_ndt.cpp, src/particle_filter.cpp, src/occupancy_grid.cpp, compiled
unmodified and in place by oracle/Makefile into oracle/_ref/libndt2d_ref.so, g++ -O3
-DNDEBUG, no -march -- the reference's Release flags).  The inputs are stored next to
the outputs so the fixtures do not depend on the synthetic generator staying unchanged.
tests/test_golden.py checks the oracle restatement (CPU) and the CUDA path (GPU) against
them.  /root/reference is only needed to RE-generate; the committed files travel.
"""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from ndt_2d_b200 import synth  # noqa: E402  (host-only synthetic world generator)
from oracle import binding as B  # noqa: E402


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def occupied(cells):
    idx = np.nonzero(cells[:, 1] > 0)[0]
    return idx.astype(np.int64), cells[idx]


def matcher_case(r, name, params, map_poses, map_offsets, map_points, queries, score_poses):
    """queries: list of (pose3, points); score_poses: [K,3] poses scored with queries[0]'s points."""
    m = r.new_matcher(params)
    m.add_scans(map_poses, map_offsets, map_points)
    cells = m.dump_cells()
    idx, occ = occupied(cells)
    out = dict(
        params=np.array([params[k] for k in synth.PARAM_KEYS], dtype=np.float64),
        map_poses=map_poses, map_offsets=map_offsets.astype(np.uint64), map_points=map_points,
        grid=np.array(m.grid(), dtype=np.float64), cell_index=idx, cell_values=occ,
        candidate_count=np.array([r.matcher_candidate_count(m.h)], dtype=np.uint64),
        score_poses=np.asarray(score_poses, dtype=np.float64),
    )
    res = []
    for k, (pose, pts) in enumerate(queries):
        s, d, written, cov, _ = m.match_scan(pose, pts)
        out[f"q{k}_pose"] = np.asarray(pose, dtype=np.float64)
        out[f"q{k}_points"] = np.asarray(pts, dtype=np.float64)
        out[f"q{k}_score"] = np.array([s])
        out[f"q{k}_delta"] = d
        out[f"q{k}_written"] = np.array([int(written)])
        out[f"q{k}_cov"] = cov
        res.append(s)
    out["n_queries"] = np.array([len(queries)])
    pts0 = np.asarray(queries[0][1], dtype=np.float64)
    out["score_values"] = np.array([m.score_points(pts0, p) for p in out["score_poses"]])
    out["likelihood_scan"] = np.array([
        r.ndt_likelihood_scan(r.matcher_ndt(m.h), _d(np.ascontiguousarray(p)), _d(pts0), pts0.shape[0])
        for p in out["score_poses"]])
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(f"{name}: grid {m.grid()[:2]} occupied {idx.size} scores {res}")


def filter_case(r, name):
    w = synth.config2(n_side=8, n_particles=400)
    m = r.new_matcher(w.params)
    m.add_scans(w.map_poses, w.map_offsets, w.map_points)
    P = w.particles.shape[0]
    alphas = np.full(5, 0.2)
    f = r.pf_create(50, P, _d(alphas))
    w0 = np.full(P, 1.0 / P)
    cov0 = np.zeros((3, 3))
    r.pf_set(f, _d(w.particles), _d(w0), P)
    r.pf_set_cov(f, _d(cov0))
    r.pf_measure(f, m.h, _d(w.scan_points), w.scan_points.shape[0])
    pw = np.zeros(P)
    pp = np.zeros((P, 3))
    r.pf_get(f, _d(pp), _d(pw))
    mean, cov = np.zeros(3), np.zeros((3, 3))
    r.pf_stats(f, _d(mean), _d(cov))
    out = dict(
        params=np.array([w.params[k] for k in synth.PARAM_KEYS], dtype=np.float64),
        map_poses=w.map_poses, map_offsets=w.map_offsets.astype(np.uint64), map_points=w.map_points,
        particles=w.particles, scan_points=w.scan_points, measured_weights=pw, mean=mean, cov=cov)
    # resample with the reference's generator re-seeded; the uniform stream its
    # discrete_distribution consumes is stored so that other implementations can replay it
    for k, (seed, kld_err, kld_z, min_p, max_p) in enumerate(
            [(123, 0.01, 2.3, 50, 400), (7, 0.05, 1.0, 50, 400), (99, 0.99, 0.01, 50, 400)]):
        g = r.pf_create(min_p, max_p, _d(alphas))
        r.pf_set(g, _d(w.particles), _d(pw), P)
        r.pf_seed(g, seed, seed)
        r.pf_set_cov(g, _d(cov0))
        r.pf_resample(g, kld_err, kld_z)
        n = r.pf_size(g)
        rp, rw = np.zeros((n, 3)), np.zeros(n)
        r.pf_get(g, _d(rp), _d(rw))
        rmean, rcov = np.zeros(3), np.zeros((3, 3))
        r.pf_stats(g, _d(rmean), _d(rcov))
        u = np.zeros(max_p)
        r.canonical_uniforms(seed, max_p, _d(u))
        out[f"r{k}_args"] = np.array([kld_err, kld_z, min_p, max_p], dtype=np.float64)
        out[f"r{k}_uniforms"] = u
        out[f"r{k}_particles"] = rp
        out[f"r{k}_weights"] = rw
        out[f"r{k}_mean"] = rmean
        out[f"r{k}_cov"] = rcov
        r.pf_destroy(g)
        print(f"{name}: resample {k} -> {n} particles")
    out["n_resamples"] = np.array([3])
    r.pf_destroy(f)
    np.savez_compressed(HERE / f"{name}.npz", **out)


def kd_case(r, name):
    poses = np.stack([3.0 * synth.normal(31, 500), 3.0 * synth.normal(32, 500),
                      2.0 * synth.normal(33, 500)], 1)
    poses[:5] = [[0, 0, 0], [0, 0, 0], [0.75, 0, 0], [-0.75, 0, 0], [0.75, 0.75, 0]]  # particle_tests.cpp:47-72
    counts = r.kd_counts(poses)
    np.savez_compressed(HERE / f"{name}.npz", poses=poses, counts=counts)
    print(f"{name}: leaf counts {counts[:5].tolist()} ... {int(counts[-1])}")


def occupancy_case(r, name):
    """ndt_2d::OccupancyGrid::getMsg twice on one instance: 6 scans, then 10 (the bounds persist)."""
    w = synth.config1()
    g = B.OccupancyGrid(r, 0.05, 0.25)
    out = dict(poses=w.map_poses, offsets=w.map_offsets.astype(np.uint64), points=w.map_points)
    for k, n in enumerate((6, 10)):
        info, data = g.get_msg(w.map_poses[:n], w.map_offsets[:n + 1], w.map_points)
        out[f"c{k}_n"] = np.array([n])
        out[f"c{k}_info"] = np.array([info["width"], info["height"], info["origin_x"], info["origin_y"],
                                      info["resolution"]])
        out[f"c{k}_data"] = data
        print(f"{name}: {n} scans -> {info['width']} x {info['height']}, "
              f"{int((data == 100).sum())} occupied, {int((data == 0).sum())} free")
    np.savez_compressed(HERE / f"{name}.npz", **out)


def main():
    B.build(quiet=True)
    r = B.load_ref()
    if r is None:
        raise SystemExit("oracle/_ref/libndt2d_ref.so missing: /root/reference is needed to regenerate")
    # 1. config 1 (BASELINE.json configs[0]): 360-beam scan vs rolling NDT of 10 scans
    for beams in (360, 100):
        w = synth.config1(laser_max_beams=beams)
        poses = [w.query_pose, w.true_pose, w.true_pose + np.array([3.0, -2.0, 0.5]),
                 w.true_pose + np.array([0.01, 0.02, -0.01]), np.array([500.0, 500.0, 0.0])]
        matcher_case(r, f"config1_beams{beams}", w.params, w.map_poses, w.map_offsets, w.map_points,
                     [(w.query_pose, w.query_points), (np.array([500.0, 500.0, 0.0]), w.query_points),
                      (w.true_pose, w.query_points[:37])], poses)
    # 2. the plugin's default parameters (scan_matcher_ndt.cpp:37-44): 21 x 21 x 80 candidates
    w = synth.config1()
    defaults = dict(ndt_resolution=0.25, search_angular_resolution=0.0025, search_angular_size=0.1,
                    search_linear_resolution=0.005, search_linear_size=0.05, laser_max_beams=100,
                    range_max=10.0)
    guess = w.true_pose - np.array([0.02, -0.03, 0.04])
    matcher_case(r, "plugin_defaults", defaults, w.map_poses, w.map_offsets, w.map_points,
                 [(guess, w.query_points)], [guess, w.true_pose])
    # 3. large-search shapes on a small window (config 4 at scale 0.04)
    w = synth.config4(scale=0.04)
    guess = w.true_pose - np.array([0.05, -0.03, 0.06])
    matcher_case(r, "config4_window", w.params, w.map_poses, w.map_offsets, w.map_points,
                 [(guess, w.query_points)], [guess, w.true_pose])
    # 4. all-negative poses: the DBL_MIN bounding-box quirk, points outside the grid
    p = dict(ndt_resolution=0.25, search_angular_resolution=0.01, search_angular_size=0.02,
             search_linear_resolution=0.05, search_linear_size=0.1, laser_max_beams=100, range_max=5.0)
    poses = np.array([[-30.0, -40.0, 0.3], [-31.0, -40.5, -0.2]])
    pts = np.concatenate([np.array([[1.0, 0.2], [1.01, 0.21], [0.99, 0.19], [1.0, 0.22], [1.02, 0.2],
                                    [9.0, 9.0], [-4.9, 0.0]]),
                          np.array([[2.0, 0.7], [2.01, 0.71], [1.99, 0.69]])])
    offs = np.array([0, 7, 10], dtype=np.uint64)
    q = np.array([[1.0, 0.2], [1.01, 0.2], [2.0, 0.7]])
    matcher_case(r, "bbox_quirk", p, poses, offs, pts, [(poses[0], q)], [poses[0], poses[1]])
    filter_case(r, "particle_filter")
    kd_case(r, "kd_tree")
    occupancy_case(r, "occupancy_grid")


if __name__ == "__main__":
    main()
