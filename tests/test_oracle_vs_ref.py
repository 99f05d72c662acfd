"""Pins the oracle restatement (oracle/ndt2d_oracle.c) against the reference's own
sources compiled in place (oracle/_ref): ScanMatcherNDT, NDT and ParticleFilter on
identical synthetic inputs.  The reference's gtests never call these paths
(SURVEY.md section 4), so this is where they get pinned."""
import ctypes as C

import numpy as np
import pytest

from ndt_2d_b200 import synth
from oracle import binding as B


@pytest.fixture(scope="module")
def both(oracle, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built (reference tree not available here)")
    return oracle, ref


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("beams", [360, 100])
def test_matcher_config1(both, beams):
    o, r = both
    w = synth.config1(laser_max_beams=beams)
    mo, mr = o.new_matcher(w.params), r.new_matcher(w.params)
    for m in (mo, mr):
        m.add_scans(w.map_poses, w.map_offsets, w.map_points)
    assert mo.grid() == mr.grid()
    assert np.array_equal(mo.dump_cells(), mr.dump_cells())          # bit-identical cells
    so, do, wo, co, _ = mo.match_scan(w.query_pose, w.query_points)
    sr, dr, wr, cr, _ = mr.match_scan(w.query_pose, w.query_points)
    assert wo == wr and np.array_equal(do, dr)
    assert so == sr
    np.testing.assert_allclose(co, cr, rtol=1e-12)
    for pose in (w.query_pose, w.true_pose, w.true_pose + [3.0, -2.0, 0.5]):
        assert mo.score_points(w.query_points, pose) == mr.score_points(w.query_points, pose)


def test_matcher_no_map_and_quirks(both):
    o, r = both
    w = synth.config1()
    for lib in both:
        m = lib.new_matcher(w.params)
        s, d, written, cov, _ = m.match_scan(w.query_pose, w.query_points)
        assert s == 0.0 and not written and np.all(np.isnan(cov))     # outputs untouched (:80)
        assert m.score_points(w.query_points, w.query_pose) == 0.0    # :159
    # bounding box max initialised with DBL_MIN: all-negative poses stretch the grid to ~0 (Q4)
    poses = np.array([[-30.0, -40.0, 0.1]])
    offs = np.array([0, 3], dtype=np.uint64)
    pts = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.5]])
    p = dict(w.params, range_max=5.0)
    go = o.new_matcher(p)
    gr = r.new_matcher(p)
    go.add_scans(poses, offs, pts)
    gr.add_scans(poses, offs, pts)
    assert go.grid() == gr.grid()
    assert go.grid()[0] == int((2.2250738585072014e-308 - (-35.0)) / 0.25 + 1)


def test_matcher_far_scan_leaves_pose_untouched(both):
    w = synth.config1()
    for lib in both:
        m = lib.new_matcher(w.params)
        m.add_scans(w.map_poses, w.map_offsets, w.map_points)
        # a scan that lands outside every occupied cell: every candidate scores 0
        s, d, written, cov, _ = m.match_scan([500.0, 500.0, 0.0], w.query_points)
        assert not written and s == 0.0
        assert np.all(np.isnan(cov))          # s == 0 -> (1/s) k is NaN (:146)


@pytest.mark.parametrize("seed", range(10))
def test_random_small_worlds_oracle_vs_reference(both, seed):
    """Randomised worlds, poses (negative coordinates included), scan counts and matcher
    parameters: the restatement must reproduce the compiled reference bit for bit -- grid,
    every cell, best pose, score, covariance, scorePoints."""
    o, r = both
    rng = np.random.default_rng(4000 + seed)
    arena = float(rng.choice([12.0, 25.0, 60.0]))
    rects = synth.world(seed=300 + seed, arena=arena, n_obstacles=int(rng.integers(3, 25)),
                        side_min=0.3, side_max=3.0)
    n_scans = int(rng.integers(1, 9))
    beams = int(rng.choice([45, 180, 360]))
    rmax = float(rng.choice([3.5, 8.0, 20.0]))
    centre = rng.uniform(0.2 * arena, 0.8 * arena, size=2)
    poses = np.column_stack([centre[0] + rng.normal(0, 0.4, n_scans + 1), centre[1] + rng.normal(0, 0.4, n_scans + 1),
                             rng.uniform(-np.pi, np.pi, n_scans + 1)])
    offs, pts = synth.scans(rects, poses, beams, rmax, seed=500 + seed,
                            noise_sigma=float(rng.choice([0.0, 0.01, 0.05])), arena=arena)
    poses[:, :2] += rng.choice([0.0, -arena, -3.0 * arena])      # all-negative / mixed-sign coordinates
    lres = float(rng.choice([0.02, 0.05, 0.1]))
    p = dict(ndt_resolution=float(rng.choice([0.1, 0.25, 0.5, 1.0])),
             search_angular_resolution=float(rng.choice([0.005, 0.02])),
             search_angular_size=float(rng.choice([0.01, 0.05])),
             search_linear_resolution=lres, search_linear_size=lres * float(rng.choice([1.5, 4.0])),
             laser_max_beams=int(rng.choice([30, 100, 1000])), range_max=rmax)
    mo, mr = o.new_matcher(p), r.new_matcher(p)
    map_offs, map_pts = offs[: n_scans + 1], pts[: int(offs[n_scans])]
    for m in (mo, mr):
        m.add_scans(poses[:n_scans], map_offs, map_pts)
    assert mo.grid() == mr.grid()
    assert np.array_equal(mo.dump_cells(), mr.dump_cells(), equal_nan=True)
    q = pts[int(offs[n_scans]):int(offs[n_scans + 1])]
    guess = poses[n_scans] + np.array([0.5 * lres, -1.2 * lres, 0.004])
    so, do, wo, co, _ = mo.match_scan(guess, q)
    sr, dr, wr, cr, _ = mr.match_scan(guess, q)
    assert wo == wr and (not wo or np.array_equal(do, dr))
    assert so == sr or (np.isnan(so) and np.isnan(sr))
    np.testing.assert_allclose(co, cr, rtol=1e-12, equal_nan=True)
    if q.shape[0]:
        assert mo.score_points(q, guess) == mr.score_points(q, guess)


def test_particle_filter_measure_and_stats(both):
    o, r = both
    w = synth.config2(n_side=10, n_particles=300)
    mo, mr = o.new_matcher(w.params), r.new_matcher(w.params)
    for m in (mo, mr):
        m.add_scans(w.map_poses, w.map_offsets, w.map_points)
    P = w.particles.shape[0]
    # reference: ParticleFilter::measure through the ScanMatcher interface
    alphas = np.full(5, 0.2)
    f = r.pf_create(50, P, _d(alphas))
    w0 = np.full(P, 1.0 / P)
    r.pf_set(f, _d(w.particles), _d(w0), P)
    cov0 = np.zeros((3, 3))
    r.pf_set_cov(f, _d(cov0))
    r.pf_measure(f, mr.h, _d(w.scan_points), w.scan_points.shape[0])
    pr, wr = np.zeros((P, 3)), np.zeros(P)
    r.pf_get(f, _d(pr), _d(wr))
    mean_r, cov_r = np.zeros(3), np.zeros((3, 3))
    r.pf_stats(f, _d(mean_r), _d(cov_r))
    # oracle
    raw = B.pf_measure(o, mo, w.particles, w.scan_points)
    wo, mean_o, cov_o = B.pf_update_statistics(o, w.particles, raw, cov0)
    assert np.array_equal(wo, wr)
    np.testing.assert_allclose(mean_o, mean_r, rtol=0, atol=1e-15)
    np.testing.assert_allclose(cov_o, cov_r, rtol=1e-13, atol=1e-18)
    # cov(2,2) accumulates across calls (particle_filter.cpp:216)
    r.pf_update_statistics(f)
    cov_r2 = np.zeros((3, 3))
    r.pf_stats(f, _d(mean_r), _d(cov_r2))
    _, _, cov_o2 = B.pf_update_statistics(o, w.particles, wo, cov_o)
    assert cov_r2[2, 2] > cov_r[2, 2]
    np.testing.assert_allclose(cov_o2, cov_r2, rtol=1e-13, atol=1e-18)
    r.pf_destroy(f)


@pytest.mark.parametrize("seed,kld_err,kld_z", [(123, 0.01, 2.3), (7, 0.05, 1.0), (99, 0.99, 0.01)])
def test_particle_filter_resample(both, seed, kld_err, kld_z):
    """resample with the reference's generator re-seeded and its uniform stream replayed."""
    o, r = both
    P, min_p, max_p = 400, 50, 400
    particles = np.stack([1.25 + 0.6 * synth.normal(seed, P), 0.5 + 0.6 * synth.normal(seed + 1, P),
                          1.57 + 0.3 * synth.normal(seed + 2, P)], 1)
    weights = synth.uniform(seed + 3, P) + 0.01
    weights /= weights.sum()
    alphas = np.full(5, 0.2)
    f = r.pf_create(min_p, max_p, _d(alphas))
    r.pf_set(f, _d(particles), _d(weights), P)
    r.pf_seed(f, seed, seed)
    cov0 = np.zeros((3, 3))
    r.pf_set_cov(f, _d(cov0))
    r.pf_resample(f, kld_err, kld_z)
    n = r.pf_size(f)
    pr, wr = np.zeros((n, 3)), np.zeros(n)
    r.pf_get(f, _d(pr), _d(wr))
    u = np.zeros(max_p)
    r.canonical_uniforms(seed, max_p, _d(u))
    po, wo, idx = B.pf_resample(o, particles, weights, min_p, max_p, kld_err, kld_z, u)
    assert po.shape[0] == n
    assert np.array_equal(po, pr)
    wn, mean_o, cov_o = B.pf_update_statistics(o, po, wo, cov0)
    assert np.array_equal(wn, wr)
    r.pf_destroy(f)
