"""Parity of the CUDA path (through the C ABI) with the oracle, on identical
synthetic inputs.  Contract (BASELINE.json north_star):
  * cell indices and candidate counts: bit-exact
  * per-cell mean / covariance and scores: within 1e-5 relative
  * best pose identical unless reference scores tie within that tolerance
Scores additionally get an absolute floor of 1e-30: a per-point likelihood below
2^-126 flushes to zero in the FP32 exp2 the device uses for the final 2^t.
"""
import math

import numpy as np
import pytest

from ndt_2d_b200 import ParticleFilter, Pose2d, Scan, ScanMatcherNDT, synth
from ndt_2d_b200 import _lib as L
from oracle import binding as B

pytestmark = pytest.mark.gpu

RTOL = 1e-5       # the north_star tolerance
ATOL_SCORE = 1e-30


def world_points(poses, offsets, points):
    """NDT::addScan's transform (ndt_model.cpp:135-143) in numpy: separate IEEE
    multiplies/adds, no FMA, cos/sin from the host libm."""
    out = np.empty_like(points)
    for k in range(poses.shape[0]):
        lo, hi = int(offsets[k]), int(offsets[k + 1])
        c, s = math.cos(poses[k, 2]), math.sin(poses[k, 2])
        px, py = points[lo:hi, 0], points[lo:hi, 1]
        out[lo:hi, 0] = poses[k, 0] + (px * c - py * s)
        out[lo:hi, 1] = poses[k, 1] + (px * s + py * c)
    return out


def ref_keys(grid, wp):
    """NDT::getIndex (ndt_model.cpp:203-218), vectorised."""
    sx, sy, ox, oy, cs = grid
    x, y = wp[:, 0], wp[:, 1]
    inside = (x >= ox) & (y >= oy)
    gx = np.zeros(x.shape, dtype=np.int64)
    gy = np.zeros(x.shape, dtype=np.int64)
    gx[inside] = np.trunc((x[inside] - ox) / cs).astype(np.int64)
    gy[inside] = np.trunc((y[inside] - oy) / cs).astype(np.int64)
    inside &= (gx < sx) & (gy < sy)
    return np.where(inside, gy * sx + gx, -1).astype(np.int32)


def check_cells(gpu_cells, orc_cells):
    """valid, n, mean, covariance, correlation exact; information to 1e-12 (same IEEE
    operations on both sides), far inside the 1e-5 contract."""
    assert np.array_equal(gpu_cells[:, 1], orc_cells[:, 1])                       # n
    assert np.array_equal(gpu_cells[:, 0], orc_cells[:, 0])                       # valid
    assert np.array_equal(gpu_cells[:, 2:4], orc_cells[:, 2:4])                   # mean
    assert np.array_equal(gpu_cells[:, 8:12], orc_cells[:, 8:12])                 # correlation
    assert np.array_equal(gpu_cells[:, 4:8], orc_cells[:, 4:8])                   # covariance
    np.testing.assert_allclose(gpu_cells[:, 12:16], orc_cells[:, 12:16], rtol=1e-12, atol=0)


def check_match(gpu, orc, scores_gpu=None, scores_orc=None):
    sg, dg, wg, cg = gpu
    so, do, wo, co = orc
    assert wg == wo
    np.testing.assert_allclose(sg, so, rtol=RTOL, atol=ATOL_SCORE)
    np.testing.assert_allclose(cg, co, rtol=RTOL, atol=RTOL * np.abs(co).max())
    if wo and not np.array_equal(dg, do):
        # allowed only when the two candidates tie within tolerance in the reference
        assert scores_orc is not None, "best pose differs and no score volume to justify a tie"
        flat = scores_orc.ravel()
        best = flat.min()
        tied = np.abs(flat - best) <= RTOL * abs(best)
        assert tied.sum() > 1, f"best pose differs without a tie: gpu {dg} oracle {do}"
    if scores_gpu is not None:
        assert scores_gpu.shape == scores_orc.shape                               # candidate count
        np.testing.assert_allclose(scores_gpu, scores_orc, rtol=RTOL, atol=ATOL_SCORE)


@pytest.fixture(scope="module")
def o(oracle, gpu):
    return oracle


# ------------------------------------------------------------------ build
def test_divide_by_point_count_is_the_ieee_divide(o):
    """The single-CTA build replaces Cell::addPoint's divide by n + 1 (ndt_model.cpp:55-61) with a
    quotient formed from the count's reciprocal (csrc/build_common.cuh); it has to be the correctly
    rounded quotient bit for bit -- 2.6e8 pseudo-random (numerator, count) pairs per seed."""
    for seed in (1, 0xDEADBEEF, 2 ** 61 + 7):
        bad = np.zeros(1, dtype=np.uint64)
        L.check(L.lib.ndt2d_probe_div_by_count(0, seed, 1 << 28, L.u64ptr(bad)), "probe_div_by_count")
        assert int(bad[0]) == 0


def test_build_long_cells_small_model(o):
    """A rolling window whose points pile up in a few cells (hundreds of points per cell): the
    longest dependency chains of the single-CTA build, cells still bit-identical."""
    rng = np.random.default_rng(11)
    n_scans, per = 8, 500
    poses = np.zeros((n_scans, 3))
    poses[:, 0] = 0.01 * np.arange(n_scans)
    poses[:, 2] = 0.002 * np.arange(n_scans)
    centres = np.array([[2.1, 0.3], [2.4, 0.35], [-1.2, 3.3], [0.4, -2.7]])
    pts = (centres[rng.integers(0, 4, n_scans * per)] + rng.normal(0, 0.05, (n_scans * per, 2)))
    offsets = (per * np.arange(n_scans + 1)).astype(np.uint64)
    prm = synth.config1().params
    m = ScanMatcherNDT.from_params(prm)
    mo = o.new_matcher(prm)
    m.add_scans_raw(poses, offsets, pts)
    mo.add_scans(poses, offsets, pts)
    assert m.grid_info() == mo.grid()
    cells = m.dump_cells()
    assert cells[:, 1].max() >= 300
    check_cells(cells, mo.dump_cells())
    m.close()


@pytest.mark.parametrize("cfg", ["config1", "config4"])
def test_build_parity(o, cfg):
    w = getattr(synth, cfg)()
    m = ScanMatcherNDT.from_params(w.params)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    mo = o.new_matcher(w.params)
    mo.add_scans(w.map_poses, w.map_offsets, w.map_points)
    assert m.grid_info() == mo.grid()                                             # grid, bit-exact
    keys = m.dump_keys(w.map_points.shape[0])
    wp = world_points(w.map_poses, w.map_offsets, w.map_points)
    expect = ref_keys(mo.grid(), wp)
    assert np.array_equal(keys, expect)                                           # cell indices
    # spot-check the vectorised restatement against the oracle's own getIndex
    for i in range(0, wp.shape[0], 97):
        assert mo.get_index(wp[i, 0], wp[i, 1]) == expect[i]
    check_cells(m.dump_cells(), mo.dump_cells())
    assert m.counters()["valid_cells"] == int((mo.dump_cells()[:, 1] >= 5).sum())


def test_build_bounding_box_quirk(o):
    """max_* start at DBL_MIN (scan_matcher_ndt.cpp:54,56): all-negative poses stretch
    the grid up to ~0; points outside the grid are dropped, not clamped."""
    p = dict(ndt_resolution=0.25, search_angular_resolution=0.01, search_angular_size=0.02,
             search_linear_resolution=0.05, search_linear_size=0.1, laser_max_beams=100,
             range_max=5.0)
    poses = np.array([[-30.0, -40.0, 0.3], [-31.0, -40.5, -0.2]])
    pts = np.concatenate([np.array([[1.0, 0.2], [1.01, 0.21], [0.99, 0.19], [1.0, 0.22], [1.02, 0.2],
                                    [9.0, 9.0], [-4.9, 0.0]]),
                          np.array([[2.0, 0.7], [2.01, 0.71], [1.99, 0.69]])])
    offs = np.array([0, 7, 10], dtype=np.uint64)
    m = ScanMatcherNDT.from_params(p)
    m.add_scans_raw(poses, offs, pts)
    mo = o.new_matcher(p)
    mo.add_scans(poses, offs, pts)
    assert m.grid_info() == mo.grid()
    assert np.array_equal(m.dump_keys(10), ref_keys(mo.grid(), world_points(poses, offs, pts)))
    check_cells(m.dump_cells(), mo.dump_cells())


def test_build_empty_and_ragged(o):
    w = synth.config1()
    m = ScanMatcherNDT.from_params(w.params)
    mo = o.new_matcher(w.params)
    # scans with zero points in the middle, and a model with no points at all
    poses = w.map_poses[:3]
    offs = np.array([0, int(w.map_offsets[1]), int(w.map_offsets[1]), int(w.map_offsets[2])], dtype=np.uint64)
    m.add_scans_raw(poses, offs, w.map_points)
    mo.add_scans(poses, offs, w.map_points)
    assert m.grid_info() == mo.grid()
    check_cells(m.dump_cells(), mo.dump_cells())
    offs0 = np.zeros(4, dtype=np.uint64)
    m.add_scans_raw(poses, offs0, np.zeros((0, 2)))
    mo.add_scans(poses, offs0, np.zeros((0, 2)))
    assert m.grid_info() == mo.grid()
    assert not m.dump_cells().any()
    s, d, written, cov, _ = m.match_scan_raw(w.query_pose, w.query_points)
    so, do, wo, co, _ = mo.match_scan(w.query_pose, w.query_points)
    assert (s, written) == (so, wo) and np.all(np.isnan(cov)) and np.all(np.isnan(co))


def test_build_degenerate_cell_is_nan_like_reference(o):
    """Q7: >= 5 identical points -> zero covariance -> inf/NaN information; the score
    volume must be NaN-poisoned in the same candidates."""
    p = dict(ndt_resolution=0.5, search_angular_resolution=0.01, search_angular_size=0.02,
             search_linear_resolution=0.05, search_linear_size=0.1, laser_max_beams=100,
             range_max=4.0)
    poses = np.array([[0.0, 0.0, 0.0]])
    pts = np.array([[1.1, 1.1]] * 6 + [[2.3, 0.4], [2.35, 0.45], [2.32, 0.42], [2.31, 0.47], [2.36, 0.41]])
    offs = np.array([0, pts.shape[0]], dtype=np.uint64)
    m = ScanMatcherNDT.from_params(p)
    m.add_scans_raw(poses, offs, pts)
    mo = o.new_matcher(p)
    mo.add_scans(poses, offs, pts)
    gc, oc = m.dump_cells(), mo.dump_cells()
    assert np.array_equal(np.isnan(gc), np.isnan(oc)) and np.array_equal(np.isinf(gc), np.isinf(oc))
    q = np.array([[1.1, 1.1], [2.3, 0.45]])
    sg = m.dump_scores([0.0, 0.0, 0.0], q)
    _, _, _, _, so = mo.match_scan([0.0, 0.0, 0.0], q, want_scores=True)
    assert np.array_equal(np.isnan(sg), np.isnan(so))
    ok = ~np.isnan(so)
    np.testing.assert_allclose(sg[ok], so[ok], rtol=RTOL, atol=ATOL_SCORE)


# ------------------------------------------------------------------ search
@pytest.mark.parametrize("beams", [360, 100])
@pytest.mark.parametrize("variant", [0, 1, 3, 4, 5])
def test_match_scan_config1(o, beams, variant):
    w = synth.config1(laser_max_beams=beams)
    m = ScanMatcherNDT.from_params(w.params, kernel_variant=variant)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    mo = o.new_matcher(w.params)
    mo.add_scans(w.map_poses, w.map_offsets, w.map_points)
    assert m.search_shape() == (200, 10)                                          # candidate counts
    so, do, wo, co, scores_o = mo.match_scan(w.query_pose, w.query_points, want_scores=True)
    sg, dg, wg, cg, st = m.match_scan_raw(w.query_pose, w.query_points)
    scores_g = m.dump_scores(w.query_pose, w.query_points)
    check_match((sg, dg, wg, cg), (so, do, wo, co), scores_g, scores_o)
    # the reference-style call: pose is the in/out correction
    score, pose, cov = m.matchScan(Scan(0, Pose2d(*w.query_pose), w.query_points))
    assert score == sg and (pose.x, pose.y, pose.theta) == tuple(dg)


def test_match_scan_plugin_defaults(o):
    """Plugin defaults (scan_matcher_ndt.cpp:37-44): 21 x 21 x 80 candidates."""
    w = synth.config1()
    p = dict(range_max=10.0)
    m = ScanMatcherNDT.from_params(p)
    assert m.search_shape() == (80, 21)
    full = dict(ndt_resolution=0.25, search_angular_resolution=0.0025, search_angular_size=0.1,
                search_linear_resolution=0.005, search_linear_size=0.05, laser_max_beams=100,
                range_max=10.0)
    mo = o.new_matcher(full)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    mo.add_scans(w.map_poses, w.map_offsets, w.map_points)
    guess = w.true_pose - np.array([0.02, -0.03, 0.04])
    so, do, wo, co, scores_o = mo.match_scan(guess, w.query_points, want_scores=True)
    sg, dg, wg, cg, _ = m.match_scan_raw(guess, w.query_points)
    check_match((sg, dg, wg, cg), (so, do, wo, co), m.dump_scores(guess, w.query_points), scores_o)
    dth, dlin = m.search_values()
    assert dlin[-1] == 0.04999999999999999 and len(dlin) == 21                   # accumulated bound


def test_match_scan_edge_cases(o):
    w = synth.config1()
    m = ScanMatcherNDT.from_params(w.params)
    mo = o.new_matcher(w.params)
    # no map: 0.0 and outputs untouched (scan_matcher_ndt.cpp:80)
    pose, cov = Pose2d(1.0, 2.0, 3.0), np.full((3, 3), 7.0)
    score, pose, cov = m.matchScan(Scan(0, Pose2d(*w.query_pose), w.query_points), pose, cov)
    assert score == 0.0 and (pose.x, pose.y, pose.theta) == (1.0, 2.0, 3.0) and np.all(cov == 7.0)
    assert m.scorePoints(w.query_points, w.query_pose) == 0.0
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    mo.add_scans(w.map_poses, w.map_offsets, w.map_points)
    # scan far outside the map: every candidate scores 0 -> pose untouched, NaN covariance
    sg, dg, wg, cg, _ = m.match_scan_raw([500.0, 500.0, 0.0], w.query_points)
    so, do, wo, co, _ = mo.match_scan([500.0, 500.0, 0.0], w.query_points)
    assert (sg, wg) == (so, wo) == (0.0, False) and np.all(np.isnan(cg)) and np.all(np.isnan(co))
    # empty scan: n == 0 -> 0/0
    sg, dg, wg, cg, _ = m.match_scan_raw(w.query_pose, np.zeros((0, 2)))
    so, do, wo, co, _ = mo.match_scan(w.query_pose, np.zeros((0, 2)))
    assert math.isnan(sg) and math.isnan(so) and not wg and not wo
    # reset drops the model
    m.reset()
    assert m.match_scan_raw(w.query_pose, w.query_points)[4] == L.ERR_NO_MAP


def test_match_scan_config4_reduced(o):
    """The large-search shapes (1080 beams, 0.01 m / 0.002 rad steps) on a window the
    oracle finishes in seconds."""
    w = synth.config4(scale=0.04)
    m = ScanMatcherNDT.from_params(w.params)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    mo = o.new_matcher(w.params)
    mo.add_scans(w.map_poses, w.map_offsets, w.map_points)
    guess = w.true_pose - np.array([0.05, -0.03, 0.06])
    so, do, wo, co, scores_o = mo.match_scan(guess, w.query_points, want_scores=True)
    sg, dg, wg, cg, _ = m.match_scan_raw(guess, w.query_points)
    check_match((sg, dg, wg, cg), (so, do, wo, co), m.dump_scores(guess, w.query_points), scores_o)


def test_region_kernel_tables_in_global_memory(o):
    """A map whose dilated bitmap + threshold tables (57 KB at 0.1 m cells over +-30 m) do not fit
    next to the region kernel's per-warp shared memory: the instantiation that reads them
    through L1, with and without the coordinate pre-pass (3 regions per axis are needed for it)."""
    w = synth.config4(scale=0.03)
    for lin_size in (0.06, 0.5):
        p = dict(w.params, ndt_resolution=0.1, search_linear_size=lin_size,
                 search_angular_size=0.02 if lin_size > 0.1 else w.params["search_angular_size"])
        m = ScanMatcherNDT.from_params(p, kernel_variant=4)
        mo = o.new_matcher(p)
        m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
        mo.add_scans(w.map_poses, w.map_offsets, w.map_points)
        sx, sy = m.grid_info()[:2]
        assert (sx + 2) * (sy + 2) // 8 + 8 * (sx + sy + 4) > 24 * 1024
        guess = w.true_pose - np.array([0.03, -0.02, 0.01])
        so, do, wo, co, scores_o = mo.match_scan(guess, w.query_points, want_scores=True)
        sg, dg, wg, cg, _ = m.match_scan_raw(guess, w.query_points)
        check_match((sg, dg, wg, cg), (so, do, wo, co), m.dump_scores(guess, w.query_points), scores_o)
        m.close()


def test_instrumentation_switches(o):
    """Work tallies and small-search event timing are off by default (they cost run time on the
    hot path) and available on request; results do not depend on them; the C-ABI latency probe
    times real calls."""
    w = synth.config4(scale=0.04)
    m = ScanMatcherNDT.from_params(w.params, kernel_variant=4)       # the region kernel, forced
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    guess = w.true_pose - np.array([0.05, -0.03, 0.06])
    a = m.match_scan_raw(guess, w.query_points)
    st = m.search_stats()
    assert st["useful_evaluations"] == 0 and st["items"] == 0
    m.set_tallies(True)
    b = m.match_scan_raw(guess, w.query_points)
    st = m.search_stats()
    assert st["useful_evaluations"] > 0 and 0 < st["items"] < st["useful_evaluations"]
    m.set_tallies(False)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])
    m.close()
    w = synth.config1(laser_max_beams=100)
    m = ScanMatcherNDT.from_params(w.params)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    r0 = m.match_scan_raw(w.query_pose, w.query_points)
    assert m.search_stats()["kernel_ms"] == 0.0                      # small search: no event records
    m.set_timing(True)
    r1 = m.match_scan_raw(w.query_pose, w.query_points)
    assert m.search_stats()["kernel_ms"] > 0.0
    assert r0[0] == r1[0] and np.array_equal(r0[1], r1[1])
    us = m.probe_call_latency(w.query_pose, w.query_points, 20)
    assert us.shape == (20,) and np.all(us > 1.0) and np.all(us < 1.0e5)
    us = m.probe_call_latency(w.query_pose, w.query_points, 5, (w.map_poses, w.map_offsets, w.map_points))
    assert np.all(us > 1.0)
    # the sequence left the handle with the same model: the match is still the same
    r2 = m.match_scan_raw(w.query_pose, w.query_points)
    assert r0[0] == r2[0] and np.array_equal(r0[1], r2[1])
    m.close()


def test_match_scan_dense_clutter_reduced(o):
    """The cluttered short-range world of bench.py's floor workload (a third of the (candidate,
    point) pairs in occupied 0.5 m cells), on a window the oracle finishes in seconds: full score
    volume, pose and covariance."""
    w = synth.config4_dense(scale=0.04)
    m = ScanMatcherNDT.from_params(w.params)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    mo = o.new_matcher(w.params)
    mo.add_scans(w.map_poses, w.map_offsets, w.map_points)
    assert m.grid_info() == mo.grid()
    check_cells(m.dump_cells(), mo.dump_cells())
    for guess in (w.true_pose - np.array([0.05, -0.03, 0.06]), w.true_pose + np.array([0.9, 0.4, 2.0])):
        so, do, wo, co, scores_o = mo.match_scan(guess, w.query_points, want_scores=True)
        sg, dg, wg, cg, _ = m.match_scan_raw(guess, w.query_points)
        check_match((sg, dg, wg, cg), (so, do, wo, co), m.dump_scores(guess, w.query_points), scores_o)
    m.close()


SWEEP = [
    # ndt_res, ang_res, ang_size, lin_res, lin_size, beams   -- what it exercises
    (0.05, 0.01, 0.03, 0.05, 0.25, 360),     # step == cell: one-candidate regions
    (0.05, 0.01, 0.03, 0.11, 0.30, 360),     # step > cell
    (0.30, 0.004, 0.02, 0.07, 0.33, 200),    # non-integer cell/step ratio, n_lin not a region multiple
    (0.25, 0.0025, 0.1, 0.005, 0.05, 100),   # 50 steps per cell: region capped at 25, 21 x 21 lattice
    (0.25, 0.5, 0.2, 0.01, 0.13, 360),       # a single theta slice (accumulated loop: 1 value)
    (0.25, 0.01, 0.05, 0.2, 0.1, 360),       # a single (dx, dy) candidate per slice
    (1.00, 0.01, 0.05, 0.03, 0.30, 37),      # coarse cells, few beams (37 of 360, subsampled)
    (0.10, 0.002, 0.01, 0.004, 0.06, 1000),  # laser_max_beams > points in the scan
    (0.25, 0.003, 0.0, 0.01, 0.1, 360),      # angular size 0: no candidates at all
]


@pytest.mark.parametrize("case", range(len(SWEEP)))
def test_search_parameter_sweep(o, case):
    res, ares, asize, lres, lsize, beams = SWEEP[case]
    w = synth.config1()
    p = dict(ndt_resolution=res, search_angular_resolution=ares, search_angular_size=asize,
             search_linear_resolution=lres, search_linear_size=lsize, laser_max_beams=beams,
             range_max=10.0)
    m = ScanMatcherNDT.from_params(p)
    mo = o.new_matcher(p)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    mo.add_scans(w.map_poses, w.map_offsets, w.map_points)
    assert m.grid_info() == mo.grid()
    check_cells(m.dump_cells(), mo.dump_cells())
    guess = w.true_pose - np.array([0.03, -0.02, 0.01])
    so, do, wo, co, scores_o = mo.match_scan(guess, w.query_points, want_scores=True)
    sg, dg, wg, cg, _ = m.match_scan_raw(guess, w.query_points)
    na, nl = m.search_shape()
    assert scores_o.shape == (na, nl, nl)                                         # candidate counts
    if scores_o.size == 0:
        assert (sg, wg) == (so, wo) and np.all(np.isnan(cg)) and np.all(np.isnan(co))
        return
    check_match((sg, dg, wg, cg), (so, do, wo, co), m.dump_scores(guess, w.query_points), scores_o)


@pytest.mark.parametrize("seed", range(12))
def test_random_small_worlds(o, seed):
    """Randomised worlds, poses (negative coordinates included), scan counts and matcher
    parameters: build bit-exact, full score volume within 1e-5, same best pose."""
    rng = np.random.default_rng(1000 + seed)
    arena = float(rng.choice([12.0, 25.0, 60.0]))
    rects = synth.world(seed=700 + seed, arena=arena, n_obstacles=int(rng.integers(3, 25)),
                        side_min=0.3, side_max=3.0)
    n_scans = int(rng.integers(1, 9))
    beams = int(rng.choice([45, 180, 360, 720]))
    rmax = float(rng.choice([3.5, 8.0, 20.0]))
    centre = rng.uniform(0.2 * arena, 0.8 * arena, size=2)
    poses = np.column_stack([centre[0] + rng.normal(0, 0.4, n_scans + 1), centre[1] + rng.normal(0, 0.4, n_scans + 1),
                             rng.uniform(-np.pi, np.pi, n_scans + 1)])
    offs, pts = synth.scans(rects, poses, beams, rmax, seed=900 + seed, noise_sigma=float(rng.choice([0.0, 0.01, 0.05])),
                            arena=arena)
    shift = rng.choice([0.0, -arena, -3.0 * arena])          # all-negative / mixed-sign coordinates
    poses[:, :2] += shift
    lres = float(rng.choice([0.01, 0.02, 0.05, 0.1]))
    p = dict(ndt_resolution=float(rng.choice([0.1, 0.25, 0.5, 1.0])),
             search_angular_resolution=float(rng.choice([0.002, 0.005, 0.02])),
             search_angular_size=float(rng.choice([0.01, 0.05, 0.1])),
             search_linear_resolution=lres, search_linear_size=lres * float(rng.choice([1.5, 4.0, 9.5])),
             laser_max_beams=int(rng.choice([30, 100, 360, 1000])), range_max=rmax)
    m = ScanMatcherNDT.from_params(p)
    mo = o.new_matcher(p)
    map_offs = offs[: n_scans + 1]
    map_pts = pts[: int(offs[n_scans])]
    m.add_scans_raw(poses[:n_scans], map_offs, map_pts)
    mo.add_scans(poses[:n_scans], map_offs, map_pts)
    assert m.grid_info() == mo.grid()
    check_cells(m.dump_cells(), mo.dump_cells())
    q = pts[int(offs[n_scans]):int(offs[n_scans + 1])]
    guess = poses[n_scans] + np.array([0.5 * lres, -1.2 * lres, 0.004])
    so, do, wo, co, scores_o = mo.match_scan(guess, q, want_scores=True)
    sg, dg, wg, cg, _ = m.match_scan_raw(guess, q)
    if scores_o.size == 0 or q.shape[0] == 0:
        assert wg == wo
        return
    check_match((sg, dg, wg, cg), (so, do, wo, co), m.dump_scores(guess, q), scores_o)
    np.testing.assert_allclose(m.scorePoints(q, guess), mo.score_points(q, guess), rtol=RTOL, atol=ATOL_SCORE)


@pytest.mark.parametrize("variant", [0, 3, 4, 5])
def test_long_scan_small_lattice(o, variant):
    """A 3,240-point scan (more than the dense kernel stages per pass, more than one chunk of
    the region kernel) on a small lattice."""
    w = synth.config1()
    p = dict(w.params, laser_max_beams=5000, search_angular_size=0.02, search_linear_size=0.15)
    pts = np.concatenate([w.query_points] + [w.query_points + 1e-3 * (k + 1) for k in range(8)])
    assert pts.shape[0] > 2048
    m = ScanMatcherNDT.from_params(p, kernel_variant=variant)
    mo = o.new_matcher(p)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    mo.add_scans(w.map_poses, w.map_offsets, w.map_points)
    so, do, wo, co, scores_o = mo.match_scan(w.query_pose, pts, want_scores=True)
    sg, dg, wg, cg, _ = m.match_scan_raw(w.query_pose, pts)
    check_match((sg, dg, wg, cg), (so, do, wo, co), m.dump_scores(w.query_pose, pts), scores_o)


WINDOW_CASES = [
    # lin_res, lin_size, ndt_res, beams, guess offset      -- window kernel shape (variant 5)
    (0.005, 0.05, 0.25, 100, (0.02, -0.03, 0.01)),      # K = 2: plugin defaults, 441 candidates = 4 CTAs per slice
    (0.05, 0.25, 0.25, 360, (0.12, -0.07, 0.02)),       # K = 3: config 1's window
    (0.04, 0.34, 0.25, 360, (0.0, 0.0, 0.0)),           # K = 4: 0.64 m window over 0.25 m cells
    (0.03, 0.2, 0.1, 360, (0.05, 0.05, 0.0)),           # K = 4 again with small cells: 0.39 m / 0.1 m
    (0.02, 0.1, 0.5, 1000, (0.3, -0.2, 0.05)),          # K = 2, every beam
    (0.05, 0.25, 0.25, 360, (9.0, -8.5, 0.3)),          # window over the edge of the grid (most points outside)
    (0.05, 0.25, 0.25, 360, (40.0, 40.0, 0.0)),         # scan entirely outside: every score 0, pose untouched
    (0.1, 0.45, 0.1, 360, (0.0, 0.0, 0.0)),             # too wide (9 cells): falls back to the dense kernel
]


@pytest.mark.parametrize("case", range(len(WINDOW_CASES)))
def test_window_kernel_shapes(o, case):
    """search_window.cu (thread per candidate, per-point tables shared by the CTA) forced with
    kernel_variant 5 over its K = 2 / 3 / 4 instantiations, several CTAs per slice, grid edges."""
    lres, lsize, res, beams, off = WINDOW_CASES[case]
    w = synth.config1()
    p = dict(ndt_resolution=res, search_angular_resolution=0.005, search_angular_size=0.03,
             search_linear_resolution=lres, search_linear_size=lsize, laser_max_beams=beams, range_max=10.0)
    m = ScanMatcherNDT.from_params(p, kernel_variant=5)
    mo = o.new_matcher(p)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    mo.add_scans(w.map_poses, w.map_offsets, w.map_points)
    guess = w.true_pose - np.array(off)
    so, do, wo, co, scores_o = mo.match_scan(guess, w.query_points, want_scores=True)
    sg, dg, wg, cg, _ = m.match_scan_raw(guess, w.query_points)
    if not wo:
        assert not wg and sg == so
        assert np.all(m.dump_scores(guess, w.query_points) == 0.0)
        return
    check_match((sg, dg, wg, cg), (so, do, wo, co), m.dump_scores(guess, w.query_points), scores_o)


def test_theta_sliced_search_matches_full(o):
    """Partial searches over theta ranges + one combine == the full search (the
    multi-GPU path, exercised on one device)."""
    w = synth.config1()
    m = ScanMatcherNDT.from_params(w.params)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    full = m.match_scan_raw(w.query_pose, w.query_points)
    na, _ = m.search_shape()
    m.stage_scan(w.query_pose, w.query_points)
    for cuts in ([0, na], [0, 67, 134, na], [0, 0, 1, na - 1, na], [0, 25, 50, 75, 100, 125, 150, 175, na]):
        parts = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            m.search_staged(a, b)
            parts.append(m.fetch_partial())
        parts = np.array(parts)
        assert parts[:, 12].sum() == na * 100
        s, d, written, cov = m.combine_partials(parts)
        assert written == full[2] and np.array_equal(d, full[1])
        np.testing.assert_allclose(s, full[0], rtol=1e-13)
        np.testing.assert_allclose(cov, full[3], rtol=1e-9, atol=1e-12)


def test_theta_interleaved_search_matches_full(o):
    """Strided theta slices (rank, rank + N, ...): the multi-GPU partition of bench.py."""
    from ndt_2d_b200 import sharded
    w = synth.config1()
    m = ScanMatcherNDT.from_params(w.params)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    full = m.match_scan_raw(w.query_pose, w.query_points)
    na, nl = m.search_shape()
    m.stage_scan(w.query_pose, w.query_points)
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            b, e, st = sharded.theta_slices(na, r, world)
            m.search_staged(b, e, stride=st)
            parts.append(m.fetch_partial())
        parts = np.array(parts)
        assert parts[:, 12].sum() == na * nl * nl
        s, d, written, cov = m.combine_partials(parts)
        assert written == full[2] and np.array_equal(d, full[1])
        np.testing.assert_allclose(s, full[0], rtol=1e-13)
        np.testing.assert_allclose(cov, full[3], rtol=1e-9, atol=1e-12)


# ------------------------------------------------------------------ scoring
def test_score_points_and_poses(o):
    w = synth.config1(laser_max_beams=100)
    m = ScanMatcherNDT.from_params(w.params)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    mo = o.new_matcher(w.params)
    mo.add_scans(w.map_poses, w.map_offsets, w.map_points)
    poses = w.true_pose[None, :] + 0.05 * synth.normal(3, 3 * 257).reshape(-1, 3)
    poses = np.concatenate([poses, [[500.0, 500.0, 0.0], list(w.true_pose)]])
    expect = np.array([mo.score_points(w.query_points, p) for p in poses])
    got = m.scorePoses(w.query_points, poses)
    np.testing.assert_allclose(got, expect, rtol=RTOL, atol=ATOL_SCORE)
    assert got[-2] == 0.0
    for p in poses[:5]:
        np.testing.assert_allclose(m.scorePoints(w.query_points, p), mo.score_points(w.query_points, p),
                                   rtol=RTOL, atol=ATOL_SCORE)
    scan = Scan(0, Pose2d(*w.true_pose), w.query_points)
    np.testing.assert_allclose(m.scoreScan(scan), mo.score_points(w.query_points, w.true_pose), rtol=RTOL)
    # NDT::likelihood(ScanPtr): all points, positive, not normalised
    import ctypes as C
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    lk = o.ndt_likelihood_scan(o.matcher_ndt(mo.h), d(np.ascontiguousarray(w.true_pose)),
                               d(w.query_points), w.query_points.shape[0])
    np.testing.assert_allclose(m.likelihoodScan(scan), lk, rtol=RTOL)
    assert lk > 0


# ------------------------------------------------------------------ loop closure batch
def test_match_scan_batch(o):
    w = synth.config3(n_jobs=4)
    m = ScanMatcherNDT.from_params(w.params)
    score, delta, written, cov = m.match_scan_batch(w.job_scan_offsets, w.map_poses, w.map_offsets,
                                                    w.map_points, w.query_poses, w.query_offsets,
                                                    w.query_points)
    mo = o.new_matcher(w.params)
    for j in range(4):
        s0, s1 = int(w.job_scan_offsets[j]), int(w.job_scan_offsets[j + 1])
        mo.reset()
        offs = w.map_offsets[s0:s1 + 1]
        mo.add_scans(w.map_poses[s0:s1], offs - offs[0], w.map_points[int(offs[0]):int(offs[-1])])
        q0, q1 = int(w.query_offsets[j]), int(w.query_offsets[j + 1])
        so, do, wo, co, sc = mo.match_scan(w.query_poses[j], w.query_points[q0:q1], want_scores=True)
        check_match((score[j], delta[j], written[j], cov[j]), (so, do, wo, co), None, sc)
    assert m.match_scan_raw(w.query_poses[0], w.query_points[:10])[4] == L.ERR_NO_MAP


def _batch_against_oracle(o, params, jobs, variant=0):
    """jobs: list of (map_poses, map_offsets, map_points, query_pose, query_points)."""
    so_, mo_, qo_ = [0], [0], [0]
    mposes, mpts, qposes, qpts = [], [], [], []
    for poses, offs, pts, qpose, qp in jobs:
        offs = np.asarray(offs, dtype=np.int64)
        for k in range(poses.shape[0]):
            mposes.append(poses[k])
            mpts.append(pts[offs[k]:offs[k + 1]])
            mo_.append(mo_[-1] + int(offs[k + 1] - offs[k]))
        so_.append(so_[-1] + poses.shape[0])
        qposes.append(qpose)
        qpts.append(qp)
        qo_.append(qo_[-1] + qp.shape[0])
    m = ScanMatcherNDT.from_params(params, kernel_variant=variant)
    score, delta, written, cov = m.match_scan_batch(
        np.array(so_, dtype=np.uint64), np.array(mposes), np.array(mo_, dtype=np.uint64), np.concatenate(mpts),
        np.array(qposes), np.array(qo_, dtype=np.uint64), np.concatenate(qpts))
    mo = o.new_matcher(params)
    for j, (poses, offs, pts, qpose, qp) in enumerate(jobs):
        mo.reset()
        mo.add_scans(poses, np.asarray(offs, dtype=np.uint64), pts)
        s, d, wr, c, sc = mo.match_scan(qpose, qp, want_scores=True)
        check_match((score[j], delta[j], written[j], cov[j]), (s, d, wr, c), None, sc)


def test_match_scan_batch_other_kernel_paths(o):
    """match_scan_batch beyond the local-window shape: a window too wide for the window kernel
    (dense batch kernel), searches above 2e7 (candidate, point) pairs (region batch kernel, planned
    for the whole batch) and maps too large for the one-CTA build (pipelined lanes)."""
    w = synth.config1(laser_max_beams=100)
    two = slice(0, 2)
    offs2 = w.map_offsets[:3]
    small_map = (w.map_poses[two], offs2, w.map_points[:int(offs2[-1])])
    guesses = [w.true_pose - np.array([0.03 * k, -0.02 * k, 0.01 * k]) for k in range(1, 4)]
    # (a) 10 x 10 lattice 0.9 m wide over 0.1 m cells: 11 cells per axis -> dense batch
    p = dict(w.params, ndt_resolution=0.1, search_linear_resolution=0.1, search_linear_size=0.45,
             search_angular_size=0.05)
    _batch_against_oracle(o, p, [small_map + (g, w.query_points) for g in guesses])
    # (b) 40 x 100 x 100 candidates x 100 beams = 4e7 pairs per job -> region batch
    p = dict(w.params, search_linear_resolution=0.01, search_linear_size=0.5, search_angular_size=0.05)
    _batch_against_oracle(o, p, [small_map + (g, w.query_points) for g in guesses[:2]])
    # (c) 17-scan maps (more points than one CTA's 4,096) -> lanes
    poses17 = np.concatenate([w.map_poses, w.map_poses[:7] + np.array([0.05, 0.05, 0.01])])
    sizes = np.diff(w.map_offsets.astype(np.int64))
    offs17 = np.concatenate([[0], np.cumsum(np.concatenate([sizes, sizes[:7]]))])
    pts17 = np.concatenate([w.map_points, w.map_points[:int(w.map_offsets[7])]])
    assert pts17.shape[0] > 4096
    _batch_against_oracle(o, w.params, [(poses17, offs17, pts17, g, w.query_points) for g in guesses[:2]])


def test_close_loop_empty_window_is_skipped_per_candidate(o):
    """rolling == 0 makes candidate 0's window empty (ndt_mapper.cpp:628-631): it scores 0.0 like a
    matcher without a map and is not accepted; the other candidates are matched as usual."""
    w = synth.config1()
    poses, offs, pts = w.map_poses, w.map_offsets.astype(np.int64), w.map_points
    m = ScanMatcherNDT.from_params(w.params)
    mo = o.new_matcher(w.params)
    qp, got, _ = m.close_loop(poses, offs.astype(np.uint64), pts, np.array([0, 3], dtype=np.uint64), 0, 0,
                              -10.0, w.query_pose, w.query_points)
    assert [g["candidate"] for g in got] == [0, 3]
    assert got[0]["score"] == 0.0 and not got[0]["accepted"]
    oo = offs[2:4]
    mo.add_scans(poses[2:3], oo - oo[0], pts[int(oo[0]):int(oo[1])])                # window [2, 3)
    s, *_ = mo.match_scan(w.query_pose, w.query_points)
    np.testing.assert_allclose(got[1]["score"], s, rtol=RTOL, atol=ATOL_SCORE)
    m.close()


def _sequential_loop_closure(mo, poses, offs, pts, candidates, rolling, limit, typical, qpose, qpts):
    """The reference's inner loop (ndt_mapper.cpp:619-671) with the oracle matcher, one
    candidate at a time."""
    qpose = np.array(qpose, dtype=np.float64)
    out, left = [], limit
    for i in candidates:
        i = int(i)
        if offs[i + 1] == offs[i]:
            continue                                                              # :625
        b, e = (i - 1 if i > 0 else i), (i + 1 if i < rolling else i)             # :628-631
        mo.reset()
        o = offs[b:e + 1]
        mo.add_scans(poses[b:e], o - o[0], pts[int(o[0]):int(o[-1])])
        s, d, written, cov, _ = mo.match_scan(qpose, qpts)
        accept = bool(np.isfinite(s) and s < typical)                             # :645
        if accept:
            qpose = (d if written else np.zeros(3)) + qpose                       # :652-655
        out.append(dict(candidate=i, score=s, accepted=accept, pose=qpose.copy(), covariance=cov))
        left -= 1
        if left == 0:
            break
    return qpose, out


@pytest.mark.parametrize("typical,limit", [(-0.05, 5), (-0.3, 6), (-10.0, 3), (0.5, 4), (-10.0, 0)])
def test_close_loop_matches_sequential_reference_loop(o, typical, limit):
    """Speculative batches + exact re-issue after an acceptance == the sequential loop."""
    w = synth.config1()
    poses, offs, pts = w.map_poses.copy(), w.map_offsets.astype(np.int64), w.map_points
    # graph scan 5 has no points: skipped without counting
    keep = np.ones(pts.shape[0], dtype=bool)
    keep[offs[5]:offs[6]] = False
    sizes = np.diff(offs)
    sizes[5] = 0
    offs = np.concatenate([[0], np.cumsum(sizes)])
    pts = np.ascontiguousarray(pts[keep])
    candidates = np.array([3, 5, 8, 0, 9, 6, 1, 2], dtype=np.uint64)
    rolling = 8
    m = ScanMatcherNDT.from_params(w.params)
    mo = o.new_matcher(w.params)
    qp_o, seq = _sequential_loop_closure(mo, poses, offs, pts, candidates, rolling, limit, typical,
                                         w.query_pose, w.query_points)
    qp_g, got, n_batches = m.close_loop(poses, offs.astype(np.uint64), pts, candidates, rolling, limit,
                                        typical, w.query_pose, w.query_points)
    assert [g["candidate"] for g in got] == [s["candidate"] for s in seq]
    assert [g["accepted"] for g in got] == [s["accepted"] for s in seq]
    for g, s in zip(got, seq):
        np.testing.assert_allclose(g["score"], s["score"], rtol=RTOL, atol=ATOL_SCORE)
        assert np.array_equal(g["pose"], s["pose"])                               # identical corrections
        if np.all(np.isfinite(s["covariance"])):
            np.testing.assert_allclose(g["covariance"], s["covariance"], rtol=RTOL,
                                       atol=RTOL * np.abs(s["covariance"]).max())
    assert np.array_equal(qp_g, qp_o)
    acc = [s["accepted"] for s in seq]
    assert n_batches == 1 + sum(acc[:-1])
    assert 5 not in [g["candidate"] for g in got]
    # the model is left empty, like after the reference's last reset/addScans/matchScan it is not reused
    assert m.match_scan_raw(w.query_pose, w.query_points)[4] == L.ERR_NO_MAP


def test_two_handles_from_two_threads(o):
    """local_scan_matcher_ and global_scan_matcher_ live on different threads in the node
    (ndt_mapper.cpp:508-515 vs :634-643): distinct handles are independent (own stream, own
    buffers), one handle serialises its own calls."""
    import threading
    w1, w4 = synth.config1(), synth.config4(scale=0.04)
    m1 = ScanMatcherNDT.from_params(w1.params)
    m4 = ScanMatcherNDT.from_params(w4.params)
    guess4 = w4.true_pose - np.array([0.05, -0.03, 0.06])
    m1.add_scans_raw(w1.map_poses, w1.map_offsets, w1.map_points)
    m4.add_scans_raw(w4.map_poses, w4.map_offsets, w4.map_points)
    ref1 = m1.match_scan_raw(w1.query_pose, w1.query_points)
    ref4 = m4.match_scan_raw(guess4, w4.query_points)
    errors = []

    def worker(m, w, pose, ref, reps, rebuild):
        try:
            for _ in range(reps):
                if rebuild:
                    m.reset()
                    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
                s, d, wr, cov, _ = m.match_scan_raw(pose, w.query_points)
                assert s == ref[0] and np.array_equal(d, ref[1]) and np.array_equal(cov, ref[3])
        except Exception as e:  # surfaced in the main thread
            errors.append(e)

    ts = [threading.Thread(target=worker, args=(m1, w1, w1.query_pose, ref1, 40, True)),
          threading.Thread(target=worker, args=(m4, w4, guess4, ref4, 10, False)),
          threading.Thread(target=worker, args=(m4, w4, guess4, ref4, 10, False))]   # same handle twice
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


# ------------------------------------------------------------------ particle filter
def test_filter_measure_and_statistics(o):
    w = synth.config2(n_side=12, n_particles=700)
    m = ScanMatcherNDT.from_params(w.params)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    mo = o.new_matcher(w.params)
    mo.add_scans(w.map_poses, w.map_offsets, w.map_points)
    P = w.particles.shape[0]
    f = ParticleFilter(100, P)
    assert f.size() == 100 and np.all(f.get_particles()[0] == 0.0)                # ctor state
    f.set_particles(w.particles, np.full(P, 1.0 / P))
    f.measure(m, Scan(0, Pose2d(), w.scan_points))
    raw = B.pf_measure(o, mo, w.particles, w.scan_points)
    wn, mean_o, cov_o = B.pf_update_statistics(o, w.particles, raw, np.zeros((3, 3)))
    pg, wg = f.get_particles()
    assert np.array_equal(pg, w.particles)
    np.testing.assert_allclose(wg, wn, rtol=RTOL, atol=1e-30)
    np.testing.assert_allclose(f.getMean(), mean_o, rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(f.getCovariance(), cov_o, rtol=RTOL, atol=RTOL * np.abs(cov_o).max())
    # cov(2,2) accumulates across calls like the reference (particle_filter.cpp:216)
    f.measure(m, Scan(0, Pose2d(), w.scan_points))
    _, _, cov_o2 = B.pf_update_statistics(o, w.particles, raw, cov_o)
    np.testing.assert_allclose(f.getCovariance()[2, 2], cov_o2[2, 2], rtol=RTOL)
    assert f.getCovariance()[2, 2] > cov_o[2, 2]


@pytest.mark.parametrize("seed,kld_err,kld_z,min_p,max_p", [
    (11, 0.01, 2.3, 500, 5000), (12, 0.05, 1.0, 50, 400), (13, 0.99, 0.01, 50, 100),
    (14, 0.01, 2.3, 10, 37)])
def test_filter_resample(o, seed, kld_err, kld_z, min_p, max_p):
    P = max_p
    particles = np.stack([1.25 + 0.8 * synth.normal(seed, P), 0.5 + 0.8 * synth.normal(seed + 1, P),
                          1.57 + 0.4 * synth.normal(seed + 2, P)], 1)
    weights = synth.uniform(seed + 3, P) + 0.01
    weights /= weights.sum()
    u = synth.uniform(seed + 4, max_p)
    po, wo, idx = B.pf_resample(o, particles, weights, min_p, max_p, kld_err, kld_z, u)
    f = ParticleFilter(min_p, max_p)
    f.set_particles(particles, weights)
    f.set_covariance(np.zeros((3, 3)))
    f.resample(kld_err, kld_z, uniforms=u)
    assert f.size() == po.shape[0]                                                # KLD stop index
    assert np.array_equal(f.last_draws(), idx)                                    # drawn indices
    pg, wg = f.get_particles()
    assert np.array_equal(pg, po)
    wn, mean_o, cov_o = B.pf_update_statistics(o, po, wo, np.zeros((3, 3)))
    np.testing.assert_allclose(wg, wn, rtol=1e-12)
    np.testing.assert_allclose(f.getMean(), mean_o, rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(f.getCovariance(), cov_o, rtol=RTOL, atol=RTOL * np.abs(cov_o).max())
    # device-generated uniforms replay the documented host stream
    f2 = ParticleFilter(min_p, max_p)
    f2.set_particles(particles, weights)
    f2.resample(kld_err, kld_z, uniforms=None, seed=seed + 4)
    assert f2.size() == po.shape[0] and np.array_equal(f2.last_draws(), idx)


def test_filter_init_update_statistical(o):
    """test/particle_tests.cpp:160-204: loose statistical checks (the reference's RNG is
    random_device-seeded, so only distributions can be compared)."""
    from ndt_2d_b200 import MotionModel
    f = ParticleFilter(2000, 4000, MotionModel(0.1, 0.1, 0.1, 0.1, 0.0))
    f.init(1.25, 0.5, 1.57, 0.1, 0.1, 0.3, seed=5)
    mean, cov = f.getMean(), f.getCovariance()
    assert abs(mean[0] - 1.25) < 0.02 and abs(mean[1] - 0.5) < 0.02 and abs(mean[2] - 1.57) < 0.05
    assert abs(cov[0, 0] - 0.01) < 0.003 and abs(cov[1, 1] - 0.01) < 0.003 and abs(cov[2, 2] - 0.09) < 0.02
    f.update(1.5, 0.0, 0.0, seed=6)
    mean = f.getMean()
    # tolerances of the reference's own test (EXPECT_NEAR(..., 0.4)); the expected y is
    # 0.5 + 1.5 * E[sin(theta + rot1 noise)] ~ 1.78, not 2.0
    assert abs(mean[0] - 1.25) < 0.4 and abs(mean[1] - 2.0) < 0.4 and abs(mean[2] - 1.57) < 0.4
    f.resample(0.99, 0.01, seed=7)
    mean = f.getMean()
    assert abs(mean[0] - 1.25) < 0.4 and abs(mean[1] - 2.0) < 0.4
    f.init(0.0, 0.0, 3.14, 0.1, 0.1, 0.1, seed=8)                                  # circular mean
    assert abs(o.shortest_angular_distance(f.getMean()[2], 3.14)) < 0.05
    f.init(0.0, 0.0, 0.0, 0.1, 0.1, 0.1, seed=9)
    f.update(-1.0, 0.0, 0.0, seed=10)
    mean = f.getMean()
    assert abs(mean[0] + 1.0) < 0.2 and abs(mean[1]) < 0.2 and abs(mean[2]) < 0.2
