"""Golden vectors produced by the reference's own sources (tests/golden/make_golden.py,
run where /root/reference exists).  CPU tests pin the oracle restatement to them; GPU
tests pin the CUDA path (through the C ABI) to them with the north_star tolerances:
cell indices / counts / candidate counts bit-exact, means / covariances / scores within
1e-5 relative, identical best pose."""
from pathlib import Path

import numpy as np
import pytest

from oracle import binding as B

GOLDEN = Path(__file__).resolve().parent / "golden"
MATCHER_CASES = ["config1_beams360", "config1_beams100", "plugin_defaults", "config4_window",
                 "bbox_quirk"]
RTOL = 1e-5
ATOL_SCORE = 1e-30

PARAM_KEYS = ("ndt_resolution", "search_angular_resolution", "search_angular_size",
              "search_linear_resolution", "search_linear_size", "laser_max_beams", "range_max")


def load(name):
    g = np.load(GOLDEN / f"{name}.npz")
    params = dict(zip(PARAM_KEYS, g["params"].tolist()))
    params["laser_max_beams"] = int(params["laser_max_beams"])
    return g, params


def dense(g):
    sx, sy = int(g["grid"][0]), int(g["grid"][1])
    cells = np.zeros((sx * sy, 16))
    cells[g["cell_index"]] = g["cell_values"]
    return cells


# ----------------------------------------------------------------------------- CPU: oracle
@pytest.mark.parametrize("name", MATCHER_CASES)
def test_oracle_matches_reference_golden(oracle, name):
    g, params = load(name)
    m = oracle.new_matcher(params)
    m.add_scans(g["map_poses"], g["map_offsets"], g["map_points"])
    assert np.array_equal(np.array(m.grid(), dtype=np.float64), g["grid"])
    assert np.array_equal(m.dump_cells(), dense(g))                               # bit-identical
    na = oracle.loop_values(params["search_angular_size"], params["search_angular_resolution"], None, 0)
    nl = oracle.loop_values(params["search_linear_size"], params["search_linear_resolution"], None, 0)
    assert na * nl * nl == int(g["candidate_count"][0])
    for k in range(int(g["n_queries"][0])):
        s, d, written, cov, _ = m.match_scan(g[f"q{k}_pose"], g[f"q{k}_points"])
        assert s == g[f"q{k}_score"][0] or (np.isnan(s) and np.isnan(g[f"q{k}_score"][0]))
        assert written == bool(g[f"q{k}_written"][0])
        if written:
            assert np.array_equal(d, g[f"q{k}_delta"])
        np.testing.assert_allclose(cov, g[f"q{k}_cov"], rtol=1e-12, equal_nan=True)
    got = np.array([m.score_points(g["q0_points"], p) for p in g["score_poses"]])
    assert np.array_equal(got, g["score_values"])


def test_oracle_particle_filter_golden(oracle):
    g, params = load("particle_filter")
    m = oracle.new_matcher(params)
    m.add_scans(g["map_poses"], g["map_offsets"], g["map_points"])
    raw = B.pf_measure(oracle, m, g["particles"], g["scan_points"])
    w, mean, cov = B.pf_update_statistics(oracle, g["particles"], raw, np.zeros((3, 3)))
    assert np.array_equal(w, g["measured_weights"])
    np.testing.assert_allclose(mean, g["mean"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(cov, g["cov"], rtol=1e-13, atol=1e-18)
    for k in range(int(g["n_resamples"][0])):
        kld_err, kld_z, min_p, max_p = g[f"r{k}_args"]
        po, wo, _ = B.pf_resample(oracle, g["particles"], g["measured_weights"], int(min_p), int(max_p),
                                  kld_err, kld_z, g[f"r{k}_uniforms"])
        assert np.array_equal(po, g[f"r{k}_particles"])
        wn, mean, cov = B.pf_update_statistics(oracle, po, wo, np.zeros((3, 3)))
        assert np.array_equal(wn, g[f"r{k}_weights"])
        np.testing.assert_allclose(mean, g[f"r{k}_mean"], rtol=0, atol=1e-15)
        np.testing.assert_allclose(cov, g[f"r{k}_cov"], rtol=1e-13, atol=1e-18)


def test_oracle_kd_tree_golden(oracle):
    g = np.load(GOLDEN / "kd_tree.npz")
    assert np.array_equal(oracle.kd_counts(g["poses"]), g["counts"])
    assert g["counts"][:5].tolist() == [1, 1, 2, 3, 4]                            # particle_tests.cpp:47-72


# ----------------------------------------------------------------------------- GPU: CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1, 3, 4, 5])
@pytest.mark.parametrize("name", MATCHER_CASES)
def test_device_matches_reference_golden(gpu, name, variant):
    from ndt_2d_b200 import ScanMatcherNDT
    g, params = load(name)
    m = ScanMatcherNDT.from_params(params, kernel_variant=variant)
    m.add_scans_raw(g["map_poses"], g["map_offsets"], g["map_points"])
    assert np.array_equal(np.array(m.grid_info(), dtype=np.float64), g["grid"])   # grid, bit-exact
    gold = dense(g)
    cells = m.dump_cells()
    assert np.array_equal(np.nonzero(cells[:, 1])[0], g["cell_index"])            # cell indices
    assert np.array_equal(cells[:, :2], gold[:, :2])                              # valid, n
    np.testing.assert_allclose(cells[:, 2:12], gold[:, 2:12], rtol=RTOL, atol=0)  # mean, cov, moments
    np.testing.assert_allclose(cells[:, 12:16], gold[:, 12:16], rtol=RTOL, atol=0)
    na, nl = m.search_shape()
    assert na * nl * nl == int(g["candidate_count"][0])                           # candidate count
    for k in range(int(g["n_queries"][0])):
        s, d, written, cov, _ = m.match_scan_raw(g[f"q{k}_pose"], g[f"q{k}_points"])
        np.testing.assert_allclose(s, g[f"q{k}_score"][0], rtol=RTOL, atol=ATOL_SCORE, equal_nan=True)
        assert written == bool(g[f"q{k}_written"][0])
        if written:
            assert np.array_equal(d, g[f"q{k}_delta"])                            # identical best pose
        gc = g[f"q{k}_cov"]
        if np.all(np.isfinite(gc)):
            np.testing.assert_allclose(cov, gc, rtol=RTOL, atol=RTOL * np.abs(gc).max())
        else:
            assert np.array_equal(np.isnan(cov), np.isnan(gc))
    got = m.scorePoses(g["q0_points"], g["score_poses"])
    np.testing.assert_allclose(got, g["score_values"], rtol=RTOL, atol=ATOL_SCORE)


@pytest.mark.gpu
def test_device_particle_filter_golden(gpu):
    from ndt_2d_b200 import ParticleFilter, Pose2d, Scan, ScanMatcherNDT
    g, params = load("particle_filter")
    m = ScanMatcherNDT.from_params(params)
    m.add_scans_raw(g["map_poses"], g["map_offsets"], g["map_points"])
    P = g["particles"].shape[0]
    f = ParticleFilter(50, P)
    f.set_particles(g["particles"], np.full(P, 1.0 / P))
    f.set_covariance(np.zeros((3, 3)))
    f.measure(m, Scan(0, Pose2d(), g["scan_points"]))
    _, w = f.get_particles()
    np.testing.assert_allclose(w, g["measured_weights"], rtol=RTOL, atol=1e-30)
    np.testing.assert_allclose(f.getMean(), g["mean"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(f.getCovariance(), g["cov"], rtol=RTOL, atol=RTOL * np.abs(g["cov"]).max())
    for k in range(int(g["n_resamples"][0])):
        kld_err, kld_z, min_p, max_p = g[f"r{k}_args"]
        f2 = ParticleFilter(int(min_p), int(max_p))
        f2.set_particles(g["particles"], g["measured_weights"])
        f2.set_covariance(np.zeros((3, 3)))
        f2.resample(kld_err, kld_z, uniforms=g[f"r{k}_uniforms"])
        assert f2.size() == g[f"r{k}_particles"].shape[0]                         # KLD stop index
        p2, w2 = f2.get_particles()
        assert np.array_equal(p2, g[f"r{k}_particles"])                           # same draws
        np.testing.assert_allclose(w2, g[f"r{k}_weights"], rtol=1e-12)
        np.testing.assert_allclose(f2.getMean(), g[f"r{k}_mean"], rtol=RTOL, atol=1e-9)
        gc = g[f"r{k}_cov"]
        np.testing.assert_allclose(f2.getCovariance(), gc, rtol=RTOL, atol=RTOL * np.abs(gc).max())
