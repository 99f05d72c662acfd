"""Known-answer tests of the reference's own gtests, replayed on the oracle
restatement (orc) and on the reference's sources compiled in place (ref).

test/ndt_model_tests.cpp:32-230 and test/particle_tests.cpp:47-72 of the reference.
"""
import math

import numpy as np
import pytest


def _cell(lib, pts):
    c = lib.cell_new()
    for x, y in pts:
        lib.cell_add_point(c, x, y)
    return c


def test_ndt_cell(either):
    """test/ndt_model_tests.cpp:32-105"""
    lib = either
    c = _cell(lib, [(3.5, 3.5), (3.5, 3.5), (3.4, 3.45), (3.6, 3.55)])
    assert lib.cell_state(c)[0] == 0.0           # EXPECT_FALSE(cell.valid)
    lib.cell_compute(c)
    st = lib.cell_state(c)
    assert st[0] == 1.0                          # valid
    assert st[2] == 3.5 and st[3] == 3.5         # EXPECT_DOUBLE_EQ mean
    assert abs(lib.cell_score(c, 3.5, 3.5)) <= 0.001   # n < 5 -> 0
    lib.cell_add_point(c, 3.6, 3.45)
    lib.cell_add_point(c, 3.4, 3.55)
    lib.cell_compute(c)
    st = lib.cell_state(c)
    assert abs(st[4] - 0.008) <= 0.001           # covariance(0,0)
    assert abs(st[5] - 0.0) <= 0.001             # covariance(0,1)
    assert abs(st[7] - 0.002) <= 0.001           # covariance(1,1)
    assert abs(lib.cell_score(c, 3.5, 3.5) - 1.0) <= 0.001
    assert abs(lib.cell_score(c, 3.5 + math.sqrt(0.008), 3.5) - 0.6065) <= 0.001
    assert abs(lib.cell_score(c, 3.5 + 2 * math.sqrt(0.008), 3.5) - 0.1353) <= 0.001
    assert abs(lib.cell_score(c, 3.5, 3.5 + math.sqrt(0.002)) - 0.6065) <= 0.001
    assert abs(lib.cell_score(c, 3.5, 3.5 + 2 * math.sqrt(0.002)) - 0.1353) <= 0.001
    assert abs(lib.cell_score(c, 0.0, 0.0)) <= 0.001
    lib.cell_free(c)


def test_ndt_cell_no_x_variation(either):
    """test/ndt_model_tests.cpp:107-147"""
    lib = either
    c = _cell(lib, [(3.5, 3.5)])
    corr = lib.cell_state(c)[8:12]
    assert list(corr) == [12.25, 12.25, 0.0, 12.25]   # (1,0) never written
    for p in [(3.5, 3.45)] * 2 + [(3.5, 3.55)] * 2:
        lib.cell_add_point(c, *p)
    corr = lib.cell_state(c)[8:12]
    assert corr[0] == 12.25 and corr[1] == 12.25 and corr[2] == 0.0
    assert corr[3] == pytest.approx(12.252, rel=4e-16)  # EXPECT_DOUBLE_EQ (4 ulp)
    lib.cell_compute(c)
    st = lib.cell_state(c)
    assert st[2] == 3.5 and st[3] == pytest.approx(3.5, rel=4e-16)
    assert st[4] == 0.0 and st[5] == 0.0 and st[6] == 0.0
    assert abs(st[7] - 0.0025) <= 1e-6
    assert abs(st[12] - 400000.0) <= 1e-6             # information(0,0): clamp branch
    assert st[13] == 0.0 and st[14] == 0.0 and st[15] == 0.0
    lib.cell_free(c)


def test_ndt_cell_no_y_variation(either):
    """test/ndt_model_tests.cpp:149-189"""
    lib = either
    c = _cell(lib, [(3.5, 3.5)])
    for p in [(3.45, 3.5)] * 2 + [(3.55, 3.5)] * 2:
        lib.cell_add_point(c, *p)
    corr = lib.cell_state(c)[8:12]
    assert corr[0] == pytest.approx(12.252, rel=4e-16)
    assert corr[1] == 12.25 and corr[2] == 0.0 and corr[3] == 12.25
    lib.cell_compute(c)
    st = lib.cell_state(c)
    assert st[2] == pytest.approx(3.5, rel=4e-16) and st[3] == 3.5
    assert abs(st[4] - 0.0025) <= 1e-6
    assert st[5] == 0.0 and st[6] == 0.0 and st[7] == 0.0
    assert st[12] == 0.0 and st[13] == 0.0 and st[14] == 0.0
    assert abs(st[15] - 400000.0) <= 1e-6
    lib.cell_free(c)


def _d(a):
    import ctypes as C
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_ndt(either):
    """test/ndt_model_tests.cpp:191-230: NDT(1,10,10,-5,-5), 5 points, likelihood 0.7659"""
    lib = either
    ndt = lib.ndt_create(1.0, 10.0, 10.0, -5.0, -5.0)
    info = np.zeros(5)
    lib.ndt_grid(ndt, _d(info))
    assert (info[0], info[1]) == (11, 11)            # size_t(10/1 + 1)
    pose = np.zeros(3)
    pts = np.array([[3.5, 3.5], [3.45, 3.4], [3.55, 3.6], [3.45, 3.6], [3.45, 3.6]])
    lib.ndt_add_scan(ndt, _d(pose), _d(pts), 5)
    lib.ndt_compute(ndt)
    q = np.array([[3.5, 3.5]])
    score = lib.ndt_likelihood_points(ndt, _d(q), 1)
    assert abs(score - 0.7659) <= 0.001
    assert lib.ndt_get_index(ndt, 3.5, 3.5) == 8 * 11 + 8
    assert lib.ndt_get_index(ndt, -5.5, 0.0) == -1
    assert lib.ndt_get_index(ndt, 0.0, 6.5) == -1
    # likelihood(ScanPtr) at the identity pose == sum over points, positive
    s2 = lib.ndt_likelihood_scan(ndt, _d(pose), _d(q), 1)
    assert s2 == score
    lib.ndt_destroy(ndt)


def test_kd_tree(either):
    """test/particle_tests.cpp:47-72: truncation-toward-zero bins, leaf counts 1,1,2,3,4"""
    poses = np.array([[0, 0, 0], [0, 0, 0], [0.75, 0, 0], [-0.75, 0, 0], [0.75, 0.75, 0]], float)
    counts = either.kd_counts(poses, sizes=(0.5, 0.5, 0.25))
    assert list(counts) == [1, 1, 2, 3, 4]
    # -0.4 and +0.4 truncate to the same (double-width) bin 0
    counts = either.kd_counts(np.array([[0.4, 0, 0], [-0.4, 0, 0], [-0.6, 0, 0]]), sizes=(0.5, 0.5, 0.25))
    assert list(counts) == [1, 1, 2]


def test_plugin_defaults(ref):
    """scan_matcher_ndt.cpp:37-44 defaults, read back from the compiled reference."""
    if ref is None:
        pytest.skip("oracle/_ref not built")
    out = np.zeros(6)
    ref.matcher_defaults(_d(out))
    assert list(out) == [0.25, 0.0025, 0.1, 0.005, 0.05, 100.0]


def test_loop_counts(oracle):
    """Accumulated-double loop bounds (scan_matcher_ndt.cpp:103,117,119): SURVEY.md section 6."""
    assert oracle.loop_values(0.05, 0.005, None, 0) == 21       # plugin default, not 20
    assert oracle.loop_values(0.1, 0.0025, None, 0) == 80
    assert oracle.loop_values(0.25, 0.05, None, 0) == 10
    assert oracle.loop_values(0.25, 0.0025, None, 0) == 200
    assert oracle.loop_values(2.0, 0.01, None, 0) == 400
    assert oracle.loop_values(math.pi, 0.002, None, 0) == 3142
    v = np.zeros(10)
    oracle.loop_values(0.25, 0.05, _d(v), 10)
    assert v[5] == -1.3877787807814457e-17                     # the accumulated "zero"


def test_angles(either):
    assert either.normalize_angle(0.5) == pytest.approx(0.5)
    assert either.normalize_angle(4.0) == pytest.approx(4.0 - 2 * math.pi)
    assert either.shortest_angular_distance(3.0, -3.0) == pytest.approx(2 * math.pi - 6.0)
