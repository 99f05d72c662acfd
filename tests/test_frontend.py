"""LaserScan -> Scan points (Mapper::laserCallback, ndt_mapper.cpp:385-453).
The reference code lives inside the ROS node and has no test of its own, so the oracle here is the
C restatement alone (parity unpinned); the CPU tests check it against hand-computed cases, the GPU
tests check the device path against it: kept beams and their order exact, coordinates to 1e-12."""
import math

import numpy as np
import pytest

from oracle import binding as B


def scan_msg(n=720, seed=1):
    rng = np.random.default_rng(seed)
    ranges = (1.0 + 20.0 * rng.random(n)).astype(np.float32)
    ranges[rng.random(n) < 0.05] = np.nan
    ranges[rng.random(n) < 0.05] = np.inf
    ranges[7] = 12.0                      # exactly range_max: kept (the test is `> range_max`)
    angle_min = np.float32(-2.35619449)
    inc = np.float32(2.0 * 2.35619449 / (n - 1))
    return ranges, float(angle_min), float(inc)


def test_oracle_hand_cases(oracle):
    # one beam straight ahead, no transform, no motion
    p = B.laser_to_points(oracle, [2.0], 0.0, 0.1, 10.0, [0, 0, 0], [0, 0, 0], False)
    assert p.shape == (1, 2) and p[0, 0] == 2.0 and p[0, 1] == 0.0
    # laser mounted at (0.2, 0.1) rotated by 90 degrees: x axis of the laser = +y of the robot
    p = B.laser_to_points(oracle, [1.0], 0.0, 0.1, 10.0, [0.2, 0.1, math.pi / 2], [0, 0, 0], False)
    np.testing.assert_allclose(p[0], [0.2, 1.1], atol=1e-15)
    # NaN and beyond-range beams dropped, order kept; r == range_max kept
    r = np.array([1.0, np.nan, 5.0, 10.0, 10.5], dtype=np.float32)
    p = B.laser_to_points(oracle, r, 0.0, 0.0, 10.0, [0, 0, 0], [0, 0, 0], False)
    assert p[:, 0].tolist() == [1.0, 5.0, 10.0]
    # de-skew: beam i is moved by i / n of the translation during the scan
    p = B.laser_to_points(oracle, np.ones(4, dtype=np.float32), 0.0, 0.0, 10.0, [0, 0, 0], [0.4, 0.0, 0.0], False)
    np.testing.assert_allclose(p[:, 0], [1.0, 1.1, 1.2, 1.3], atol=1e-15)
    # inverted: backwards from n-1 down to 1 (index 0 never visited), negated angles
    r = np.array([9.0, 1.0, 2.0, 3.0], dtype=np.float32)
    p = B.laser_to_points(oracle, r, 0.0, 0.0, 10.0, [0, 0, 0], [0, 0, 0], True)
    assert p[:, 0].tolist() == [3.0, 2.0, 1.0]
    p = B.laser_to_points(oracle, np.ones(3, dtype=np.float32), 0.0, np.float32(math.pi / 2), 10.0,
                          [0, 0, 0], [0, 0, 0], True)
    np.testing.assert_allclose(p, [[-1.0, 0.0], [0.0, -1.0]], atol=1e-6)   # angles -pi, -pi/2 (float pi/2)


@pytest.mark.gpu
@pytest.mark.parametrize("inverted", [False, True])
@pytest.mark.parametrize("n", [1, 2, 33, 720, 1080, 4099])
def test_device_laser_to_points(oracle, gpu, n, inverted):
    from ndt_2d_b200 import laser_to_points
    ranges, amin, inc = scan_msg(max(n, 8), seed=n)
    ranges = ranges[:n]
    laser_tf = [0.21, -0.03, 0.37]
    translation = [0.034, -0.012, 0.021]
    want = B.laser_to_points(oracle, ranges, amin, inc, 12.0, laser_tf, translation, inverted)
    got = laser_to_points(ranges, amin, inc, 12.0, laser_tf, translation, inverted)
    assert got.shape == want.shape                                  # same beams kept
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)   # same order, same coordinates


@pytest.mark.gpu
def test_device_laser_all_filtered(oracle, gpu):
    from ndt_2d_b200 import laser_to_points
    r = np.full(100, np.nan, dtype=np.float32)
    assert laser_to_points(r, 0.0, 0.01, 10.0, [0, 0, 0], [0, 0, 0]).shape == (0, 2)
    assert laser_to_points(np.zeros(0, dtype=np.float32), 0.0, 0.01, 10.0, [0, 0, 0], [0, 0, 0]).shape == (0, 2)
