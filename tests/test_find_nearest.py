"""Graph::findNearest (graph.cpp:167-189) + Scan::getBarycenterPose (scan.cpp:55-59, 72-91): the
candidate selection in front of the loop-closure batch.

findNearest runs nanoflann (un-vendored, not installed here) and no reference test calls it, so the
oracle is the restatement of nanoflann's published radius search alone (PARITY UNPINNED): squared L2
accumulated dimension by dimension, dist < radius (a SQUARED radius), nearest first.  The CPU tests
check it on hand cases; the barycenter IS pinned against the reference's own Scan class (oracle/_ref).
The GPU tests check the device path against the oracle: same indices, same order, same distances."""
import numpy as np
import pytest

from oracle import binding as B


def test_oracle_find_nearest_hand_cases(oracle):
    xy = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 2.0], [3.0, 4.0], [0.5, 0.5], [1.0, 0.0]])
    idx, d2 = B.find_nearest(oracle, xy, [0.0, 0.0], 4.0)
    # squared radius 4: (0,2) is at squared distance exactly 4 -> NOT kept (strict <);
    # equal distances (1 and 5) in index order
    assert idx.tolist() == [0, 4, 1, 5] and d2.tolist() == [0.0, 0.5, 1.0, 1.0]
    idx, _ = B.find_nearest(oracle, xy, [0.0, 0.0], np.nextafter(4.0, 5.0))
    assert idx.tolist() == [0, 4, 1, 5, 2]
    # limit_scan_index > 0: only scans [0, limit) (graph.cpp:171); <= 0: all
    idx, _ = B.find_nearest(oracle, xy, [0.0, 0.0], 30.0, limit_scan_index=4)
    assert idx.tolist() == [0, 1, 2, 3]
    idx, _ = B.find_nearest(oracle, xy, [0.0, 0.0], 30.0, limit_scan_index=0)
    assert idx.tolist() == [0, 4, 1, 5, 2, 3]
    # the node's global_search_size 0.2 is a squared radius: 0.44 m is inside, 0.45 m is not
    idx, _ = B.find_nearest(oracle, [[0.44, 0.0], [0.45, 0.0]], [0.0, 0.0], 0.2)
    assert idx.tolist() == [0]
    idx, _ = B.find_nearest(oracle, np.zeros((0, 2)), [0.0, 0.0], 1.0)
    assert idx.size == 0


def test_barycenter_oracle_vs_reference_scan_class(oracle, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(3)
    for n in (0, 1, 7, 360, 1080):
        pose = np.array([rng.normal(0, 30), rng.normal(0, 30), rng.uniform(-3.2, 3.2)])
        pts = rng.normal(0, 5, (n, 2))
        want = B.scan_barycenter(ref, pose, pts)
        assert np.array_equal(B.scan_barycenter(oracle, pose, pts), want)          # bit-identical


def test_mirror_scan_barycenter_matches_oracle(oracle):
    from ndt_2d_b200.scan_matcher import Pose2d, Scan
    rng = np.random.default_rng(4)
    for n in (0, 1, 360):
        pose = np.array([rng.normal(0, 30), rng.normal(0, 30), rng.uniform(-3.2, 3.2)])
        pts = rng.normal(0, 5, (n, 2))
        b = Scan(0, Pose2d(*pose), pts).getBarycenterPose()
        assert np.array_equal([b.x, b.y, b.theta], B.scan_barycenter(oracle, pose, pts))


def graph_positions(n, seed):
    rng = np.random.default_rng(seed)
    xy = rng.uniform(0.0, 100.0, (n, 2))
    if n > 20:
        xy[11] = xy[3]                       # exact duplicates: equal distances, index order
        xy[17] = xy[3]
    return xy


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 31, 257, 20000])
def test_device_find_nearest(oracle, gpu, n):
    from ndt_2d_b200 import find_nearest
    xy = graph_positions(n, seed=n + 1)
    queries = [np.array([50.0, 50.0])] + ([xy[3], xy[n // 2] + 0.01] if n > 20 else [])
    for q in queries:
        for dist in (0.2, 25.0, 1e9):
            for limit in (-1, 0, n // 3, n + 5):
                want_i, want_d = B.find_nearest(oracle, xy, q, dist, limit)
                got_i, got_d = find_nearest(xy, q, dist, limit, return_distances=True)
                assert np.array_equal(got_i, want_i) and np.array_equal(got_d, want_d)


@pytest.mark.gpu
def test_device_find_nearest_boundary_and_capacity(oracle, gpu):
    import ctypes as C
    from ndt_2d_b200 import _lib as L, find_nearest
    xy = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 2.0], [3.0, 4.0], [0.5, 0.5], [1.0, 0.0]])
    assert find_nearest(xy, [0.0, 0.0], 4.0).tolist() == [0, 4, 1, 5]                # strict <
    assert find_nearest(xy, [0.0, 0.0], np.nextafter(4.0, 5.0)).tolist() == [0, 4, 1, 5, 2]
    # capacity smaller than the number of matches: the nearest are written, the count is the total
    idx = np.zeros(2, dtype=np.uint64)
    n = C.c_size_t(0)
    q = np.zeros(2)
    L.check(L.lib.ndt2d_find_nearest(-1, L.dptr(xy), 6, -1, L.dptr(q), 30.0,
                                     idx.ctypes.data_as(C.POINTER(C.c_uint64)), None, 2, C.byref(n)), "find")
    assert n.value == 6 and idx.tolist() == [0, 4]


@pytest.mark.gpu
def test_graph_find_nearest_feeds_close_loop(oracle, gpu):
    """findNearest -> closeLoop, as Mapper::loopClosureThread chains them (ndt_mapper.cpp:612-671)."""
    from ndt_2d_b200 import Pose2d, Scan, ScanMatcherNDT, graph_find_nearest, synth
    w = synth.config3(n_jobs=12)
    scans = []
    for k in range(w.map_poses.shape[0]):
        a, b = int(w.map_offsets[k]), int(w.map_offsets[k + 1])
        scans.append(Scan(k, Pose2d(*w.map_poses[k]), w.map_points[a:b]))
    query = Scan(999, Pose2d(*w.query_poses[0]), w.query_points[int(w.query_offsets[0]):int(w.query_offsets[1])])
    for bary in (False, True):
        cand = graph_find_nearest(scans, query, 25.0, len(scans) - 2, use_barycenter=bary)
        pick = (lambda s: s.getBarycenterPose()) if bary else (lambda s: s.getPose())
        xy = np.array([[pick(s).x, pick(s).y] for s in scans])
        want, _ = B.find_nearest(oracle, xy, [pick(query).x, pick(query).y], 25.0, len(scans) - 2)
        assert np.array_equal(cand, want) and cand.size > 0
    m = ScanMatcherNDT.from_params(w.params)
    pose, out, _ = m.close_loop(w.map_poses, w.map_offsets, w.map_points, cand, len(scans) - 2, 3, -0.01,
                                query.getPose(), query.getPoints())
    assert [o["candidate"] for o in out] == cand[:len(out)].tolist()
