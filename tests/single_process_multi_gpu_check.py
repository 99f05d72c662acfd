"""Run as ONE plain python process on a box with N >= 2 GPUs (not collected by pytest): one
multi-device handle (ndt2d_params.n_devices, the path the C++ plugin's setDevices / "<name>.n_gpus"
takes) must reproduce the single-GPU search and the committed full-size golden.

    python tests/single_process_multi_gpu_check.py [N]
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ndt_2d_b200 import ScanMatcherNDT, lib, synth  # noqa: E402


def main():
    n_vis = lib.ndt2d_device_count()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else n_vis
    assert 2 <= n <= n_vis, f"needs >= 2 GPUs, {n_vis} visible"
    devices = list(range(n))
    # (1) a reduced window: many searches back to back (sequence numbers / mailbox parities)
    w = synth.config4(scale=0.25)
    one = ScanMatcherNDT.from_params(w.params, device=0)
    many = ScanMatcherNDT.from_params(w.params, devices=devices)
    many.set_group_threshold(1.0e9)          # this reduced window is below the default threshold
    for m in (one, many):
        m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    full = one.match_scan_raw(w.query_pose, w.query_points)
    for rep in range(6):
        s, d, wr, cov, _ = many.match_scan_raw(w.query_pose, w.query_points)
        assert wr == full[2] and np.array_equal(d, full[1]), (rep, d, full[1])
        np.testing.assert_allclose(s, full[0], rtol=1e-6)
        np.testing.assert_allclose(cov, full[3], rtol=1e-6, atol=1e-9 * np.abs(full[3]).max())
    gi = many.group_info()
    assert gi["devices"] == n and gi["group_searches"] == 6, gi
    print(f"config4/4: {n} GPUs behind one handle match the single-GPU search; p2p exchange: {gi['p2p']}")
    # a local match through the same handle stays on devices[0]
    w1 = synth.config1()
    loc1 = ScanMatcherNDT.from_params(w1.params, device=0)
    locn = ScanMatcherNDT.from_params(w1.params, devices=devices)
    for m in (loc1, locn):
        m.add_scans_raw(w1.map_poses, w1.map_offsets, w1.map_points)
    a, b = loc1.match_scan_raw(w1.query_pose, w1.query_points), locn.match_scan_raw(w1.query_pose, w1.query_points)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and locn.group_info()["group_searches"] == 0
    for m in (one, many, loc1, locn):
        m.close()
    # (2) the headline workload against the committed golden (oracle over all 3142 slices)
    g = np.load(ROOT / "tests" / "golden" / "config4_full.npz")
    w = synth.config4()
    many = ScanMatcherNDT.from_params(w.params, devices=devices)
    many.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    for _ in range(3):
        many.match_scan_raw(w.query_pose, w.query_points)
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        s, d, wr, cov, _ = many.match_scan_raw(w.query_pose, w.query_points)
    dt = (time.perf_counter() - t0) / reps
    assert wr and np.array_equal(d, g["delta"])
    np.testing.assert_allclose(s, g["score"][0], rtol=1e-5)
    np.testing.assert_allclose(cov, g["cov"], rtol=1e-5, atol=1e-5 * np.abs(g["cov"]).max())
    st = many.group_search_stats()
    print(f"config 4 (502,720,000 candidates) on {n} GPUs, one process: {dt * 1e3:.3f} ms per matchScan end "
          f"to end, search kernels per device {['%.3f' % x for x in st['kernel_ms']]} ms; golden parity ok")
    many.close()
    print("SINGLE_PROCESS_MULTI_GPU_CHECK_OK")


if __name__ == "__main__":
    main()
