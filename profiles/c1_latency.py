"""A/B helper: matchScan latency of the local match (BASELINE config 1), many calls, one process."""
import time

import numpy as np

from ndt_2d_b200 import ScanMatcherNDT, synth

for beams in (360, 100):
    w = synth.config1(laser_max_beams=beams)
    m = ScanMatcherNDT.from_params(w.params)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    for _ in range(50):
        m.match_scan_raw(w.query_pose, w.query_points)
    ts = []
    for _ in range(2000):
        t = time.perf_counter()
        m.match_scan_raw(w.query_pose, w.query_points)
        ts.append(time.perf_counter() - t)
    ts = np.array(ts) * 1e6
    print(f"beams {beams}: matchScan p50 {np.percentile(ts, 50):.1f} us  mean {ts.mean():.1f} us  p99 {np.percentile(ts, 99):.1f} us")
