"""A/B helper: latency of the local match (BASELINE config 1) and of the node's per-scan sequence
reset / addScans / scoreScan / matchScan (ndt_mapper.cpp:508-515), many calls, one process.

    python profiles/c1_latency.py [calls]
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ndt_2d_b200 import Pose2d, Scan, ScanMatcherNDT, synth  # noqa: E402

CALLS = int(sys.argv[1]) if len(sys.argv) > 1 else 2000


def dist(fn, calls=CALLS, warm=50):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(calls):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    ts = np.array(ts) * 1e6
    return f"p50 {np.percentile(ts, 50):.1f} us  mean {ts.mean():.1f} us  p99 {np.percentile(ts, 99):.1f} us"


for beams in (360, 100):
    w = synth.config1(laser_max_beams=beams)
    m = ScanMatcherNDT.from_params(w.params)
    scan = Scan(0, Pose2d(*w.query_pose), w.query_points)
    m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
    print(f"beams {beams}: matchScan  {dist(lambda: m.match_scan_raw(w.query_pose, w.query_points))}")
    print(f"beams {beams}: scoreScan  {dist(lambda: m.scoreScan(scan))}")

    def add():
        m.reset()
        m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)

    def add_wait():
        add()
        m.counters()   # (synchronises nothing by itself; the next value-returning call waits)

    def seq():
        add()
        m.scoreScan(scan)
        return m.match_scan_raw(w.query_pose, w.query_points)
    print(f"beams {beams}: reset+addScans (enqueue only)  {dist(add)}")
    print(f"beams {beams}: reset+addScans+scoreScan+matchScan  {dist(seq)}")
    for name, ms in (("matchScan", None), ("reset+addScans+scoreScan+matchScan",
                                           (w.map_poses, w.map_offsets, w.map_points))):
        m.probe_call_latency(w.query_pose, w.query_points, 50, ms)
        us = m.probe_call_latency(w.query_pose, w.query_points, CALLS, ms)
        print(f"beams {beams}: C ABI {name}  p50 {np.percentile(us, 50):.1f} us  mean {us.mean():.1f} us  "
              f"p99 {np.percentile(us, 99):.1f} us")
    m.close()
