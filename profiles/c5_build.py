"""Profiling helper: the large model build (BASELINE config 5: 20,000 scans, 5.6 M points, 0.1 m cells)."""
import time

import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ndt_2d_b200 import ScanMatcherNDT, synth  # noqa: E402

w = synth.config5()
m = ScanMatcherNDT.from_params(w.params)
for k in range(3):
    t = time.perf_counter()
    m.add_scans_raw(w.poses, w.offsets, w.points)
    print(f"addScans {1e3 * (time.perf_counter() - t):.2f} ms  stats {m.build_stats()}")
