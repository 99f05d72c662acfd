"""One rank's share of the 8-GPU config-4 search on one GPU (theta slices 0, 8, 16, ...): the launch
list of this script (ncu --metrics gpu__time_duration.sum) shows the fixed per-search costs that
limit scaling.   python profiles/c4_rank_of_8.py [stride]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ndt_2d_b200 import ScanMatcherNDT, synth  # noqa: E402

stride = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w = synth.config4()
m = ScanMatcherNDT.from_params(w.params)
m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
na, nl = m.search_shape()
m.stage_scan(w.query_pose, w.query_points)
for _ in range(3):
    m.search_staged(0, na, stride=stride)
    m.fetch_partial()
t0 = time.perf_counter()
n = 20
for _ in range(n):
    m.search_staged(0, na, stride=stride)
    m.fetch_partial()
print(f"stride {stride}: {(time.perf_counter() - t0) / n * 1e3:.3f} ms per search_staged + fetch; kernel "
      f"{m.search_stats()['kernel_ms']:.3f} ms")
