"""One rank's share of the 8-GPU config-4 search on one GPU (theta slices r, r + 8, r + 16, ...): the
launch list of this script (ncu --metrics gpu__time_duration.sum) shows the fixed per-search costs
that limit scaling.   python profiles/c4_rank_of_8.py [stride] [--all-ranks]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ndt_2d_b200 import ScanMatcherNDT, synth  # noqa: E402

stride = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ranks = range(stride) if "--all-ranks" in sys.argv else (0,)
w = synth.config4()
m = ScanMatcherNDT.from_params(w.params)
m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
na, nl = m.search_shape()
m.stage_scan(w.query_pose, w.query_points)
total = 0.0
for r in ranks:
    for _ in range(3):
        m.search_staged(r, na, stride=stride)
        m.fetch_partial()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        m.search_staged(r, na, stride=stride)
        m.fetch_partial()
    k = m.search_stats()['kernel_ms']
    total += k
    print(f"stride {stride} rank {r}: {(time.perf_counter() - t0) / n * 1e3:.3f} ms per search_staged + fetch; "
          f"kernel {k:.3f} ms")
if len(ranks) > 1:
    print(f"sum of the ranks' kernels {total:.3f} ms")
