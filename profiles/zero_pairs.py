"""Experiment (variant build -DNDT2D_COUNT_ZERO_PAIRS, see profiles/ab_variant.py): how many of the row
pairs the region kernel evaluates are zero on every lane (all 64 exponents below -126), and how many
phases consist of such pairs only.

    python profiles/ab_variant.py zero search_region.cu -DNDT2D_COUNT_ZERO_PAIRS
    NDT2D_B200_LIB=ndt_2d_b200/lib/ab/libndt2d_b200_zero.so python profiles/zero_pairs.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ndt_2d_b200 import ScanMatcherNDT, synth  # noqa: E402

for wl in (synth.config4(), synth.config4_dense()):
    m = ScanMatcherNDT.from_params(wl.params)
    m.add_scans_raw(wl.map_poses, wl.map_offsets, wl.map_points)
    na, nl = m.search_shape()
    m.set_tallies(True)
    m.stage_scan(wl.query_pose, wl.query_points)
    m.search_staged(0, na, stride=8)
    m.fetch_partial()
    st = m.search_stats()
    a, b = st["useful_evaluations"], st["items"]   # the variant build reuses the two tallies
    M = (1 << 40) - 1
    print(wl.name, "pairs", b & M, "zero pairs", a & M, "phases", b >> 40, "zero phases", a >> 40,
          "zero pair frac %.3f zero phase frac %.3f" % ((a & M) / (b & M), (a >> 40) / max(b >> 40, 1)))
