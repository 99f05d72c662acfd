import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ndt_2d_b200 import ScanMatcherNDT, synth
for wl in (synth.config4(), synth.config4_dense()):
    m = ScanMatcherNDT.from_params(wl.params)
    m.add_scans_raw(wl.map_poses, wl.map_offsets, wl.map_points)
    na, nl = m.search_shape()
    m.stage_scan(wl.query_pose, wl.query_points)
    m.search_staged(0, na, stride=8)
    m.fetch_partial()
    st = m.search_stats()
    a, b = st['useful_evaluations'], st['items']
    M = (1 << 40) - 1
    print(wl.name, 'pairs', b & M, 'zero pairs', a & M, 'phases', b >> 40, 'zero phases', a >> 40,
          'zero pair frac %.3f zero phase frac %.3f' % ((a & M) / (b & M), (a >> 40) / (b >> 40)))
