"""Timing helper for the loop-closure batch (BASELINE config 3): wall time of match_scan_batch
against the number of jobs.  Used under ncu for the launch list in profiles/README.md."""
import sys
import time

import numpy as np

from ndt_2d_b200 import ScanMatcherNDT, synth

w = synth.config3()
m = ScanMatcherNDT.from_params(w.params)


def run(n_jobs, reps=20):
    so = w.job_scan_offsets[:n_jobs + 1]
    s1 = int(so[-1])
    mo = w.map_offsets[:s1 + 1]
    qo = w.query_offsets[:n_jobs + 1]
    args = (so, w.map_poses[:s1], mo, w.map_points[:int(mo[-1])], w.query_poses[:n_jobs], qo,
            w.query_points[:int(qo[-1])])
    for _ in range(3):
        m.match_scan_batch(*args)
    t = time.perf_counter()
    for _ in range(reps):
        m.match_scan_batch(*args)
    return (time.perf_counter() - t) / reps * 1e3


for n in ([50] if len(sys.argv) > 1 else [1, 5, 10, 25, 50]):
    print(f"jobs {n:3d}  batch {run(n):.3f} ms")
