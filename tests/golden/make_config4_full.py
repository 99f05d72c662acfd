#!/usr/bin/env python
"""Generates tests/golden/config4_full.npz: the reference-held answer for the headline workload.

    python tests/golden/make_config4_full.py slices [--threads 7]   (~7 min on 7 cores)
    python tests/golden/make_config4_full.py reference               (~50 min, one core)
    python tests/golden/make_config4_full.py merge

BASELINE.json configs[3]: 1080-beam scan, +-2 m @0.01 m, +-pi @0.002 rad = 3142 x 400 x 400 =
502,720,000 candidates (scan_matcher_ndt.cpp:103-148).

* `slices`    runs the C restatement (oracle/ndt2d_oracle.c, bit-identical to the compiled reference on
              every case of tests/test_oracle_vs_ref.py) once per theta slice -- orc_matcher_partial, the
              sequential loop restricted to one slice -- on a thread pool, and keeps the 3142 16-double
              records (best score / best global index / k, u, s sums of the slice).
* `reference` runs the REFERENCE ITSELF (oracle/_ref/libndt2d_ref.so: src/scan_matcher_ndt.cpp compiled
              unmodified) through its own matchScan over all 502.7M candidates, single-threaded as it is.
* `merge`     writes config4_full.npz with the inputs, the per-slice records, their sequential fold
              (argmin on strict '<' in loop order, covariance K/s + u u^T/s^2, best/n) and, when the
              reference run has finished, its score / delta / covariance.  The merge asserts that the
              folded slices and the reference agree (same delta; score and covariance to 1e-12: the
              only difference is the association of the k/u/s sums).
"""
from __future__ import annotations

import ctypes as C
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from ndt_2d_b200 import synth  # noqa: E402  (host-only synthetic world generator)
from oracle import binding as B  # noqa: E402

SLICES = HERE / "_config4_slices.npy"
REFOUT = HERE / "_config4_reference.npz"


def workload():
    return synth.config4()


def lattices(o, params):
    na = o.loop_values(params["search_angular_size"], params["search_angular_resolution"], None, 0)
    nl = o.loop_values(params["search_linear_size"], params["search_linear_resolution"], None, 0)
    dth, dlin = np.zeros(na), np.zeros(nl)
    dp = C.POINTER(C.c_double)
    o.loop_values(params["search_angular_size"], params["search_angular_resolution"],
                  dth.ctypes.data_as(dp), na)
    o.loop_values(params["search_linear_size"], params["search_linear_resolution"],
                  dlin.ctypes.data_as(dp), nl)
    return dth, dlin


def fold(records, dth, dlin):
    """Sequential fold of per-slice records in theta order == the reference's loop order."""
    best, best_idx = 0.0, 1.0e300
    sums = np.zeros(10)
    for p in records:
        if p[0] < best:
            best, best_idx = p[0], p[1]
        sums += p[2:12]
    n = records[:, 13].max()
    nl = dlin.shape[0]
    written = best < 0.0
    delta = np.zeros(3)
    if written:
        idx = int(best_idx)
        it, rem = divmod(idx, nl * nl)
        delta[:] = (dlin[rem // nl], dlin[rem % nl], dth[it])
    s = sums[9]
    k = np.array([[sums[0], sums[1], sums[2]], [sums[1], sums[3], sums[4]], [sums[2], sums[4], sums[5]]])
    u = sums[6:9]
    cov = (1.0 / s) * k + np.outer((1.0 / (s * s)) * u, u)
    return (best if written else 0.0) / n, delta, written, cov, sums, best, best_idx


def run_slices(threads: int):
    B.build(quiet=True)
    o = B.load_oracle()
    w = workload()
    na = o.loop_values(w.params["search_angular_size"], w.params["search_angular_resolution"], None, 0)
    out = np.zeros((na, 16))
    done = [0]
    t0 = time.time()

    def work(chunk):
        m = o.new_matcher(w.params)          # one model per thread: the oracle is not re-entrant per handle
        m.add_scans(w.map_poses, w.map_offsets, w.map_points)
        for i in chunk:
            out[i] = m.partial(w.query_pose, w.query_points, i, i + 1)
            done[0] += 1
            if done[0] % 100 == 0:
                print(f"  {done[0]}/{na} slices, {time.time() - t0:.0f} s", flush=True)
        m.close()

    chunks = [list(range(t, na, threads)) for t in range(threads)]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, chunks))
    np.save(SLICES, out)
    print(f"slices: {na} records in {time.time() - t0:.0f} s -> {SLICES.name}")


def run_reference():
    B.build(quiet=True)
    r = B.load_ref()
    if r is None:
        raise SystemExit("oracle/_ref/libndt2d_ref.so missing: /root/reference is needed")
    w = workload()
    m = r.new_matcher(w.params)
    m.add_scans(w.map_poses, w.map_offsets, w.map_points)
    t0 = time.time()
    s, d, written, cov, _ = m.match_scan(w.query_pose, w.query_points)
    dt = time.time() - t0
    np.savez(REFOUT, score=np.array([s]), delta=d, written=np.array([int(written)]), cov=cov,
             seconds=np.array([dt]), candidates=np.array([r.matcher_candidate_count(m.h)], dtype=np.uint64))
    print(f"reference: score {s!r} delta {d.tolist()} in {dt:.0f} s -> {REFOUT.name}")


def merge():
    o = B.load_oracle()
    w = workload()
    rec = np.load(SLICES)
    dth, dlin = lattices(o, w.params)
    score, delta, written, cov, sums, best_sum, best_idx = fold(rec, dth, dlin)
    out = dict(
        params=np.array([w.params[k] for k in synth.PARAM_KEYS], dtype=np.float64),
        map_poses=w.map_poses, map_offsets=w.map_offsets.astype(np.uint64), map_points=w.map_points,
        query_pose=w.query_pose, query_points=w.query_points,
        dth=dth, dlin=dlin, slice_records=rec,
        score=np.array([score]), delta=delta, written=np.array([int(written)]), cov=cov,
        sums=sums, best_sum=np.array([best_sum]), best_index=np.array([best_idx]),
        candidates=np.array([rec[:, 12].sum()], dtype=np.uint64), n_points=np.array([rec[:, 13].max()]),
        has_reference=np.array([0]))
    print(f"slices: score {score!r} delta {delta.tolist()} candidates {int(rec[:, 12].sum())}")
    if REFOUT.exists():
        ref = np.load(REFOUT)
        assert int(ref["written"][0]) == int(written)
        assert np.array_equal(ref["delta"], delta), (ref["delta"], delta)
        np.testing.assert_allclose(ref["score"][0], score, rtol=1e-12)
        np.testing.assert_allclose(ref["cov"], cov, rtol=1e-9, atol=1e-12 * np.abs(cov).max())
        assert int(ref["candidates"][0]) == int(rec[:, 12].sum())
        out.update(has_reference=np.array([1]), ref_score=ref["score"], ref_delta=ref["delta"],
                   ref_written=ref["written"], ref_cov=ref["cov"], ref_seconds=ref["seconds"])
        print(f"reference: score {ref['score'][0]!r} in {ref['seconds'][0]:.0f} s -- agrees with the fold")
    else:
        print("reference run not finished: has_reference = 0")
    np.savez_compressed(HERE / "config4_full.npz", **out)
    print(f"-> {HERE / 'config4_full.npz'}")


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else ""
    if mode == "slices":
        run_slices(int(sys.argv[sys.argv.index("--threads") + 1]) if "--threads" in sys.argv else 7)
    elif mode == "reference":
        run_reference()
    elif mode == "merge":
        merge()
    else:
        raise SystemExit(__doc__)
