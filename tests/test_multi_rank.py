"""The N > 1 host logic of the theta-sliced search on CPU: world_size-2 (and 3) gloo process
groups; each rank contributes the partial record of its theta range (computed by the oracle,
since there is no GPU here), the product's exchange + host combine must reproduce the
sequential search of the whole lattice -- same pose (first-wins across rank boundaries),
same score, covariance sums within rounding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ndt_2d_b200 import sharded, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    from oracle import binding as B
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = synth.config1(laser_max_beams=100)
        o = B.load_oracle()
        mo = o.new_matcher(w.params)
        mo.add_scans(w.map_poses, w.map_offsets, w.map_points)
        dth = sharded.lattice(w.params["search_angular_size"], w.params["search_angular_resolution"])
        dlin = sharded.lattice(w.params["search_linear_size"], w.params["search_linear_resolution"])
        lo, hi = sharded.theta_range(len(dth), rank, world)
        mine = torch.from_numpy(mo.partial(w.query_pose, w.query_points, lo, hi))
        gathered = torch.zeros(world * sharded.PARTIAL_DOUBLES, dtype=torch.float64)
        sharded.exchange_partials(mine, gathered)
        score, delta, written, cov = sharded.combine_host(dth, dlin, gathered.numpy())
        so, do, wo, co, _ = mo.match_scan(w.query_pose, w.query_points)
        assert written == wo and np.array_equal(delta, do), (delta, do)
        assert score == so
        np.testing.assert_allclose(cov, co, rtol=1e-9, atol=1e-12 * np.abs(co).max())
        parts = gathered.numpy().reshape(world, -1)
        assert parts[:, 12].sum() == len(dth) * len(dlin) ** 2        # every candidate exactly once
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.concatenate([[score], delta, cov.ravel()]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_theta_sliced_search_gloo(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"rank{r}.npy") for r in range(world)]
    for r in res[1:]:
        assert np.array_equal(r, res[0])                                # every rank holds the result


def test_theta_slices_interleave_exactly():
    for n_ang in (0, 1, 7, 80, 3142):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                b, e, st = sharded.theta_slices(n_ang, r, world)
                idx = list(range(b, e, st))
                assert len(idx) == sharded.n_slices(b, e, st)
                seen += idx
            assert sorted(seen) == list(range(n_ang))


def test_theta_range_partitions_exactly():
    for n_ang in (0, 1, 7, 80, 200, 3142):
        for world in (1, 2, 3, 4, 8):
            cuts = [sharded.theta_range(n_ang, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n_ang
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))


def test_lattice_replays_accumulated_bounds():
    assert len(sharded.lattice(0.05, 0.005)) == 21          # not 20: accumulated dx reaches 0.0499999...
    assert len(sharded.lattice(0.1, 0.0025)) == 80
    assert len(sharded.lattice(np.pi, 0.002)) == 3142
    assert sharded.lattice(0.05, 0.005)[-1] == 0.04999999999999999


# ------------------------------------------------------------------ loop-closure batch, jobs over ranks
def _batch_worker(rank, world, port, out_dir):
    from oracle import binding as B
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = synth.config3(n_jobs=7)
        o = B.load_oracle()
        mo = o.new_matcher(w.params)

        class OracleBatchMatcher:
            """match_scan_batch of the product's mirror, computed by the oracle (no GPU here)."""
            def match_scan_batch(self, so, mposes, moffs, mpts, qposes, qoffs, qpts):
                n = qposes.shape[0]
                score, delta, written, cov = np.zeros(n), np.zeros((n, 3)), np.zeros(n, bool), np.zeros((n, 3, 3))
                for j in range(n):
                    s0, s1 = int(so[j]), int(so[j + 1])
                    offs = moffs[s0:s1 + 1].astype(np.int64)
                    mo.reset()
                    mo.add_scans(mposes[s0:s1], (offs - offs[0]).astype(np.uint64), mpts[int(offs[0]):int(offs[-1])])
                    q0, q1 = int(qoffs[j]), int(qoffs[j + 1])
                    score[j], d, written[j], cov[j], _ = mo.match_scan(qposes[j], qpts[q0:q1])
                    if written[j]:
                        delta[j] = d
                return score, delta, written, cov

        args = (w.job_scan_offsets, w.map_poses, w.map_offsets, w.map_points, w.query_poses, w.query_offsets,
                w.query_points)
        sb = sharded.ShardedBatch(OracleBatchMatcher(), rank, world, torch.device("cpu"))
        got = sb.match_scan_batch(*args)
        jobs = np.arange(w.query_poses.shape[0])
        want = OracleBatchMatcher().match_scan_batch(*sharded.select_jobs(jobs, *args))
        for g, x in zip(got, want):
            assert np.array_equal(np.nan_to_num(g, nan=-7.0), np.nan_to_num(x, nan=-7.0))   # every job, job order
        np.save(os.path.join(out_dir, f"batch{rank}.npy"), np.nan_to_num(got[0], nan=-7.0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_loop_closure_batch_jobs_over_ranks_gloo(tmp_path, world):
    mp.spawn(_batch_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"batch{r}.npy") for r in range(world)]
    for r in res[1:]:
        assert np.array_equal(r, res[0])


def test_job_slice_covers_every_job_once():
    for n_jobs in (0, 1, 5, 50):
        for world in (1, 2, 3, 8, 64):
            seen = np.concatenate([sharded.job_slice(n_jobs, r, world) for r in range(world)])
            assert sorted(seen.tolist()) == list(range(n_jobs))
