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
