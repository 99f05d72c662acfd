"""Occupancy-grid export (ndt_2d::OccupancyGrid, src/occupancy_grid.cpp):
integer hit / empty counters -> the int8 grid must be BIT-EXACT, and so must width, height and origin.
CPU: the C restatement against the compiled reference and the golden file it produced.
GPU: the CUDA path against the restatement and the golden file."""
from pathlib import Path

import numpy as np
import pytest

from ndt_2d_b200 import synth
from oracle import binding as B

GOLDEN = Path(__file__).resolve().parent / "golden" / "occupancy_grid.npz"


def same(a, b):
    (ia, da), (ib, db) = a, b
    assert (ia["width"], ia["height"]) == (ib["width"], ib["height"])
    assert ia["origin_x"] == ib["origin_x"] and ia["origin_y"] == ib["origin_y"]
    assert np.array_equal(da, db)


def sequences():
    """Calls on one instance: growing scan lists (bounds persist), a repeated call, negative coordinates."""
    w = synth.config1()
    yield "grow", [(w.map_poses[:n], w.map_offsets[:n + 1], w.map_points) for n in (3, 3, 7, 10)]
    w4 = synth.config4()
    shift = w4.map_poses.copy()
    shift[:, :2] -= 80.0                      # everything at negative coordinates: max stays 0
    yield "negative", [(shift, w4.map_offsets, w4.map_points)]
    yield "empty", [(np.zeros((0, 3)), np.zeros(1, dtype=np.uint64), np.zeros((0, 2))),
                    (w.map_poses[:2], np.array([0, 0, 0], dtype=np.uint64), np.zeros((0, 2)))]


def test_oracle_matches_compiled_reference(oracle, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built (reference tree not available here)")
    for name, calls in sequences():
        for res, thr in ((0.05, 0.25), (0.1, 0.6)):
            go, gr = B.OccupancyGrid(oracle, res, thr), B.OccupancyGrid(ref, res, thr)
            for poses, offs, pts in calls:
                same(go.get_msg(poses, offs, pts), gr.get_msg(poses, offs, pts))


def test_oracle_matches_golden(oracle):
    g = np.load(GOLDEN)
    grid = B.OccupancyGrid(oracle, 0.05, 0.25)
    for k in range(2):
        n = int(g[f"c{k}_n"][0])
        info, data = grid.get_msg(g["poses"][:n], g["offsets"][:n + 1], g["points"])
        assert [info["width"], info["height"], info["origin_x"], info["origin_y"]] == g[f"c{k}_info"][:4].tolist()
        assert np.array_equal(data, g[f"c{k}_data"])
    assert (g["c1_data"] == 100).sum() > 100 and (g["c1_data"] == 0).sum() > 10000


@pytest.mark.gpu
def test_device_matches_oracle(oracle, gpu):
    from ndt_2d_b200 import OccupancyGrid
    for name, calls in sequences():
        for res, thr in ((0.05, 0.25), (0.1, 0.6)):
            go, gd = B.OccupancyGrid(oracle, res, thr), OccupancyGrid(res, thr)
            for poses, offs, pts in calls:
                same(gd.getMsg(poses, offs, pts), go.get_msg(poses, offs, pts))


@pytest.mark.gpu
def test_device_matches_golden(gpu):
    from ndt_2d_b200 import OccupancyGrid
    g = np.load(GOLDEN)
    grid = OccupancyGrid(0.05, 0.25)
    for k in range(2):
        n = int(g[f"c{k}_n"][0])
        info, data = grid.getMsg(g["poses"][:n], g["offsets"][:n + 1], g["points"])
        assert [info["width"], info["height"], info["origin_x"], info["origin_y"]] == g[f"c{k}_info"][:4].tolist()
        assert np.array_equal(data, g[f"c{k}_data"])


@pytest.mark.gpu
def test_device_full_map(oracle, gpu):
    """The config-2 map: 2,500 scans, 704k rays into a 2,200 x 2,200 grid."""
    from ndt_2d_b200 import OccupancyGrid
    w = synth.config2()
    same(OccupancyGrid(0.05, 0.25).getMsg(w.map_poses, w.map_offsets, w.map_points),
         B.OccupancyGrid(oracle, 0.05, 0.25).get_msg(w.map_poses, w.map_offsets, w.map_points))
