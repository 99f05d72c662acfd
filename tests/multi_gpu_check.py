"""Run under torchrun on N >= 2 GPUs (not collected by pytest): the fused peer-to-peer exchange
and the NCCL all-gather path must both reproduce the single-GPU search.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29541 tests/multi_gpu_check.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ndt_2d_b200 import ScanMatcherNDT, sharded, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    for name, w in (("config1", synth.config1()), ("config4/25", synth.config4(scale=0.04))):
        m = ScanMatcherNDT.from_params(w.params, device=local, stream=stream.cuda_stream)
        m.add_scans_raw(w.map_poses, w.map_offsets, w.map_points)
        full = m.match_scan_raw(w.query_pose, w.query_points)
        for mode in ("nccl", "p2p"):
            ss = sharded.ShardedSearch(m, rank, world, dev, exchange=mode)
            assert ss.exchange == mode, ss.exchange
            for rep in range(5):                               # several searches: sequence numbers / parities
                s, d, wr, cov = ss.match_scan(w.query_pose, w.query_points)
                assert wr == full[2] and np.array_equal(d, full[1]), (mode, rep, d, full[1])
                # a sub-range may be planned with other regions / point chunks: the float block
                # sums then round differently (~1e-8), far inside the 1e-5 contract
                np.testing.assert_allclose(s, full[0], rtol=1e-6)
                np.testing.assert_allclose(cov, full[3], rtol=1e-6, atol=1e-9 * np.abs(full[3]).max())
        m.close()
        if rank == 0:
            print(f"{name}: nccl and p2p exchanges match the single-GPU search on {world} ranks")
    # loop-closure batch: jobs interleaved over the ranks, one all-gather of the result rows
    w = synth.config3(n_jobs=13)
    m = ScanMatcherNDT.from_params(w.params, device=local, stream=stream.cuda_stream)
    args = (w.job_scan_offsets, w.map_poses, w.map_offsets, w.map_points, w.query_poses, w.query_offsets,
            w.query_points)
    full = m.match_scan_batch(*args)
    got = sharded.ShardedBatch(m, rank, world, dev).match_scan_batch(*args)
    assert np.array_equal(got[2], full[2]) and np.array_equal(got[1], full[1])
    np.testing.assert_allclose(got[0], full[0], rtol=1e-6)
    np.testing.assert_allclose(got[3], full[3], rtol=1e-6, atol=1e-9 * np.nanmax(np.abs(full[3])))
    m.close()
    if rank == 0:
        print(f"config3/13 jobs: the batch split over {world} ranks matches the single-GPU batch")
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK_OK")


if __name__ == "__main__":
    main()
