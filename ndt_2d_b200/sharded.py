"""Theta-sliced search over several ranks (one process per GPU, torch.distributed).

The correlative search shards by theta slice: the model and the scan are tiny and
replicated, rank r scores the INTERLEAVED slices r, r + world, ... (theta_slices), and ONE
exchange of a 128-byte partial record per rank -- peer stores from the search's last kernel
into mailboxes in every rank's memory ("p2p"), or an all-gather ("nccl") -- precedes a
lexicographic (score, candidate index) reduce + covariance sums, so the result is what a
single sequential search returns (first-wins argmin included).  (A single process can drive
several GPUs through one handle instead: ndt2d_params.n_devices / ScanMatcherNDT(devices=...).)  The reference has no
multi-device path; this is the N > 1 form of ScanMatcherNDT::matchScan
(scan_matcher_ndt.cpp:76-149).

torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _lib as L

PARTIAL_DOUBLES = 16


def theta_range(n_ang: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of theta indices owned by `rank` (covers [0, n_ang) exactly once)."""
    return (n_ang * rank) // world, (n_ang * (rank + 1)) // world


def theta_slices(n_ang: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(begin, end, stride) of the INTERLEAVED slices of `rank`: rank, rank + world, ...
    How much of the scan overlaps the map -- the cost of a slice -- varies smoothly with
    theta, so interleaving balances the ranks where a contiguous split does not (measured:
    59 % / 41 % of the work at 2 GPUs on config 4)."""
    return min(rank, n_ang), n_ang, world


def n_slices(begin: int, end: int, stride: int) -> int:
    return max(0, (end - begin + stride - 1) // stride)


def lattice(size: float, resolution: float) -> np.ndarray:
    """The reference's accumulated-double loop values (host replay, no device needed)."""
    n = C.c_size_t(0)
    L.check(L.lib.ndt2d_search_lattice(size, resolution, None, 0, C.byref(n)), "ndt2d_search_lattice")
    out = np.zeros(n.value, dtype=np.float64)
    L.check(L.lib.ndt2d_search_lattice(size, resolution, L.dptr(out), out.shape[0], C.byref(n)),
            "ndt2d_search_lattice")
    return out


def exchange_partials(mine, gathered, group=None):
    """The single collective of the sharded search: all-gather of the 16-double records.
    `mine` / `gathered` are torch tensors (float64) of 16 and world*16 elements on the device
    the backend works with (CUDA for NCCL, CPU for gloo)."""
    import torch.distributed as dist
    dist.all_gather_into_tensor(gathered, mine.contiguous(), group=group)
    return gathered


def combine_host(dth: np.ndarray, dlin: np.ndarray, partials: np.ndarray):
    """Host reduction of gathered partial records -> (score, delta[3], written, cov[3,3])."""
    parts = L.f64(partials).reshape(-1, PARTIAL_DOUBLES)
    dth, dlin = L.f64(dth), L.f64(dlin)
    delta, cov = np.zeros(3), np.zeros((3, 3))
    written, score = C.c_int(0), C.c_double(0.0)
    L.check(L.lib.ndt2d_combine_partials_host(L.dptr(dth), dth.shape[0], L.dptr(dlin), dlin.shape[0],
                                              L.dptr(parts), parts.shape[0], L.dptr(delta),
                                              C.byref(written), L.dptr(cov), C.byref(score)),
            "ndt2d_combine_partials_host")
    return float(score.value), delta, bool(written.value), cov


class ShardedSearch:
    """matchScan of one query scan split over the ranks of a process group (GPU ranks).

    exchange = "p2p"  : the search's last kernel stores this rank's record into every rank's
                        mailbox over NVLink (CUDA IPC peer memory), waits for the others and
                        reduces -- no collective call at all (ndt2d_matcher_search_exchange)
    exchange = "nccl" : all-gather of the records through torch.distributed, then a combine kernel
    exchange = "auto" : "p2p" if every rank could map its peers' mailboxes, else "nccl"
    """

    def __init__(self, matcher, rank: int, world: int, device, group=None, exchange: str = "auto"):
        import torch
        self.m, self.rank, self.world, self.group = matcher, rank, world, group
        na, _ = matcher.search_shape()
        self.lo, self.hi, self.stride = theta_slices(na, rank, world)
        self.n_theta = n_slices(self.lo, self.hi, self.stride)
        self.gathered = torch.zeros(world * PARTIAL_DOUBLES, dtype=torch.float64, device=device)
        self.mine = self.gathered[rank * PARTIAL_DOUBLES:(rank + 1) * PARTIAL_DOUBLES]
        self.seq = 0
        self.exchange = "nccl"
        if world > 1 and exchange in ("auto", "p2p"):
            self.exchange = "p2p" if self._connect_p2p(device) else "nccl"
            if exchange == "p2p" and self.exchange != "p2p":
                raise RuntimeError("peer mailboxes could not be mapped on every rank (CUDA IPC)")
        elif world == 1:
            self.exchange = "none"

    def _connect_p2p(self, device) -> bool:
        """All-gathers the mailboxes' IPC handles once and maps them; agreed on by all ranks."""
        import torch
        import torch.distributed as dist
        ok = 1
        try:
            mine = self.m.exchange_init(self.world, self.rank)
        except L.Ndt2dError:
            mine, ok = bytes(64), 0
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=self.group)
        if ok:
            try:
                self.m.exchange_connect(handles)
            except L.Ndt2dError:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return bool(flag.item())

    def search_staged(self):
        """Launch this rank's slices of the staged scan + the exchange (asynchronous)."""
        if self.exchange == "p2p":
            self.seq += 1
            self.m.search_exchange(self.lo, self.hi, self.stride, self.seq)
            return
        self.m.search_staged(self.lo, self.hi, self.mine.data_ptr(), stride=self.stride)
        if self.world > 1:
            # the search ran on the matcher's stream, the collective runs on torch's current
            # stream: order them explicitly in both directions (they may be different streams)
            self._order(self._matcher_stream(), None)
            exchange_partials(self.mine.clone(), self.gathered, self.group)
            self._order(None, self._matcher_stream())

    def _matcher_stream(self):
        import torch
        h = self.m.stream()
        return torch.cuda.ExternalStream(h) if h else torch.cuda.default_stream()

    @staticmethod
    def _order(first, then):
        """Work enqueued on `then` from now on waits for what is on `first` now (None = torch's
        current stream); no-op when they are the same stream."""
        import torch
        first = first if first is not None else torch.cuda.current_stream()
        then = then if then is not None else torch.cuda.current_stream()
        if first.cuda_stream != then.cuda_stream:
            then.wait_event(first.record_event())

    def result(self):
        """-> (score, delta, written, cov) of the whole search; every rank holds the same."""
        if self.exchange == "p2p":
            return self.m.fetch_result()
        return self.m.combine_device(self.gathered.data_ptr(), self.world)

    def match_scan(self, pose, points):
        self.m.stage_scan(pose, points)
        self.search_staged()
        return self.result()


# ----------------------------------------------------------------------------------------------
# Loop-closure batch over several ranks ("for config 3 alternatively split the
# candidates"): the jobs of ndt2d_matcher_match_scan_batch are independent, so rank r takes jobs
# r, r + world, ... and ONE all-gather of the 14-double result rows puts every job's result on
# every rank.  No exchange inside a job.
RESULT_DOUBLES = 14     # score, delta_written, delta[3], covariance[9]


def job_slice(n_jobs: int, rank: int, world: int) -> np.ndarray:
    """Indices of the jobs `rank` runs (interleaved: jobs are sorted by distance, so neighbours cost alike)."""
    return np.arange(min(rank, n_jobs), n_jobs, world, dtype=np.int64)


def select_jobs(jobs, job_scan_offsets, map_poses, map_offsets, map_points, query_poses, query_offsets,
                query_points):
    """The match_scan_batch arguments restricted to `jobs` (offsets rebased, order kept)."""
    so = np.asarray(job_scan_offsets, dtype=np.int64)
    mo = np.asarray(map_offsets, dtype=np.int64)
    qo = np.asarray(query_offsets, dtype=np.int64)
    mp_, qp = L.f64(map_poses).reshape(-1, 3), L.f64(query_poses).reshape(-1, 3)
    mpts, qpts = L.f64(map_points).reshape(-1, 2), L.f64(query_points).reshape(-1, 2)
    out_so, out_mo, out_qo = [0], [0], [0]
    poses, pts, qposes, qs = [], [], [], []
    for j in jobs:
        s0, s1 = int(so[j]), int(so[j + 1])
        for s in range(s0, s1):
            poses.append(mp_[s])
            pts.append(mpts[int(mo[s]):int(mo[s + 1])])
            out_mo.append(out_mo[-1] + int(mo[s + 1] - mo[s]))
        out_so.append(out_so[-1] + (s1 - s0))
        qposes.append(qp[j])
        qs.append(qpts[int(qo[j]):int(qo[j + 1])])
        out_qo.append(out_qo[-1] + int(qo[j + 1] - qo[j]))
    cat = lambda parts, width: np.concatenate(parts) if parts else np.zeros((0, width))   # noqa: E731
    return (np.array(out_so, dtype=np.uint64), cat([p[None] for p in poses], 3).reshape(-1, 3),
            np.array(out_mo, dtype=np.uint64), cat(pts, 2), cat([p[None] for p in qposes], 3).reshape(-1, 3),
            np.array(out_qo, dtype=np.uint64), cat(qs, 2))


def gather_job_rows(rows: np.ndarray, n_jobs: int, rank: int, world: int, device, group=None) -> np.ndarray:
    """All-gather of the ranks' result rows -> [n_jobs, RESULT_DOUBLES] in job order (every rank)."""
    import torch
    import torch.distributed as dist
    per_rank = (n_jobs + world - 1) // world
    mine = torch.zeros(per_rank * RESULT_DOUBLES, dtype=torch.float64, device=device)
    flat = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float64).ravel())
    mine[:flat.numel()] = flat.to(device)
    gathered = torch.zeros(world * per_rank * RESULT_DOUBLES, dtype=torch.float64, device=device)
    if world > 1:
        dist.all_gather_into_tensor(gathered, mine, group=group)
    else:
        gathered.copy_(mine)
    g = gathered.cpu().numpy().reshape(world, per_rank, RESULT_DOUBLES)
    out = np.zeros((n_jobs, RESULT_DOUBLES))
    for r in range(world):
        jobs = job_slice(n_jobs, r, world)
        out[jobs] = g[r, :jobs.shape[0]]
    return out


class ShardedBatch:
    """match_scan_batch with the jobs interleaved over the ranks of a process group."""

    def __init__(self, matcher, rank: int, world: int, device, group=None):
        self.m, self.rank, self.world, self.device, self.group = matcher, rank, world, device, group

    def match_scan_batch(self, job_scan_offsets, map_poses, map_offsets, map_points, query_poses,
                         query_offsets, query_points):
        """-> (score[n], delta[n,3], written[n], cov[n,3,3]) of ALL jobs, on every rank."""
        n_jobs = np.asarray(query_poses).reshape(-1, 3).shape[0]
        jobs = job_slice(n_jobs, self.rank, self.world)
        rows = np.zeros((jobs.shape[0], RESULT_DOUBLES))
        if jobs.shape[0]:
            score, delta, written, cov = self.m.match_scan_batch(*select_jobs(
                jobs, job_scan_offsets, map_poses, map_offsets, map_points, query_poses, query_offsets,
                query_points))
            rows[:, 0], rows[:, 1], rows[:, 2:5], rows[:, 5:14] = score, written, delta, cov.reshape(-1, 9)
        allr = gather_job_rows(rows, n_jobs, self.rank, self.world, self.device, self.group)
        return allr[:, 0].copy(), allr[:, 2:5].copy(), allr[:, 1] != 0.0, allr[:, 5:14].reshape(-1, 3, 3).copy()
