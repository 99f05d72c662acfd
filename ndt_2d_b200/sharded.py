"""Theta-sliced search over several ranks (one process per GPU, torch.distributed).

The correlative search shards by theta slice: the model and the scan are tiny and
replicated, rank r scores the slices [theta_range(n_ang, r, world)), and ONE exchange
-- an all-gather of a 128-byte partial record per rank -- precedes a lexicographic
(score, candidate index) reduce + covariance sums, so the result is what a single
sequential search returns (first-wins argmin included).  The reference has no
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
            exchange_partials(self.mine.clone(), self.gathered, self.group)

    def result(self):
        """-> (score, delta, written, cov) of the whole search; every rank holds the same."""
        if self.exchange == "p2p":
            return self.m.fetch_result()
        return self.m.combine_device(self.gathered.data_ptr(), self.world)

    def match_scan(self, pose, points):
        self.m.stage_scan(pose, points)
        self.search_staged()
        return self.result()
