"""Catalog sharding across the GPUs of one box (BASELINE.json north_star item 4).

One process per GPU (torch.distributed, NCCL over NVLink). The catalog rows are split into
contiguous ranges -- rank r owns ids [r*ceil(nb/G), (r+1)*ceil(nb/G)) -- every rank sees all
queries, runs the exact top-k kernel on its shard with global ids, the per-shard (D, I) are
all-gathered and a k-way merge kernel (nrb_merge_topk, K4) produces the global top-k on every
rank. The merged result equals the single-index answer up to the order of exact ties.

The reference has no distributed code at all; this is the north star's extension
of IndexFlat.search.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(nb: int, world: int, rank: int) -> tuple[int, int]:
    """Balanced contiguous row ranges: shard sizes differ by at most one row and no shard is empty
    while nb >= world (ceil-sized ranges leave trailing shards empty, e.g. nb = 9, world = 8)."""
    return nb * rank // world, nb * (rank + 1) // world


def _gpu_merge(Dp: torch.Tensor, Ip: torch.Tensor, metric: int):
    from ._lib import check, lib
    G, nq, k = Dp.shape
    D = torch.empty((nq, k), dtype=torch.float32, device=Dp.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=Dp.device)
    check(lib.nrb_merge_topk(Dp.data_ptr(), Ip.data_ptr(), G, nq, k, metric, D.data_ptr(), I.data_ptr(),
                             torch.cuda.current_stream().cuda_stream), "merge_topk")
    return D, I


class ShardedIndexFlat:
    """Row-sharded exact index. `make_index(d, metric)` builds the per-rank index (default: the
    CUDA IndexFlat) and `merge(Dp, Ip, metric)` merges gathered parts (default: the K4 kernel);
    both are injectable so the host-side logic can be exercised with gloo on CPU."""

    def __init__(self, d: int, metric: int = 1, group=None, make_index=None, merge=None):
        self.d, self.metric_type, self.group = d, metric, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if make_index is None:
            from .faiss import IndexFlat
            make_index = IndexFlat
        self.local = make_index(d, metric)
        self.merge = merge or _gpu_merge
        self.ntotal = 0
        self.id_base = 0

    def add_global(self, x):
        """Every rank passes the SAME full matrix; each keeps only its own row range."""
        nb = x.shape[0]
        assert self.ntotal == 0, "add_global is a one-shot build"
        lo, hi = shard_range(nb, self.world, self.rank)
        self.id_base = lo
        self.local.add(x[lo:hi])
        self.ntotal = nb

    def search_local(self, q, k: int):
        """Per-shard top-k with global ids. q: packed queries (CUDA path) or an array."""
        if hasattr(self.local, "search_packed") and not isinstance(q, (np.ndarray, torch.Tensor)):
            return self.local.search_packed(q, k, self.id_base)
        D, I = self.local.search(q, k)
        I = torch.as_tensor(I)
        I = torch.where(I >= 0, I + self.id_base, I)
        return torch.as_tensor(D), I

    def exchange(self, D: torch.Tensor, I: torch.Tensor):
        """all-gather of the per-shard results -> [G, nq, k] on every rank."""
        if self.world == 1:
            return D.unsqueeze(0), I.unsqueeze(0)
        nq, k = D.shape
        Dp = torch.empty((self.world * nq, k), dtype=D.dtype, device=D.device)
        Ip = torch.empty((self.world * nq, k), dtype=I.dtype, device=I.device)
        dist.all_gather_into_tensor(Dp, D.contiguous(), group=self.group)
        dist.all_gather_into_tensor(Ip, I.contiguous(), group=self.group)
        return Dp.view(self.world, nq, k), Ip.view(self.world, nq, k)

    def search(self, q, k: int):
        D, I = self.search_local(q, k)
        Dp, Ip = self.exchange(D, I)
        if self.world == 1:
            return D, I
        return self.merge(Dp, Ip, self.metric_type)
