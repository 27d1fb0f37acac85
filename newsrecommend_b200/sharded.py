"""Catalog sharding across the GPUs of one box (BASELINE.json north_star item 4).

One process per GPU (torch.distributed, NCCL over NVLink / NVSwitch). The catalog rows are split
into contiguous ranges -- rank r owns ids [nb*r/G, nb*(r+1)/G) -- every rank sees all queries and
runs the exact top-k kernel on its shard. The exchange step comes in two forms:

* ``exchange="alltoall"`` (default): all-to-all BY QUERY RANGE. Every rank packs its per-shard
  results as 8-byte (score bits, local row) words (nrb_pack_topk), sends to rank j the rows of the
  queries rank j owns, and merges only its own nq/G queries (nrb_merge_topk_packed, K4). Per rank
  this moves and merges 1/G of what the all-gather does. The batch is processed in chunks; the
  exchange of chunk i runs on NCCL's stream while chunk i+1 is being searched, and every chunk is
  split G ways, so all ranks merge all the time.
* ``exchange="allgather"``: the contract north_star names literally -- all-gather of the packed
  per-shard results and a k-way merge of all queries on every rank.

The merged result equals the single-index answer up to the order of exact ties. The reference has
no distributed code at all; this is the north star's extension of IndexFlat.search.

Host logic only; `make_index` (per-rank index) and `codec` (pack / merge) are injectable so that
the same code runs under gloo on CPU with the oracle index and numpy stand-ins (tests/).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

CHUNK_WAVES = 7  # queries per chunk = CHUNK_WAVES full waves of the search kernel (7 x 18,944)


def shard_range(nb: int, world: int, rank: int) -> tuple[int, int]:
    """Balanced contiguous row ranges: shard sizes differ by at most one row and no shard is empty
    while nb >= world (ceil-sized ranges leave trailing shards empty, e.g. nb = 9, world = 8)."""
    return nb * rank // world, nb * (rank + 1) // world


def chunk_slices(nq: int, world: int, chunk: int) -> list[tuple[int, int, int]]:
    """[(c0, c1, per)]: the batch in chunks of `chunk` queries; inside a chunk rank r owns rows
    [c0 + r*per, min(c1, c0 + (r+1)*per)) -- ceil-sized slices, so that the concatenation of the
    (padded) slices in rank order is the chunk itself (what all_gather_into_tensor produces)."""
    out = []
    for c0 in range(0, nq, chunk):
        c1 = min(nq, c0 + chunk)
        out.append((c0, c1, -(-(c1 - c0) // world)))
    return out


def owned_rows(nq: int, world: int, rank: int, chunk: int) -> list[tuple[int, int]]:
    """Query rows whose final result lands on `rank`, as (lo, hi) spans in batch order."""
    spans = []
    for c0, c1, per in chunk_slices(nq, world, chunk):
        lo, hi = min(c1, c0 + rank * per), min(c1, c0 + (rank + 1) * per)
        spans.append((lo, hi))
    return spans


class GpuCodec:
    """Wire format + K4 merge on the device (libnrb200)."""

    @staticmethod
    def pack(D: torch.Tensor, I: torch.Tensor, id_base: int) -> torch.Tensor:
        from ._lib import check, lib
        P = torch.empty(D.shape, dtype=torch.int64, device=D.device)
        check(lib.nrb_pack_topk(D.data_ptr(), I.data_ptr(), id_base, D.numel(), P.data_ptr(),
                                torch.cuda.current_stream().cuda_stream), "pack_topk")
        return P

    @staticmethod
    def merge(P: torch.Tensor, bases: torch.Tensor, metric: int):
        from ._lib import check, lib
        G, nq, k = P.shape
        D = torch.empty((nq, k), dtype=torch.float32, device=P.device)
        I = torch.empty((nq, k), dtype=torch.int64, device=P.device)
        check(lib.nrb_merge_topk_packed(P.data_ptr(), bases.data_ptr(), G, nq, k, metric, D.data_ptr(), I.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream), "merge_topk_packed")
        return D, I


class ShardedIndexFlat:
    """Row-sharded exact index (IndexFlatIP / IndexFlatL2 semantics over the whole catalog)."""

    def __init__(self, d: int, metric: int = 1, group=None, make_index=None, codec=None, exchange: str = "alltoall",
                 chunk_queries: int | None = None):
        assert exchange in ("alltoall", "allgather")
        self.d, self.metric_type, self.group, self.exchange_mode = d, metric, group, exchange
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._cuda_index = make_index is None
        if make_index is None:
            from .faiss import IndexFlat
            make_index = IndexFlat
        self.local = make_index(d, metric)
        self.codec = codec or GpuCodec
        self.ntotal = 0
        self.id_base = 0
        self.bases = None  # i64[G]: first global id of every shard
        self.chunk_queries = chunk_queries

    # ------------------------------------------------------------------ build
    def add_global(self, x):
        """Every rank passes the SAME full matrix; each keeps only its own row range."""
        nb = x.shape[0]
        assert self.ntotal == 0, "add_global is a one-shot build"
        lo, hi = shard_range(nb, self.world, self.rank)
        self.id_base = lo
        self.local.add(x[lo:hi])
        self.ntotal = nb
        self._bases_host = [shard_range(nb, self.world, r)[0] for r in range(self.world)]
        self.bases = None

    def add_local(self, x_local, id_base: int, ntotal: int):
        """Each rank passes only ITS rows (global ids id_base ... id_base + len - 1): for catalogs
        that no single process ever holds whole (10M x 256 sharded: BASELINE configs[4])."""
        assert self.ntotal == 0, "add_local is a one-shot build"
        self.id_base = int(id_base)
        self.local.add(x_local)
        self.ntotal = int(ntotal)
        if self.world > 1:
            t = torch.tensor([self.id_base], dtype=torch.int64, device=self._dev())
            out = torch.empty(self.world, dtype=torch.int64, device=self._dev())
            dist.all_gather_into_tensor(out, t, group=self.group)
            self._bases_host = out.cpu().tolist()
        else:
            self._bases_host = [self.id_base]
        self.bases = None

    def _dev(self):
        if self._cuda_index:
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def _bases(self):
        if self.bases is None:
            self.bases = torch.tensor(self._bases_host, dtype=torch.int64, device=self._dev())
        return self.bases

    def _chunk(self) -> int:
        if self.chunk_queries:
            return int(self.chunk_queries)
        if self._cuda_index:
            from .faiss import _wave_rows
            return CHUNK_WAVES * _wave_rows()
        return 1 << 17

    # ------------------------------------------------------------------ search pieces
    def search_local(self, xq, k: int):
        """Per-shard top-k of a raw query chunk with GLOBAL ids: (D f32[n,k], I i64[n,k]) tensors."""
        if self._cuda_index:
            from .faiss import PackedMatrix
            q = PackedMatrix.from_tensor(xq, planes=self.local._query_planes(k))  # K0 on the fresh chunk
            return self.local.search_packed(q, k, self.id_base)
        D, I = self.local.search(np.ascontiguousarray(xq), k)
        I = torch.as_tensor(I)
        return torch.as_tensor(D), torch.where(I >= 0, I + self.id_base, I)

    def _exchange_start(self, P: torch.Tensor, per: int):
        """Starts the exchange of one chunk's packed per-shard results P i64[cn, k]. Returns
        (recv i64[G, rows, k], work handle or None): alltoall -> rows = this rank's slice of the
        chunk, allgather -> rows = the whole chunk."""
        cn, k = P.shape
        G, r = self.world, self.rank
        if G == 1:
            return P.unsqueeze(0), None
        if self.exchange_mode == "allgather":
            recv = torch.empty((G * cn, k), dtype=P.dtype, device=P.device)
            work = dist.all_gather_into_tensor(recv, P, group=self.group, async_op=True)
            return recv.view(G, cn, k), work
        in_rows = [max(0, min(cn, (j + 1) * per) - min(cn, j * per)) for j in range(G)]
        mine = in_rows[r]
        recv = torch.empty((G * mine, k), dtype=P.dtype, device=P.device)
        work = dist.all_to_all_single(recv, P, [mine] * G, in_rows, group=self.group, async_op=True)
        return recv.view(G, mine, k), work

    def search(self, xq, k: int, gather: bool = True):
        """xq: the FULL query batch on every rank (CUDA tensor fp32 [nq, d]; ndarray on the CPU test
        path). gather=True: (D, I) of all queries on every rank (a final all-gather of the merged
        slices). gather=False: the merged results of this rank's own rows only, as
        (D, I, spans) with spans = owned_rows(nq, ...)."""
        nq = xq.shape[0]
        G = self.world
        chunk = self._chunk()
        pending = None
        parts = []

        def finish(p):
            recv, work = p
            if work is not None:
                work.wait()
            parts.append(self.codec.merge(recv, self._bases(), self.metric_type))

        for c0, c1, per in chunk_slices(nq, G, chunk):
            D, I = self.search_local(xq[c0:c1], k)
            P = self.codec.pack(D, I, self.id_base)
            nxt = self._exchange_start(P, per)
            if pending is not None:
                finish(pending[:2])  # merged after the NEXT chunk's search was queued: its exchange overlaps that search
            pending = nxt + (P,)  # P stays referenced until its exchange has been waited for
        if pending is not None:
            finish(pending[:2])
        if self.exchange_mode == "allgather" or G == 1:
            D = torch.cat([p[0] for p in parts]) if len(parts) != 1 else parts[0][0]
            I = torch.cat([p[1] for p in parts]) if len(parts) != 1 else parts[0][1]
            if gather:
                return D, I
            spans = owned_rows(nq, G, self.rank, chunk)
            if G == 1:
                return D, I, spans
            sel = torch.cat([torch.arange(lo, hi) for lo, hi in spans]).to(D.device)
            return D[sel], I[sel], spans
        spans = owned_rows(nq, G, self.rank, chunk)
        if not gather:
            D = torch.cat([p[0] for p in parts]) if len(parts) != 1 else parts[0][0]
            I = torch.cat([p[1] for p in parts]) if len(parts) != 1 else parts[0][1]
            return D, I, spans
        return self._gather_all(parts, nq, k, chunk)

    def _gather_all(self, parts, nq: int, k: int, chunk: int):
        """All-gather of the merged per-rank slices, chunk by chunk, back into batch order."""
        G = self.world
        dev = parts[0][0].device
        D = torch.empty((nq, k), dtype=torch.float32, device=dev)
        I = torch.empty((nq, k), dtype=torch.int64, device=dev)
        for (c0, c1, per), (Dm, Im) in zip(chunk_slices(nq, G, chunk), parts):
            Dp = torch.zeros((per, k), dtype=torch.float32, device=dev)
            Ip = torch.zeros((per, k), dtype=torch.int64, device=dev)
            Dp[: Dm.shape[0]] = Dm
            Ip[: Im.shape[0]] = Im
            Dg = torch.empty((G * per, k), dtype=torch.float32, device=dev)
            Ig = torch.empty((G * per, k), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(Dg, Dp, group=self.group)
            dist.all_gather_into_tensor(Ig, Ip, group=self.group)
            D[c0:c1] = Dg[: c1 - c0]
            I[c0:c1] = Ig[: c1 - c0]
        return D, I

    # ------------------------------------------------------------------ end to end from host memory
    def search_host(self, xq_host: np.ndarray, k: int, D_out: np.ndarray, I_out: np.ndarray):
        """The call a user of the sharded index makes with HOST arrays: every rank passes the same
        page-locked fp32 [nq, d] batch and the same-shaped output arrays; each rank copies only ITS
        rows host-to-device (1/G of the batch), the ranks all-gather the query rows over NVLink,
        search, exchange, merge, and each rank copies the results of its own rows back into
        D_out / I_out (rows owned by other ranks are left untouched). Returns the owned spans."""
        assert self._cuda_index, "search_host is the CUDA path"
        nq, d = xq_host.shape
        G, r = self.world, self.rank
        dev = self._dev()
        chunk = self._chunk()
        src = torch.from_numpy(xq_host)
        Dt, It = torch.from_numpy(D_out), torch.from_numpy(I_out)
        pending = None
        spans = []

        def finish(p):
            recv, work, lo, hi, c0 = p
            if work is not None:
                work.wait()
            if self.exchange_mode == "allgather" and G > 1:
                recv = recv[:, lo - c0:hi - c0].contiguous()  # merge (and return) only the owned rows
            Dm, Im = self.codec.merge(recv, self._bases(), self.metric_type)
            if hi > lo:
                Dt[lo:hi].copy_(Dm, non_blocking=True)
                It[lo:hi].copy_(Im, non_blocking=True)
            spans.append((lo, hi))
            return Dm, Im

        keep = []
        for c0, c1, per in chunk_slices(nq, G, chunk):
            lo, hi = min(c1, c0 + r * per), min(c1, c0 + (r + 1) * per)
            mine = torch.zeros((per, d), dtype=torch.float32, device=dev) if hi - lo < per else \
                torch.empty((per, d), dtype=torch.float32, device=dev)
            if hi > lo:
                mine[: hi - lo].copy_(src[lo:hi], non_blocking=True)  # H2D of this rank's rows only
            if G > 1:
                full = torch.empty((G * per, d), dtype=torch.float32, device=dev)
                dist.all_gather_into_tensor(full, mine, group=self.group)
            else:
                full = mine
            D, I = self.search_local(full[: c1 - c0], k)
            P = self.codec.pack(D, I, self.id_base)
            recv, work = self._exchange_start(P, per)
            if pending is not None:
                keep.append(finish(pending[:5]))
            pending = (recv, work, lo, hi, c0, P)
        if pending is not None:
            keep.append(finish(pending[:5]))
        torch.cuda.current_stream().synchronize()
        return spans
