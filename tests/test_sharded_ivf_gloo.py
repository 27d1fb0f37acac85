"""world_size-2 gloo tests (CPU) of the sharded IVF host logic ("Partitioning (IVF)"):
shared coarse quantizer trained from the row-sharded catalog (the faiss subsample assembled from
the shards; data-parallel Lloyd iterations with one fp64 all-reduce each), row-sharding within
lists, and the exchange + merge. The per-rank IVF index, the k-means pieces and the codec are
injected (oracle + numpy stand-ins for the CUDA kernels); tests/sharded_nccl_worker.py covers the
real kernels over NCCL."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_sharded_gloo import NumpyCodec, _free_port


class NumpyKMeansOps:
    """numpy / oracle restatement of GpuKMeansOps (newsrecommend_b200/sharded.py)."""

    def __init__(self, d, k, metric):
        from oracle import faiss_oracle as fo
        self.fo, self.d, self.k, self.metric = fo, d, k, metric

    def device(self):
        return torch.device("cpu")

    def to_device(self, x):
        return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32))

    def set_rows(self, x):
        self.x = x.numpy()
        self.n = self.x.shape[0]

    def assign(self, cent):
        if not self.n:
            return None, torch.zeros((), dtype=torch.float64)
        D, I = self.fo.knn(self.x, cent.numpy(), 1, self.metric)
        return torch.from_numpy(I.reshape(-1)), torch.tensor(float(D.astype(np.float64).sum()), dtype=torch.float64)

    def partial_sums(self, assign, table):
        t = np.zeros((self.k, self.d + 1), dtype=np.float64)
        if self.n:
            a = assign.numpy()
            np.add.at(t[:, : self.d], a, self.x.astype(np.float64))
            np.add.at(t[:, self.d], a, 1.0)
        table[: self.k * (self.d + 1)] = torch.from_numpy(t.reshape(-1))

    def means_and_split(self, table, n_total, spherical):
        t = table[: self.k * (self.d + 1)].numpy().reshape(self.k, self.d + 1)
        hassign = t[:, self.d].astype(np.float32)
        inv = np.where(hassign > 0, np.float32(1.0) / np.maximum(hassign, np.float32(1.0)), np.float32(0.0)).astype(np.float32)
        cent = np.ascontiguousarray(t[:, : self.d].astype(np.float32) * inv[:, None])
        imb = float((hassign.astype(np.float64) ** 2).sum() * self.k / float(n_total) ** 2)
        nsplit = self.fo.split_clusters(cent, hassign, n_total)
        if spherical:
            self.fo.normalize_L2(cent)
        return torch.from_numpy(cent), imb, nsplit


def _worker(rank, world, port, metric, mode, nb, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from newsrecommend_b200.sharded import ShardedIndexIVFFlat, shard_range
        from oracle import faiss_oracle as fo
        rng = np.random.default_rng(1)
        d, nlist, k = 24, 4, 10
        topics = rng.standard_normal((6, d)).astype(np.float32) * 3
        xb = (topics[rng.integers(0, 6, nb)] + rng.standard_normal((nb, d))).astype(np.float32)
        xq = (topics[rng.integers(0, 6, 41)] + rng.standard_normal((41, d))).astype(np.float32)
        idx = ShardedIndexIVFFlat(d, nlist, metric, make_index=fo.IndexFlat, make_ivf=fo.IndexIVFFlat, codec=NumpyCodec,
                                  kmeans_ops=NumpyKMeansOps(d, nlist, metric), chunk_queries=16)
        lo, hi = shard_range(nb, world, rank)
        idx.train_local(xb[lo:hi], lo, nb, mode=mode)
        assert idx.is_trained and idx.quantizer.ntotal == nlist
        cent = np.array(idx.quantizer.xb, copy=True)
        # single-process oracle on the whole catalog
        qo = fo.IndexFlat(d, metric)
        single = fo.IndexIVFFlat(qo, d, nlist, metric)
        single.train(xb)
        if mode == "gather":
            # the subsample assembled from the shards is the one faiss draws: bit-equal centroids
            assert np.array_equal(cent, qo.xb), "gathered training differs from single-index training"
        else:
            # data-parallel: fp64 sums instead of faiss's sequential fp32 -> same clustering up to rounding
            o1, o2 = idx.iteration_stats[-1].obj, single.clustering.iteration_stats[-1].obj
            assert abs(o1 - o2) <= 0.02 * abs(o2), (o1, o2)
            assert len(idx.iteration_stats) == idx.cp.niter
        # every rank holds the same centroids
        t = torch.from_numpy(cent.copy())
        dist.broadcast(t, src=0)
        assert np.array_equal(t.numpy(), cent)
        # teacher-forced search parity: single oracle IVF with THESE centroids
        q2 = fo.IndexFlat(d, metric)
        q2.add(cent)
        ref = fo.IndexIVFFlat(q2, d, nlist, metric)
        ref.train(xb)
        ref.add(xb)
        idx.add_local(xb[lo:hi], lo, nb)
        assert idx.ntotal == nb and idx.local.ntotal == hi - lo
        # row-sharding WITHIN lists: the shards' list sizes add up to the single index's
        sizes = torch.from_numpy(np.asarray(idx.local.list_sizes(), dtype=np.int64))
        dist.all_reduce(sizes)
        assert np.array_equal(sizes.numpy(), ref.list_sizes())
        for nprobe in (1, 3, nlist):
            idx.nprobe = ref.nprobe = nprobe
            Do, Io = ref.search(xq, k)
            D, I = idx.search(xq, k)
            assert np.array_equal(I.numpy(), Io), (nprobe, "merged ids differ from the single IVF index")
            assert np.allclose(D.numpy(), Do, rtol=1e-5, atol=1e-5)
        # the same index through the *_global entry points (every rank passes the whole matrix)
        idx3 = ShardedIndexIVFFlat(d, nlist, metric, make_index=fo.IndexFlat, make_ivf=fo.IndexIVFFlat, codec=NumpyCodec,
                                   kmeans_ops=NumpyKMeansOps(d, nlist, metric), exchange="allgather")
        idx3.train_global(xb, mode=mode)
        idx3.add_global(xb)
        assert np.array_equal(np.asarray(idx3.quantizer.xb), cent)  # deterministic: the same centroids again
        idx3.nprobe = ref.nprobe = 2
        D3, I3 = idx3.search(xq, k)
        Do, Io = ref.search(xq, k)
        assert np.array_equal(I3.numpy(), Io) and np.allclose(D3.numpy(), Do, rtol=1e-5, atol=1e-5)
        out.put((rank, True))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("metric,mode,nb", [
    (1, "gather", 2001), (0, "gather", 2001), (1, "data_parallel", 2001), (0, "data_parallel", 2001),
    (1, "gather", 601),  # 601 < 4 * 256: no subsampling, every row trains
])
def test_sharded_ivf_world2_gloo(metric, mode, nb):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, metric, mode, nb, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    res = sorted(out.get(timeout=5) for _ in range(2))
    assert res == [(0, True), (1, True)], res


def test_data_parallel_kmeans_one_rank_matches_oracle_iteration():
    """One rank, one iteration: the data-parallel trainer's centroids equal the oracle's
    compute_centroids on the same assignment (fp64 vs sequential fp32 sums: 1e-5)."""
    from newsrecommend_b200.sharded import train_kmeans_data_parallel
    from oracle import faiss_oracle as fo
    rng = np.random.default_rng(2)
    n, d, k = 1500, 24, 5
    x = rng.standard_normal((n, d)).astype(np.float32)
    cp = fo.ClusteringParameters()
    cp.niter = 1
    cent, stats = train_kmeans_data_parallel(x, 0, n, k, cp, 1, ops=NumpyKMeansOps(d, k, 1))
    clus = fo.Clustering(d, k, cp)
    q = fo.IndexFlat(d, 1)
    clus.train(x, q)
    assert np.allclose(cent.numpy(), q.xb, rtol=1e-5, atol=1e-6)
    assert abs(stats[0].obj - clus.iteration_stats[0].obj) <= 1e-4 * abs(clus.iteration_stats[0].obj)
    assert stats[0].nsplit == clus.iteration_stats[0].nsplit
