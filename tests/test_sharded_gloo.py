"""world_size-2 gloo test (CPU) of the catalog-sharding host logic: row ranges, global ids,
all-gather layout and the merge contract. The per-rank index and the merge are injected (the
oracle index and a numpy merge stand in for the CUDA kernels, which need a GPU); the GPU suite
covers the real kernels (tests/test_gpu_misc.py::test_merge_topk, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _np_merge(Dp, Ip, metric):
    Dp, Ip = Dp.numpy(), Ip.numpy()
    G, nq, k = Dp.shape
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    for q in range(nq):
        d, i = Dp[:, q].reshape(-1), Ip[:, q].reshape(-1)
        keep = i >= 0
        d, i = d[keep], i[keep]
        o = np.argsort(-d if metric == 0 else d, kind="stable")[:k]
        D[q, :len(o)], I[q, :len(o)] = d[o], i[o]
        D[q, len(o):], I[q, len(o):] = (np.float32(3.4028235e38) if metric else -np.float32(3.4028235e38)), -1
    return torch.from_numpy(D), torch.from_numpy(I)


def _worker(rank, world, port, metric, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from newsrecommend_b200.sharded import ShardedIndexFlat, shard_range
        from oracle import faiss_oracle as fo
        rng = np.random.default_rng(0)
        xb = rng.standard_normal((1001, 24), dtype=np.float32)  # odd size: uneven shards
        xq = rng.standard_normal((37, 24), dtype=np.float32)
        idx = ShardedIndexFlat(24, metric, make_index=fo.IndexFlat, merge=_np_merge)
        idx.add_global(xb)
        lo, hi = shard_range(1001, world, rank)
        assert idx.id_base == lo and idx.local.ntotal == hi - lo and idx.ntotal == 1001
        D, I = idx.search(xq, 10)
        Do, Io = fo.knn(xq, xb, 10, metric)
        assert np.array_equal(I.numpy(), Io), "merged ids differ from the single-index answer"
        assert np.allclose(D.numpy(), Do, rtol=1e-5, atol=1e-5)
        out.put((rank, True))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("metric", [0, 1])
def test_sharded_search_world2_gloo(metric):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, metric, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    res = sorted(out.get(timeout=5) for _ in range(2))
    assert res == [(0, True), (1, True)], res


def test_shard_range_covers_catalog():
    from newsrecommend_b200.sharded import shard_range
    for nb in (0, 1, 7, 364047):
        for world in (1, 2, 3, 8):
            spans = [shard_range(nb, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
