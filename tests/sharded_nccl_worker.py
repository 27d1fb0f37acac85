"""Worker of tests/test_gpu_sharded.py::test_sharded_nccl_torchrun (launched with torchrun, one
rank per GPU, NCCL): the catalog-sharded search must equal the single-GPU index and the oracle
on the same inputs -- both exchange forms, gathered and per-rank results, and the host-array
entry point; then the sharded IVF index (shared quantizer, both training modes). Prints
SHARDED_NCCL_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import newsrecommend_b200.faiss as nf
    from newsrecommend_b200.parity import compare_topk
    from newsrecommend_b200.sharded import ShardedIndexFlat, owned_rows
    from oracle import faiss_oracle as fo
    fo.build()
    rng = np.random.default_rng(5)
    nb, d, nq, k = 60_001, 250, 3_001, 50
    xb = rng.standard_normal((nb, d), dtype=np.float32)
    xq = rng.standard_normal((nq, d), dtype=np.float32)
    xq_dev = torch.from_numpy(xq).cuda()
    for metric in (0, 1):
        Do, Io = fo.knn_fast(xq, xb, k, metric)
        single = nf.IndexFlat(d, metric)
        single.add(xb)
        Ds, Is = single.search(xq_dev, k)
        for exchange in ("alltoall", "allgather"):
            for chunk in (None, 1024):  # one chunk / three chunks (exchange of chunk i overlaps the search of i+1)
                idx = ShardedIndexFlat(d, metric, exchange=exchange, chunk_queries=chunk)
                idx.add_global(xb)
                D, I = idx.search(xq_dev, k)
                rep = compare_topk(D.cpu().numpy(), I.cpu().numpy(), Do, Io, metric)
                assert rep["ok"], (metric, exchange, chunk, rep)
                # merged scores are the per-shard exact fp32 scores: identical to the single index
                assert torch.equal(I, Is) or compare_topk(D.cpu().numpy(), I.cpu().numpy(), Ds.cpu().numpy(),
                                                           Is.cpu().numpy(), metric)["ok"]
                Dr, Ir, spans = idx.search(xq_dev, k, gather=False)
                rows = np.concatenate([np.arange(a, b) for a, b in spans])
                assert spans == owned_rows(nq, world, rank, idx._chunk())
                assert torch.equal(Ir, I[torch.from_numpy(rows).cuda()])
                # host arrays in / out: each rank moves only its own rows over PCIe
                xq_pin = torch.from_numpy(xq).pin_memory().numpy()
                D_pin = torch.full((nq, k), float("nan"), dtype=torch.float32).pin_memory().numpy()
                I_pin = torch.full((nq, k), -7, dtype=torch.int64).pin_memory().numpy()
                spans2 = idx.search_host(xq_pin, k, D_pin, I_pin)
                assert spans2 == spans
                assert np.array_equal(I_pin[rows], I.cpu().numpy()[rows])
                assert np.array_equal(D_pin[rows], D.cpu().numpy()[rows])
                other = np.setdiff1d(np.arange(nq), rows)
                assert (I_pin[other] == -7).all()  # rows owned by other ranks are not touched
    ivf_checks(rank, world, nf, fo, compare_topk)
    dist.barrier()
    if rank == 0:
        print("SHARDED_NCCL_OK world=%d" % world)
    dist.destroy_process_group()


def ivf_checks(rank, world, nf, fo, compare_topk):
    """ShardedIndexIVFFlat (shared quantizer, row-sharding within lists) against the single-GPU
    IndexIVFFlat and the oracle's IVF with the same centroids; both training modes."""
    from newsrecommend_b200 import synth
    from newsrecommend_b200.sharded import ShardedIndexIVFFlat, shard_range
    nb, d, nq, k, nlist = 40_003, 250, 2_001, 50, 32
    xb, topics = synth.g_skew(nb, d, 7, return_topics=True)
    xq = synth.user_profiles(xb, topics, nq, 8)
    xb_dev, xq_dev = torch.from_numpy(xb).cuda(), torch.from_numpy(xq).cuda()
    lo, hi = shard_range(nb, world, rank)
    for metric in (0, 1):
        single = nf.IndexIVFFlat(nf.IndexFlat(d, metric), d, nlist, metric)
        single.train(xb_dev)
        single.add(xb_dev)
        for mode in ("gather", "data_parallel"):
            idx = ShardedIndexIVFFlat(d, nlist, metric, chunk_queries=1024)
            idx.train_local(xb_dev[lo:hi], lo, nb, mode=mode)
            cent = idx.quantizer.reconstruct_n()
            if mode == "gather":  # same subsample, same deterministic trainer: bit-equal centroids
                assert np.array_equal(cent, single.quantizer.reconstruct_n()), (metric, "gather centroids differ")
            else:
                o1, o2 = idx.iteration_stats[-1].obj, single.clustering.iteration_stats[-1].obj
                assert abs(o1 - o2) <= 0.02 * abs(o2), (metric, o1, o2)
            t = torch.from_numpy(cent).cuda()
            dist.broadcast(t, src=0)
            assert np.array_equal(t.cpu().numpy(), cent), "ranks hold different centroids"
            idx.add_local(xb_dev[lo:hi], lo, nb)
            sizes = torch.from_numpy(idx.local.list_sizes().astype(np.int64)).cuda()
            dist.all_reduce(sizes)
            # oracle IVF with THESE centroids (teacher-forced)
            qo = fo.IndexFlat(d, metric)
            qo.add(cent)
            ref = fo.IndexIVFFlat(qo, d, nlist, metric)
            ref.train(xb)
            ref.add(xb)
            assert np.array_equal(sizes.cpu().numpy(), ref.list_sizes()), "list sizes do not add up"
            for nprobe in (1, 8):
                idx.nprobe = ref.nprobe = single.nprobe = nprobe
                D, I = idx.search(xq_dev, k)
                Do, Io = ref.search(xq, k)
                rep = compare_topk(D.cpu().numpy(), I.cpu().numpy(), Do, Io, metric)
                assert rep["ok"], (metric, mode, nprobe, rep)
                if mode == "gather":
                    Ds, Is = single.search(xq_dev, k)
                    assert compare_topk(D.cpu().numpy(), I.cpu().numpy(), Ds.cpu().numpy(), Is.cpu().numpy(), metric)["ok"]
                Dr, Ir, spans = idx.search(xq_dev, k, gather=False)
                rows = np.concatenate([np.arange(a, b) for a, b in spans])
                assert torch.equal(Ir, I[torch.from_numpy(rows).cuda()])


if __name__ == "__main__":
    main()
