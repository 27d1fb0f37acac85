"""GPU tests of the catalog-sharded search (north_star item 4): the wire-format kernels against
their numpy restatement, G virtual shards in one process (the exact data flow of the all-to-all
by query range, exchanged by hand) against the single index and the oracle, and the real NCCL
path under torchrun on every GPU of the box."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("metric", [0, 1])
def test_pack_and_merge_packed_kernels(nf, metric):
    import torch
    from newsrecommend_b200.sharded import GpuCodec
    from test_sharded_gloo import NumpyCodec
    rng = np.random.default_rng(0)
    G, nq, k = 8, 300, 50
    Dp = np.sort(rng.standard_normal((G, nq, k)).astype(np.float32), axis=2)
    Dp = Dp[:, :, ::-1].copy() if metric == 0 else np.sort(np.abs(Dp), axis=2)
    Ip = np.stack([rng.permutation(40_000)[: nq * k].reshape(nq, k) + 40_000 * s for s in range(G)]).astype(np.int64)
    Ip[3, :, 40:] = -1  # a shard with fewer than k results
    Dp[0, :, 5] = Dp[1, :, 7]  # exact score ties between shards: the lower shard (lower id) wins
    Dp[0] = -np.sort(-Dp[0], axis=1) if metric == 0 else np.sort(Dp[0], axis=1)
    bases = np.arange(G, dtype=np.int64) * 40_000
    packs, packs_np = [], []
    for s in range(G):
        D, I = torch.from_numpy(Dp[s]).cuda(), torch.from_numpy(Ip[s]).cuda()
        P = GpuCodec.pack(D, I, int(bases[s]))
        Pn = NumpyCodec.pack(torch.from_numpy(Dp[s]), torch.from_numpy(Ip[s]), int(bases[s]))
        assert torch.equal(P.cpu(), Pn)  # bit-exact wire format
        packs.append(P)
        packs_np.append(Pn)
    D, I = GpuCodec.merge(torch.stack(packs), torch.from_numpy(bases).cuda(), metric)
    Dn, In = NumpyCodec.merge(torch.stack(packs_np), torch.from_numpy(bases), metric)
    assert torch.equal(I.cpu(), In) and torch.equal(D.cpu(), Dn)


@pytest.mark.parametrize("metric", [0, 1])
def test_virtual_shards_match_single_index_and_oracle(nf, oracle, metric):
    """The data flow of ShardedIndexFlat.search(exchange="alltoall") with G = 4 shards living in
    one process: per-shard search with global ids -> nrb_pack_topk -> the rows of every query
    range from every shard (what all_to_all_single delivers) -> nrb_merge_topk_packed."""
    import torch
    from newsrecommend_b200.parity import compare_topk
    from newsrecommend_b200.sharded import GpuCodec, chunk_slices, shard_range
    rng = np.random.default_rng(3)
    nb, d, nq, k, G = 50_003, 250, 1_003, 50, 4
    xb = rng.standard_normal((nb, d), dtype=np.float32)
    xb[40_000:40_050] = xb[100:150]  # exact duplicates in different shards: tie order = ascending id
    xq = rng.standard_normal((nq, d), dtype=np.float32)
    xq_dev = torch.from_numpy(xq).cuda()
    single = nf.IndexFlat(d, metric)
    single.add(xb)
    Ds, Is = single.search(xq_dev, k)
    shards, bases = [], []
    for r in range(G):
        lo, hi = shard_range(nb, G, r)
        ix = nf.IndexFlat(d, metric)
        ix.add(xb[lo:hi])
        shards.append(ix)
        bases.append(lo)
    bases_dev = torch.tensor(bases, dtype=torch.int64, device="cuda")
    D = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    for c0, c1, per in chunk_slices(nq, G, 512):
        packed = []
        for r in range(G):
            q = nf.PackedMatrix.from_tensor(xq_dev[c0:c1], planes=shards[r]._query_planes(k))
            Dl, Il = shards[r].search_packed(q, k, bases[r])
            packed.append(GpuCodec.pack(Dl, Il, bases[r]))
        for j in range(G):  # rank j receives its rows of the chunk from every shard
            lo, hi = min(c1 - c0, j * per), min(c1 - c0, (j + 1) * per)
            recv = torch.stack([p[lo:hi] for p in packed])
            Dm, Im = GpuCodec.merge(recv, bases_dev, metric)
            D[c0 + lo:c0 + hi], I[c0 + lo:c0 + hi] = Dm, Im
    Do, Io = oracle.knn_fast(xq, xb, k, metric)
    rep = compare_topk(D.cpu().numpy(), I.cpu().numpy(), Do, Io, metric)
    assert rep["ok"], rep
    # the merge is exact on scores: position-wise identical scores to the single index
    assert torch.equal(D, Ds)
    assert compare_topk(D.cpu().numpy(), I.cpu().numpy(), Ds.cpu().numpy(), Is.cpu().numpy(), metric)["ok"]


def test_sharded_nccl_torchrun():
    """ShardedIndexFlat over NCCL, one rank per visible GPU (2 ranks when the box has more than
    one GPU; world size 1 on a single-GPU box, which still runs the NCCL process group, the
    packed merge and the host-array entry point)."""
    import torch
    n = min(torch.cuda.device_count(), 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "sharded_nccl_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and f"SHARDED_NCCL_OK world={n}" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
