"""GPU tests of the low-latency path for tiny batches (csrc/small_batch.cu): faiss's own nq < 20
route (exact fp32 distances, no GEMM; Retrieval.py:30-32 issues 50,000 such calls with nq = 1)
and the HBM-regime inverted-list scan, against the oracle's sequential path (oracle.knn with
nq < 20 -> fo_search_seq) and the oracle's IndexIVFFlat."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("nb", [300, 5000, 8192, 8193, 70_001])  # fused kernel up to 8,192 rows, two launches above
def test_small_flat_matches_oracle_seq(nf, oracle, metric, nb):
    import torch
    from newsrecommend_b200.parity import compare_topk
    rng = np.random.default_rng(nb + metric)
    d = 250
    xb = rng.standard_normal((nb, d), dtype=np.float32)
    index = nf.IndexFlat(d, metric)
    index.add(xb)
    for nq, k in [(1, 1), (1, 50), (3, 10), (4, 128), (5, 7), (16, 50), (19, 100)]:
        xq = rng.standard_normal((nq, d), dtype=np.float32)
        Do, Io = oracle.knn(xq, xb, k, metric)  # nq < 20: the oracle's sequential path
        D, I = index.search(xq, k)  # numpy in -> nrb_search_small_host
        assert isinstance(D, np.ndarray) and D.shape == (nq, k) and I.dtype == np.int64
        rep = compare_topk(D, I, Do, Io, metric)
        assert rep["ok"], (nq, k, rep)
        Dd, Id = index.search(torch.from_numpy(xq).cuda(), k)  # CUDA tensor in -> nrb_search_small
        assert np.array_equal(Id.cpu().numpy(), I) and np.array_equal(Dd.cpu().numpy(), D)
        # the same batch through the tcgen05 path: same ids outside ties
        index.small_nq = 0
        Dt, It = index.search(xq, k)
        index.small_nq = nf.SMALL_NQ
        assert compare_topk(D, I, Dt, It, metric)["ok"]


def test_small_flat_is_the_reference_call_pattern(nf, oracle):
    """Retrieval.py:25-32: IndexFlatL2 over 300 centroids, one search(profile[1, d], 1) per user."""
    rng = np.random.default_rng(7)
    d, nlist = 256, 300
    cent = rng.standard_normal((nlist, d), dtype=np.float32)
    users = rng.standard_normal((500, d), dtype=np.float32)
    index = nf.IndexFlatL2(d)
    index.add(cent)
    oi = oracle.IndexFlatL2(d)
    oi.add(cent)
    from newsrecommend_b200 import _lib
    n0 = _lib.launch_count()
    got = np.array([index.search(users[u].reshape(1, d), 1)[1][0, 0] for u in range(500)])
    assert _lib.launch_count() - n0 == 500  # ONE kernel launch per call
    want = np.array([oi.search(users[u].reshape(1, d), 1)[1][0, 0] for u in range(500)])
    assert np.array_equal(got, want)
    D1, _ = index.search(users[:1], 1)
    assert abs(D1[0, 0] - ((users[0] - cent[got[0]]) ** 2).sum()) <= 1e-4 * D1[0, 0]


@pytest.mark.parametrize("metric", [0, 1])
def test_small_flat_edges(nf, oracle, metric):
    """k > ntotal (padding), exact duplicates (ties resolve to the lowest id), heavy ties at the
    k-th value with more tied rows than slots, d that is not a multiple of 4."""
    rng = np.random.default_rng(11)
    d = 37
    xb = rng.standard_normal((40, d), dtype=np.float32)
    index = nf.IndexFlat(d, metric)
    index.add(xb)
    xq = rng.standard_normal((2, d), dtype=np.float32)
    D, I = index.search(xq, 64)
    Do, Io = oracle.knn(xq, xb, 64, metric)
    assert np.array_equal(I, Io) and (I[:, 40:] == -1).all()
    assert np.allclose(D[:, :40], Do[:, :40], rtol=1e-5, atol=1e-6)
    assert (D[:, 40:] == (np.float32(3.4028235e38) if metric else -np.float32(3.4028235e38))).all()
    # heavy ties: 20,000 copies of 3 distinct rows; the answer is the lowest ids of the best row(s)
    base = rng.standard_normal((3, d), dtype=np.float32)
    xb2 = np.tile(base, (20_000 // 3 + 1, 1))[:20_000]
    index2 = nf.IndexFlat(d, metric)
    index2.add(xb2)
    q = rng.standard_normal((1, d), dtype=np.float32)
    D2, I2 = index2.search(q, 50)
    s = base @ q[0] if metric == 0 else -((base - q[0]) ** 2).sum(1)
    best = int(np.argmax(s))
    assert np.array_equal(I2[0], best + 3 * np.arange(50))
    assert np.all(D2[0] == D2[0, 0])


@pytest.mark.parametrize("metric", [0, 1])
def test_ivf_small_scan_matches_oracle_ivf(nf, oracle, metric):
    """Small batches go through nrb_ivf_scan_small (warp per row over the probed lists): same
    centroids (teacher-forced from our trained quantizer) -> the oracle's IndexIVFFlat must give
    the same ids; nprobe = nlist equals the exact flat search."""
    from newsrecommend_b200 import synth
    from newsrecommend_b200.parity import compare_topk
    d, nlist, k = 250, 64, 50
    xb, topics = synth.g_skew(30_000, d, 3, return_topics=True)
    xq = synth.user_profiles(xb, topics, 40, 4)
    quant = nf.IndexFlatL2(d) if metric else nf.IndexFlatIP(d)
    ivf = nf.IndexIVFFlat(quant, d, nlist, metric)
    ivf.train(xb)
    ivf.add(xb)
    cent = quant.reconstruct_n(0, nlist)
    oq = oracle.IndexFlatL2(d) if metric else oracle.IndexFlatIP(d)
    oq.add(cent)
    oivf = oracle.IndexIVFFlat(oq, d, nlist, metric)
    oivf.is_trained = True
    oivf.add(xb)
    for nq, nprobe in [(1, 1), (1, 8), (7, 16), (40, 4), (3, nlist)]:
        ivf.nprobe = oivf.nprobe = nprobe
        D, I = ivf.search(xq[:nq], k)
        Do, Io = oivf.search(xq[:nq], k)
        rep = compare_topk(D, I, Do, Io, metric)
        assert rep["ok"], (nq, nprobe, rep)
    ivf.nprobe = nlist
    D, I = ivf.search(xq[:3], k)
    Df, If = oracle.knn(xq[:3], xb, k, metric)
    assert compare_topk(D, I, Df, If, metric)["ok"]
