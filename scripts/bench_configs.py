"""BASELINE configs[3] (item-item cosine neighbours, 364,047^2, top-20) and configs[4] (10M x 256
catalog, top-100; here on ONE GPU with a 131,072-query slice of the 1M batch) -- timings plus
parity samples against the oracle. One JSON line per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import newsrecommend_b200.faiss as nf
from newsrecommend_b200 import _lib, synth
from newsrecommend_b200.parity import compare_topk
from oracle import faiss_oracle as fo

which = sys.argv[1:] or ["4", "5"]

def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    return out, time.perf_counter() - t0

if "4" in which:
    x = synth.g_skew(synth.N_ARTICLES, 250, 42)
    nf.normalize_L2(x)
    xd = torch.from_numpy(x).cuda()
    index = nf.IndexFlatIP(250); index.add(xd)
    index.search(xd[:8192], 20)
    _lib.profile_enable(True); _lib.profile_read()
    (D, I), t = timed(lambda: index.search(xd, 20))
    kms, kn = _lib.profile_read(); _lib.profile_enable(False)
    Dh, Ih = D.cpu().numpy(), I.cpu().numpy()
    ns = 512
    Do, Io = fo.knn_fast(x[:ns], x, 20, 0)
    rep = compare_topk(Dh[:ns], Ih[:ns], Do, Io, 0)
    n = x.shape[0]
    print(json.dumps(dict(config="item-item cosine 364,047 x 364,047 x 250, top-20", seconds=t, qps=n / t, kernel_ms=kms,
                          alg_tflops=2.0 * n * n * 250 / (kms / 1e3) / 1e12 if kms else None,
                          self_match_rank0=float((Ih[:, 0] == np.arange(n)).mean()), top_score_min=float(Dh[:, 0].min()),
                          fallback_queries=int(_lib.lib.nrb_fallback_query_count()),
                          parity_sample=dict(queries=ns, ok=rep["ok"], exact_ordered=rep["exact_ordered"], recall=rep["recall"],
                                             tie_exempt=rep["tie_exempt_queries"], max_rel_score_err=rep["max_rel_score_err"]))))
    del index, xd, D, I
    torch.cuda.empty_cache()

if "5" in which:
    NB, D_, NQ, K, CH = 10_000_000, 256, 131_072, 100, 1_000_000
    rng = np.random.default_rng(45)
    G, r = 4096, 16
    w = 1.0 / np.arange(1, G + 1) ** 0.8; w /= w.sum()
    centers = rng.standard_normal((G, r)).astype(np.float32)
    q_, _ = np.linalg.qr(rng.standard_normal((D_, r))); W = q_.T.astype(np.float32)
    def chunk(nrows, rg):
        comp = rg.choice(G, size=nrows, p=w)
        z = centers[comp] + 0.7 * rg.standard_normal((nrows, r), dtype=np.float32)
        return np.ascontiguousarray(z @ W + 0.02 * rg.standard_normal((nrows, D_), dtype=np.float32), dtype=np.float32)
    xq = chunk(NQ, np.random.default_rng(46))
    ns = 64
    index = nf.IndexFlatIP(D_)
    cand_D, cand_I = [], []
    t_add = 0.0
    for c in range(NB // CH):
        xc = chunk(CH, rng)
        _, dt = timed(lambda: index.add(xc)); t_add += dt
        Dc, Ic = fo.knn_fast(xq[:ns], xc, K, 0)
        cand_D.append(Dc); cand_I.append(Ic + c * CH)
    Dall, Iall = np.concatenate(cand_D, 1), np.concatenate(cand_I, 1)
    o = np.argsort(-Dall, axis=1, kind="stable")[:, :K]
    Do, Io = np.take_along_axis(Dall, o, 1), np.take_along_axis(Iall, o, 1)
    xqd = torch.from_numpy(xq).cuda()
    index.search(xqd[:4096], K)
    _lib.profile_enable(True); _lib.profile_read()
    f0 = int(_lib.lib.nrb_fallback_query_count())
    (D, I), t = timed(lambda: index.search(xqd, K))
    kms, kn = _lib.profile_read(); _lib.profile_enable(False)
    rep = compare_topk(D[:ns].cpu().numpy(), I[:ns].cpu().numpy(), Do, Io, 0)
    print(json.dumps(dict(config="10M x 256 catalog, 131,072 of the 1M queries, top-100, one B200", ntotal=index.ntotal,
                          add_seconds=t_add, seconds=t, qps=NQ / t, kernel_ms=kms,
                          alg_tflops=2.0 * NQ * NB * D_ / (kms / 1e3) / 1e12 if kms else None,
                          hbm_gb=torch.cuda.max_memory_allocated() / 1e9,
                          fallback_queries=int(_lib.lib.nrb_fallback_query_count()) - f0,
                          parity_sample=dict(queries=ns, ok=rep["ok"], exact_ordered=rep["exact_ordered"], recall=rep["recall"],
                                             tie_exempt=rep["tie_exempt_queries"], max_rel_score_err=rep["max_rel_score_err"]))))
