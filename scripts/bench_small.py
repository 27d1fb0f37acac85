"""Low-latency path (csrc/small_batch.cu) measurements, one JSON line:

* the reference's call pattern (Retrieval.py:30-32): N sequential `search(profile[1, 256], 1)`
  calls on a 300-centroid IndexFlatL2 through the shim, numpy in / numpy out -- microseconds per
  call on the B200 small path, on the tcgen05 batch path (small_nq = 0) and on the oracle's
  sequential path (the CPU route faiss itself takes for nq < 20);
* nq = 1 / 16 exact flat search over the 364,047 x 250 catalog: ms per call and achieved GB/s
  (algorithmic bytes = n * kp * 4 per group of <= 16 queries) against the measured HBM peak;
* HBM-regime IVF list scan (nlist 250, nprobe 16): nq = 1, 16, 64 -- ms per call and GB/s over
  the algorithmic bytes sum(|probed list|) * kp * 4, next to the tcgen05 unit scan on the same batch.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import newsrecommend_b200.faiss as nf
from newsrecommend_b200 import _lib, synth
from newsrecommend_b200.parity import compare_topk
from oracle import faiss_oracle as fo

fo.build()
out = {}
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    HBM = 6545.9
out["hbm_peak_gbs"] = HBM

# ---- the reference's call pattern
rng = np.random.default_rng(0)
d, nlist, ncalls = 256, 300, 20_000
cent = rng.standard_normal((nlist, d), dtype=np.float32)
users = rng.standard_normal((ncalls, d), dtype=np.float32)
index = nf.IndexFlatL2(d)
index.add(cent)
oi = fo.IndexFlatL2(d)
oi.add(cent)


def loop(ix, n):
    t0 = time.perf_counter()
    res = [int(ix.search(users[u].reshape(1, d), 1)[1][0, 0]) for u in range(n)]
    return (time.perf_counter() - t0) / n * 1e6, res


loop(index, 200)
us_small, got = loop(index, ncalls)
index.small_nq = 0
loop(index, 50)
us_batch, got_b = loop(index, 2_000)
index.small_nq = nf.SMALL_NQ
us_cpu, want = loop(oi, ncalls)
out["reference_call_pattern"] = dict(
    what="search(profile[1,256], 1) on IndexFlatL2 over 300 centroids, numpy in/out, sequential calls (Retrieval.py:30-32)",
    calls=ncalls, us_per_call_small_path=us_small, us_per_call_tcgen05_batch_path=us_batch,
    us_per_call_oracle_cpu_seq=us_cpu, ids_equal_oracle=bool(got == want), ids_equal_batch_path=bool(got[:2000] == got_b))


def timed_ms(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- flat nq = 1 / 16 over the full catalog
xb, topics = synth.g_skew(synth.N_ARTICLES, 250, 42, return_topics=True)
xq = synth.user_profiles(xb, topics, 256, 43)
flat = nf.IndexFlatIP(250)
flat.add(xb)
xq_dev = torch.from_numpy(xq).cuda()
kp = 256
rows = []
for nq in (1, 4, 16):
    ms = timed_ms(lambda: flat._search_small_dev(xq_dev[:nq], 50))
    flat.small_nq = 0
    ms_tc = timed_ms(lambda: flat.search(xq_dev[:nq], 50), reps=5)
    flat.small_nq = nf.SMALL_NQ
    D, I = flat._search_small_dev(xq_dev[:nq], 50)
    Do, Io = fo.knn(xq[:nq], xb, 50, 0)
    rep = compare_topk(D.cpu().numpy(), I.cpu().numpy(), Do, Io, 0)
    alg = synth.N_ARTICLES * kp * 4
    rows.append(dict(nq=nq, ms_small=ms, ms_tcgen05_path=ms_tc, gbs=alg / (ms / 1e3) / 1e9, frac_hbm=alg / (ms / 1e3) / 1e9 / HBM,
                     parity_ok=rep["ok"]))
out["flat_small_nq"] = rows

# ---- IVF small scan
quant = nf.IndexFlatIP(250)
ivf = nf.IndexIVFFlat(quant, 250, 250, nf.METRIC_INNER_PRODUCT)
ivf.train(xb)
ivf.add(xb)
ivf.nprobe = 16
sizes = ivf.list_sizes()
cent = quant.reconstruct_n(0, 250)
rows = []
for nq in (1, 16, 64):
    q = xq_dev[:nq]
    ms = timed_ms(lambda: ivf.search(q, 50))
    D, I = ivf.search(q, 50)
    coarse = np.argsort(-(xq[:nq] @ cent.T), axis=1, kind="stable")[:, :16]
    scanned = int(sizes[coarse].sum())
    alg = scanned * kp * 4
    old = nf.IVF_SMALL_NQ
    nf.IVF_SMALL_NQ = 0
    ms_tc = timed_ms(lambda: ivf.search(q, 50), reps=5)
    Dt, It = ivf.search(q, 50)
    nf.IVF_SMALL_NQ = old
    rep = compare_topk(D.cpu().numpy(), I.cpu().numpy(), Dt.cpu().numpy(), It.cpu().numpy(), 0)
    rows.append(dict(nq=nq, nprobe=16, ms_small_incl_coarse=ms, ms_tcgen05_unit_scan=ms_tc, scanned_rows=scanned,
                     alg_mb=alg / 1e6, gbs=alg / (ms / 1e3) / 1e9, frac_hbm=alg / (ms / 1e3) / 1e9 / HBM,
                     same_as_unit_scan=rep["ok"]))
out["ivf_small_scan"] = rows
print(json.dumps(out))
