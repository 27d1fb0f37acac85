"""Shape sweep of the exact flat search on one B200 (device-resident, auto path): looks for cliffs away
from the benchmark point -- metric, k, d, batch size, catalog size. One JSON line per case."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import newsrecommend_b200.faiss as nf
from newsrecommend_b200 import _lib, synth

CASES = [  # (nb, d, nq, k, metric)
    (364_047, 250, 50_000, 50, 0), (364_047, 250, 50_000, 50, 1),
    (364_047, 250, 50_000, 1, 0), (364_047, 250, 50_000, 10, 0), (364_047, 250, 50_000, 100, 0), (364_047, 250, 50_000, 128, 0),
    (364_047, 64, 50_000, 50, 0), (364_047, 128, 50_000, 50, 0), (364_047, 256, 50_000, 50, 0), (364_047, 300, 50_000, 50, 0),
    (364_047, 250, 1_000, 50, 0), (364_047, 250, 5_000, 50, 0), (364_047, 250, 18_944, 50, 0), (364_047, 250, 250_000, 50, 0),
    (20_000, 250, 50_000, 50, 0), (2_000_000, 250, 50_000, 50, 0),
]
if os.environ.get("NRB_SWEEP_CASES"):  # e.g. "0,1" = only the first two cases
    CASES = [CASES[int(i)] for i in os.environ["NRB_SWEEP_CASES"].split(",")]
for nb, d, nq, k, metric in CASES:
    xb, topics = synth.g_skew(nb, d, 42, return_topics=True)
    xq = synth.user_profiles(xb, topics, nq, 43)
    index = nf.IndexFlat(d, metric)
    index.add(torch.from_numpy(xb).cuda())
    del xb
    xq_d = torch.from_numpy(xq).cuda()
    planes = index._query_planes(k)

    def step():
        q = nf.PackedMatrix.from_tensor(xq_d, planes=planes)
        return index.search_packed(q, k)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    f0 = int(_lib.lib.nrb_fallback_query_count())
    _lib.profile_enable(True)
    _lib.profile_read()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    kms, kn = _lib.profile_read()
    _lib.profile_enable(False)
    fb = (int(_lib.lib.nrb_fallback_query_count()) - f0) // reps
    flop = 2.0 * nq * nb * d
    print(json.dumps(dict(nb=nb, d=d, nq=nq, k=k, metric="l2" if metric else "ip", ms=round(dt * 1e3, 3),
                          k2_kernel_ms=round(kms / reps, 3), launches=kn / reps, qps=round(nq / dt),
                          alg_tflops=round(flop / dt / 1e12, 1), fallback_queries=fb)), flush=True)
    del index, xq_d
    torch.cuda.empty_cache()
