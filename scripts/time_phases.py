"""Diagnostic: device time (CUDA events) and host time of each phase of one flat search step, for
both tcgen05 kernel variants."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import numpy as np
import newsrecommend_b200.faiss as nf
from newsrecommend_b200 import _lib, synth

NB, D, NQ, K = 364_047, 250, 50_000, 50
xb, topics = synth.g_skew(NB, D, 42, return_topics=True)
xq = synth.user_profiles(xb, topics, NQ, 43)
index = nf.IndexFlatIP(D); index.add(xb)
xq_dev = torch.from_numpy(xq).cuda()

def ev(): return torch.cuda.Event(enable_timing=True)

for variant in (1, 2, 1, 2):
    _lib.lib.nrb_set_tc_variant(variant)
    for rep in range(4):
        torch.cuda.synchronize()
        e = [ev() for _ in range(4)]
        h0 = time.perf_counter(); e[0].record()
        q = nf.PackedMatrix.from_tensor(xq_dev, planes=("hi", "lo"))
        h1 = time.perf_counter(); e[1].record()
        _lib.profile_enable(True); _lib.profile_read()
        Dd, Id = index.search_packed(q, K)
        h2 = time.perf_counter(); e[2].record()
        torch.cuda.synchronize(); h3 = time.perf_counter()
        kms, kn = _lib.profile_read(); _lib.profile_enable(False)
        print(f"variant {variant} rep {rep}: pack dev {e[0].elapsed_time(e[1]):.3f} ms host {1e3*(h1-h0):.3f} | "
              f"search dev {e[1].elapsed_time(e[2]):.3f} ms host {1e3*(h2-h1):.3f} | kernel {kms:.3f} ms | total wall {1e3*(h3-h0):.3f}")
# back-to-back loop like bench
for variant in (1, 2):
    _lib.lib.nrb_set_tc_variant(variant)
    torch.cuda.synchronize()
    a, b = ev(), ev(); a.record()
    for _ in range(5):
        q = nf.PackedMatrix.from_tensor(xq_dev, planes=("hi", "lo"))
        Dd, Id = index.search_packed(q, K)
    b.record(); torch.cuda.synchronize()
    print(f"variant {variant}: 5 back-to-back steps {a.elapsed_time(b)/5:.3f} ms/step")
