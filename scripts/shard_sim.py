import json, os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import newsrecommend_b200.faiss as nf
from newsrecommend_b200 import _lib, synth
xb, topics = synth.g_skew(364047, 250, 42, return_topics=True)
xq = synth.user_profiles(xb, topics, 50000, 43)
for nb, nq in ((364047, 50000), (182024, 12500), (182024, 25000), (45506, 50000)):
    index = nf.IndexFlatIP(250); index.add(torch.from_numpy(xb[:nb]).cuda())
    xq_d = torch.from_numpy(xq[:nq]).cuda(); planes = index._query_planes(50)
    def step():
        q = nf.PackedMatrix.from_tensor(xq_d, planes=planes); return index.search_packed(q, 50)
    for _ in range(3): step()
    torch.cuda.synchronize(); _lib.profile_enable(True); _lib.profile_read()
    for _ in range(10): step()
    torch.cuda.synchronize(); kms, kn = _lib.profile_read(); _lib.profile_enable(False)
    print(os.environ.get('NRB_LIB','default').split('/')[-1], nb, nq, 'kernel_ms %.3f' % (kms / 10))
