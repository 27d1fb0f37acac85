import os, sys, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import newsrecommend_b200.faiss as nf
from newsrecommend_b200 import _lib, synth
from newsrecommend_b200.parity import compare_topk
from oracle import faiss_oracle as fo
fo.build()
xb, topics = synth.g_skew(364047, 250, 42, return_topics=True)
xq = synth.user_profiles(xb, topics, 6250, 43)
idx = nf.IndexFlatIP(250); idx.add(xb)
q = nf.PackedMatrix.from_tensor(torch.from_numpy(xq).cuda(), planes=idx._query_planes(50))
for _ in range(3): D, I = idx.search_packed(q, 50)
torch.cuda.synchronize()
_lib.profile_enable(True); _lib.profile_read()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): D, I = idx.search_packed(q, 50)
e1.record(); torch.cuda.synchronize()
kms, kn = _lib.profile_read()
Do, Io = fo.knn_fast(xq[:256], xb, 50, 0)
rep = compare_topk(D[:256].cpu().numpy(), I[:256].cpu().numpy(), Do, Io, 0)
print(json.dumps(dict(single=os.environ.get("NRB_NO_SINGLE_CTA") is None, step_ms=e0.elapsed_time(e1) / 10, kernel_ms=kms / kn, ok=rep["ok"])))
