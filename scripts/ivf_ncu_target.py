"""ncu launch-list target: one IndexIVFFlat.search of 250,000 queries (nlist 250, nprobe 16, k 50)
after training/adding; a marker kernel (torch fill of 7 elements) brackets the measured search."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import newsrecommend_b200.faiss as nf
from newsrecommend_b200 import synth
xb, topics = synth.g_skew(synth.N_ARTICLES, 250, 42, return_topics=True)
xq = synth.user_profiles(xb, topics, int(os.environ.get("NRB_IVF_NQ", "250000")), 44)
xb_d, xq_d = torch.from_numpy(xb).cuda(), torch.from_numpy(xq).cuda()
quant = nf.IndexFlatIP(250)
ivf = nf.IndexIVFFlat(quant, 250, 250, nf.METRIC_INNER_PRODUCT)
ivf.train(xb_d); ivf.add(xb_d); ivf.nprobe = 16
if os.environ.get("NRB_WARM", "1") == "1":
    ivf.search(xq_d, 50)
torch.cuda.synchronize()
marker = torch.empty(7, device="cuda"); marker.fill_(1.0)
D, I = ivf.search(xq_d, 50)
marker.fill_(2.0)
torch.cuda.synchronize()
print("ok")
