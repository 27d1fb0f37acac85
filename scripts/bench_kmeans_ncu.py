"""Launch-list target for ncu: 3 iterations of the fused trainer on 64,000 x 250 rows, nlist 250."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import newsrecommend_b200.faiss as nf
from newsrecommend_b200 import synth
x = synth.g_skew(64000, 250, 42)
clus = nf.Clustering(250, 250); clus.niter = 3
clus.train(torch.from_numpy(x).cuda(), nf.IndexFlatL2(250))
torch.cuda.synchronize()
print("ok")
