"""The reference's own workload: Retrieval.py:11-34 at its real sizes (364,047 articles x 256-d
learned embeddings, 300 clusters, 80 k-means iterations, 50,000 test users), run as the batched
stage on the B200 and as the oracle port on the host cores. One JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from newsrecommend_b200 import pipeline, synth
from oracle import faiss_oracle as fo

n, d, nu, k, niter = synth.N_ARTICLES, 256, synth.N_TEST_USERS, 300, 80
x, topics = synth.g_skew(n, d, 42, return_topics=True)
users = synth.user_profiles(x, topics, nu, 43)
ids = np.arange(n, dtype=np.int64)
xd, ud = torch.from_numpy(x).cuda(), torch.from_numpy(users).cuda()
pipeline.retrieve_candidates(ids, xd, ud, num_clusters=k, niter=2)  # warm-up
torch.cuda.synchronize(); t0 = time.perf_counter()
out = pipeline.retrieve_candidates(ids, xd, ud, num_clusters=k, niter=niter)
torch.cuda.synchronize(); t_gpu = time.perf_counter() - t0
sizes = out["list_sizes"].cpu().numpy()
# CPU: the same stage with the oracle (exact L2 assigner, like the product)
try:
    import ctypes
    from threadpoolctl import threadpool_limits
    threadpool_limits(limits=len(os.sched_getaffinity(0)))
except Exception:
    pass
t0 = time.perf_counter()
clus = fo.Clustering(d, k); clus.niter = niter
index = fo.IndexFlatL2(d)
clus.train(x, index)
cent = clus.centroids.reshape(k, d)
_, a = index.search(x, 1); a = a.reshape(-1)
order = np.argsort(a, kind="stable")
ci = fo.IndexFlatL2(d); ci.add(cent)
_, I = ci.search(users, 1)
t_cpu = time.perf_counter() - t0
print(json.dumps(dict(stage="Retrieval.py:11-34 (k-means 300 x 80 it, assign 364,047, lists, 50,000 users)",
                      gpu_seconds=t_gpu, cpu_oracle_seconds=t_cpu, cpu_threads=len(os.sched_getaffinity(0)),
                      speedup=t_cpu / t_gpu, list_min=int(sizes.min()), list_max=int(sizes.max()),
                      candidates_total=int(out["offsets"][-1]),
                      gpu_obj_last=None, cpu_obj_last=float(clus.iteration_stats[-1].obj))))
