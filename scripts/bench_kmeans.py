"""k-means trainer measurements (one JSON line): nrb_kmeans_train (one C-ABI call, no host sync
inside) for nlist 250 / d 250 at niter 10 (IndexIVFFlat default) and 80 (Retrieval.py:13) on the
config-1 catalog, next to the oracle port on the host cores; and the K1b update kernels alone
(64,000 x 256 and 364,047 x 256 rows) as achieved GB/s against the measured HBM peak."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import newsrecommend_b200.faiss as nf
from newsrecommend_b200 import synth
from oracle import faiss_oracle as fo

fo.build()
try:
    from threadpoolctl import threadpool_limits
    threadpool_limits(limits=len(os.sched_getaffinity(0)))
except Exception:  # noqa: BLE001
    pass
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    HBM = 6545.9
out = {"hbm_peak_gbs": HBM}
x = synth.g_skew(synth.N_ARTICLES, 250, 42)
xd = torch.from_numpy(x).cuda()
rows = []
for metric, name in ((1, "L2"), (0, "IP")):
    for niter in (10, 80):
        def train():
            clus = nf.Clustering(250, 250)
            clus.niter = niter
            index = nf.IndexFlatL2(250) if metric else nf.IndexFlatIP(250)
            clus.train(xd, index)
            return clus
        train()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        clus = train()
        torch.cuda.synchronize(); t_gpu = time.perf_counter() - t0
        row = dict(metric=name, nlist=250, niter=niter, rows=64000, gpu_seconds=t_gpu, gpu_ms_per_iteration=t_gpu / niter * 1e3,
                   obj_first=clus.iteration_stats[0].obj, obj_last=clus.iteration_stats[-1].obj,
                   imbalance_last=clus.iteration_stats[-1].imbalance_factor)
        if niter == 10 or metric == 1:
            oc = fo.Clustering(250, 250); oc.niter = niter
            oi = fo.IndexFlatL2(250) if metric else fo.IndexFlatIP(250)
            t0 = time.perf_counter(); oc.train(x, oi); t_cpu = time.perf_counter() - t0
            row.update(cpu_oracle_seconds=t_cpu, cpu_threads=len(os.sched_getaffinity(0)), speedup=t_cpu / t_gpu,
                       cpu_obj_last=oc.iteration_stats[-1].obj)
        rows.append(row)
out["train"] = rows

# K1b alone
upd = []
for n in (64000, synth.N_ARTICLES):
    p = nf.PackedMatrix.from_tensor(xd[:n], planes=("raw",))
    index = nf.IndexFlatL2(250); index.add(clus.centroids.reshape(250, 250))
    a = index.search(xd[:n], 1)[1].reshape(-1).contiguous()
    for _ in range(3):
        nf.kmeans_update(p, a, 250)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        nf.kmeans_update(p, a, 250)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    alg = n * 256 * 4
    sizes = np.bincount(a.cpu().numpy(), minlength=250)
    upd.append(dict(rows=n, ms_update_all_launches=ms, alg_mb=alg / 1e6, gbs=alg / (ms / 1e3) / 1e9, frac_hbm=alg / (ms / 1e3) / 1e9 / HBM,
                    list_min=int(sizes.min()), list_max=int(sizes.max()),
                    note="counting sort (3 launches) + plan + partial + finish, event-timed back to back"))
out["k1b_update"] = upd
print(json.dumps(out))
