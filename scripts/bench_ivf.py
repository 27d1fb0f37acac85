"""IVF-Flat (BASELINE configs[1]/[2]): k-means train + add + search timings on the B200 and a
teacher-forced parity sample against the oracle. Prints one JSON line per nlist."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import newsrecommend_b200.faiss as nf
from newsrecommend_b200 import _lib, synth
from newsrecommend_b200.parity import compare_topk, recall_at_k
from oracle import faiss_oracle as fo

NLISTS = [int(a) for a in sys.argv[1:]] or [250]
if os.environ.get("NRB_IVF_BATCH"):
    nf.IVF_QUERY_BATCH = int(os.environ["NRB_IVF_BATCH"])  # queries per nrb_ivf_search call
NQ = int(os.environ.get("NRB_IVF_NQ", "250000"))
L2 = os.environ.get("NRB_IVF_METRIC", "ip") == "l2"
MET = 1 if L2 else 0
xb, topics = synth.g_skew(synth.N_ARTICLES, 250, 42, return_topics=True)
xq = synth.user_profiles(xb, topics, NQ, 44)
xb_d, xq_d = torch.from_numpy(xb).cuda(), torch.from_numpy(xq).cuda()

def timed(fn, reps=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize(); return out, (time.perf_counter() - t0) / reps

flat = (nf.IndexFlatL2 if L2 else nf.IndexFlatIP)(250); flat.add(xb_d)
(_, I_exact), t_flat = timed(lambda: flat.search(xq_d[:20000], 50))
for nlist in NLISTS:
    quant = (nf.IndexFlatL2 if L2 else nf.IndexFlatIP)(250)
    ivf = nf.IndexIVFFlat(quant, 250, nlist, MET)
    _, t_train = timed(lambda: ivf.train(xb_d))
    _, t_add = timed(lambda: ivf.add(xb_d))
    sizes = ivf.list_sizes()
    ivf.nprobe = 16
    ivf.search(xq_d, 50)  # warm-up at full size (lists build, allocator growth for the workspaces)
    _lib.profile_enable(True); _lib.profile_read()
    (D, I), t_search = timed(lambda: ivf.search(xq_d, 50))
    kms, kn = _lib.profile_read(); _lib.profile_enable(False)
    rec = recall_at_k(I[:20000].cpu().numpy(), I_exact.cpu().numpy())
    # rows actually scanned: sum over (query, probed list) of the list length
    qp = nf.PackedMatrix.from_tensor(xq_d, planes=quant._query_planes(16))
    _, coarse = quant.search_packed(qp, 16)
    scanned = int(torch.from_numpy(sizes).cuda()[coarse.reshape(-1)].sum().item())
    del qp
    # teacher-forced parity sample: oracle IVF with the same centroids and list contents
    cent = quant.reconstruct_n()
    qo = (fo.IndexFlatL2 if L2 else fo.IndexFlatIP)(250); qo.add(cent)
    ivf_o = fo.IndexIVFFlat(qo, 250, nlist, MET); ivf_o.train(xb); ivf_o.add(xb); ivf_o.nprobe = 16
    ns = 2048
    t0 = time.perf_counter(); Do, Io = ivf_o.search(xq[:ns], 50); t_cpu = time.perf_counter() - t0
    rep = compare_topk(D[:ns].cpu().numpy(), I[:ns].cpu().numpy(), Do, Io, MET)
    same_lists = bool(np.array_equal(sizes, ivf_o.list_sizes()))
    print(json.dumps(dict(metric="l2" if L2 else "ip", nlist=nlist, nq=NQ, nprobe=16, k=50, train_s=t_train, niter=ivf.cp.niter, add_s=t_add,
                          search_s=t_search, search_qps=NQ / t_search, scan_kernel_ms=kms, scan_kernel_launches=kn,
                          scanned_rows=scanned, scan_alg_tflop=2.0 * scanned * 250 / 1e12,
                          scanned_frac_of_flat=scanned / (NQ * float(synth.N_ARTICLES)),
                          list_min=int(sizes.min()), list_max=int(sizes.max()),
                          imbalance=float((sizes.astype(np.float64) ** 2).sum() * nlist / sizes.sum() ** 2),
                          recall_vs_exact=rec, oracle_same_lists=same_lists,
                          parity_sample=dict(queries=ns, id_mismatch_queries=rep["id_mismatch_queries"],
                                             tie_exempt=rep["tie_exempt_queries"], score_violations=rep["score_violations"],
                                             exact_ordered=rep["exact_ordered"], recall=rep["recall"]),
                          cpu_oracle_qps=ns / t_cpu, flat_exact_qps=20000 / t_flat)))
