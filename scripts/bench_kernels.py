"""Roofline of the HBM-bound helper kernels (K0 pack, K1b k-means update, list build + gather,
K4 merge, select/refine): achieved GB/s = algorithmic bytes / CUDA-event time against the
measured copy bandwidth in MEASURED_PEAKS.json. One JSON object on stdout."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import newsrecommend_b200.faiss as nf
from newsrecommend_b200 import synth
from newsrecommend_b200._lib import check, lib

try:
    HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    HBM = 6650.0

def timed(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)  # evict L2 between iterations
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps / 1e3

out = {"hbm_peak_gbs": HBM, "kernels": {}}
def rec(name, bytes_, secs, note=""):
    out["kernels"][name] = dict(alg_bytes=bytes_, ms=secs * 1e3, gbs=bytes_ / secs / 1e9, frac=bytes_ / secs / 1e9 / HBM, note=note)

n, d, kp = synth.N_ARTICLES, 250, 256
x = torch.from_numpy(synth.g_skew(n, d, 42)).cuda()
# K0 pack: read n*d*4, write 3 planes + norms
p = nf.PackedMatrix(d); p._reserve(n)
st = None
def pack():
    check(lib.nrb_pack_rows(x.data_ptr(), n, d, x.stride(0), kp, p.raw.data_ptr(), p.hi.data_ptr(), p.lo.data_ptr(), p.norms.data_ptr(), None))
rec("K0 pack_rows (364,047 x 250 -> raw/hi/lo/norms)", n * d * 4 + 3 * n * kp * 4 + n * 4, timed(pack))
p.n = n
# K1b update at the config-2 training size (256*250 rows) and at the full catalog
for rows, k in ((64000, 250), (n, 300)):
    sub = nf.PackedMatrix.from_tensor(x[:rows].contiguous())
    assign = torch.randint(0, k, (rows,), device="cuda", dtype=torch.int64)
    cent = torch.empty((k, d), device="cuda"); h = torch.empty(k, device="cuda")
    wsb = lib.nrb_kmeans_update_workspace(rows, k); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    def upd():
        check(lib.nrb_kmeans_update(sub.raw.data_ptr(), rows, d, kp, assign.data_ptr(), k, cent.data_ptr(), h.data_ptr(), ws.data_ptr(), wsb, None))
    rec(f"K1b kmeans_update (+ counting sort) {rows} rows, k={k}", rows * kp * 4 + rows * 8, timed(upd), "3 sort kernels + 1 update kernel")
# list build + row gather (IVF add layout)
assign = torch.randint(0, 250, (n,), device="cuda", dtype=torch.int64)
off = torch.empty(251, dtype=torch.int32, device="cuda"); order = torch.empty(n, dtype=torch.int32, device="cuda")
wsb = lib.nrb_ivf_build_lists_workspace(n, 250); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
def build():
    check(lib.nrb_ivf_build_lists(assign.data_ptr(), n, 250, off.data_ptr(), order.data_ptr(), ws.data_ptr(), wsb, None))
rec("K3 ivf_build_lists (stable counting sort, 364,047 keys)", n * 8 * 2 + n * 4, timed(build), "keys read twice, order written")
dst = torch.empty((n, kp), device="cuda")
def gather():
    check(lib.nrb_gather_rows(p.raw.data_ptr(), kp, order.data_ptr(), n, dst.data_ptr(), None))
rec("K3 gather_rows (list-contiguous plane)", 2 * n * kp * 4 + n * 4, timed(gather))
# K4 merge: 8 shards x 50,000 queries x k=50
G, nq, k = 8, 50000, 50
Dp = torch.sort(torch.randn(G, nq, k, device="cuda"), dim=2, descending=True).values.contiguous()
Ip = torch.randint(0, 1 << 40, (G, nq, k), device="cuda", dtype=torch.int64)
Dm = torch.empty((nq, k), device="cuda"); Im = torch.empty((nq, k), dtype=torch.int64, device="cuda")
def merge():
    check(lib.nrb_merge_topk(Dp.data_ptr(), Ip.data_ptr(), G, nq, k, 0, Dm.data_ptr(), Im.data_ptr(), None))
rec("K4 merge_topk (8 x 50,000 x 50)", G * nq * k * 12 + nq * k * 12, timed(merge), "reads only the consumed heads in practice")
print(json.dumps(out))
