"""Timeline of the fp16 filter kernel's tile hand-offs on cluster 0's leader CTA.

Needs the trace build: `make -C newsrecommend_b200/csrc trace`, then
`NRB_LIB=newsrecommend_b200/libnrb200_trace.so python scripts/trace_timeline.py`.
Prints, per phase of the unit, where a tile's cycles go for the MMA warp and the selection warps.
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import newsrecommend_b200.faiss as nf  # noqa: E402
from newsrecommend_b200 import _lib, synth  # noqa: E402

NB, NQ, D, K = 364047, int(os.environ.get("NQ", 50000)), 250, 50
xb, topics = synth.g_skew(NB, D, 42, return_topics=True)
xq = synth.user_profiles(xb, topics, NQ, 43)
if os.environ.get("NRB_TRACE_DUP"):  # near-duplicate-heavy catalog: m copies per article (scripts/bench_robustness.py)
    m = int(os.environ["NRB_TRACE_DUP"])
    rng = np.random.default_rng(7)
    base = xb[: NB // m]
    rows = np.tile(np.arange(NB // m), m + 1)[:NB]
    xb = np.ascontiguousarray((base[rows] * (1.0 + 1e-4 * rng.standard_normal((NB, 1), dtype=np.float32))
                               + 1e-4 * np.linalg.norm(base[rows], axis=1, keepdims=True) / np.sqrt(D)
                               * rng.standard_normal((NB, D), dtype=np.float32)).astype(np.float32))
idx = nf.IndexFlatIP(D)
idx.add(xb)
q = nf.PackedMatrix.from_tensor(torch.from_numpy(xq).cuda(), planes=idx._query_planes(K))
idx.search_packed(q, K)
pb = np.zeros((9, 64, 4), dtype=np.int64)
_lib.lib.nrb_debug_trace_prune_read(C.c_void_p(pb.ctypes.data), C.c_size_t(pb.nbytes), 1)  # reset the prune counters
idx.search_packed(q, K)
torch.cuda.synchronize()
assert _lib.lib.nrb_debug_trace_prune_read(C.c_void_p(pb.ctypes.data), C.c_size_t(pb.nbytes), 0) == 0
T = 4096
buf = np.zeros((9, T, 4), dtype=np.int64)
rc = _lib.lib.nrb_debug_trace_read(C.c_void_p(buf.ctypes.data), C.c_size_t(buf.nbytes))
assert rc == 0, rc
mma, epi = buf[0], buf[1:]
n = int((mma[:, 1] > 0).sum())
out = {"tiles_traced": n}
cad = np.diff(mma[:n, 0])


def stats(x):
    x = np.asarray(x, dtype=np.float64)
    return {"mean": float(x.mean()), "p50": float(np.median(x)), "p90": float(np.percentile(x, 90)), "max": float(x.max())}


for name, lo, hi in (("tiles 2-31", 2, 32), ("tiles 32-255", 32, 256), ("tiles 256-1023", 256, 1024),
                     ("tiles 1024-1400", 1024, 1400), ("tiles 1500-2800 (2nd unit)", 1500, 2800)):
    hi = min(hi, n - 1)
    if hi <= lo:
        continue
    r = {"mma_cadence": stats(cad[lo:hi]),
         "mma_wait_tempty": stats(mma[lo:hi, 0] - mma[lo - 1:hi - 1, 1]),
         "mma_issue": stats(mma[lo:hi, 1] - mma[lo:hi, 0])}
    w, ld, pr, gap = [], [], [], []
    for e in epi:
        w.append(e[lo:hi, 1] - e[lo:hi, 0])
        ld.append(e[lo:hi, 2] - e[lo:hi, 1])
        pr.append(e[lo:hi, 3] - e[lo:hi, 2])
        gap.append(e[lo + 1:hi + 1, 0] - e[lo:hi, 3])
    r["epi_wait_tfull"] = stats(np.concatenate(w))
    r["epi_load_release"] = stats(np.concatenate(ld))
    r["epi_select"] = stats(np.concatenate(pr))
    r["epi_select_max_over_warps"] = stats(np.max(np.stack(pr), axis=0))
    r["epi_between_tiles"] = stats(np.concatenate(gap))
    # lag of each warp's release behind the MMA warp seeing tempty of tile+2
    rel = np.max(np.stack([e[lo:hi, 2] for e in epi]), axis=0)
    r["tempty_seen_minus_last_local_release(t->t+2)"] = stats(mma[lo + 2:hi + 2, 0] - rel)
    r["tfull_seen_minus_mma_commit"] = stats(np.min(np.stack([e[lo:hi, 1] for e in epi]), axis=0) - mma[lo:hi, 1])
    out[name] = r
# scheduled prunes of the first unit (11 of them): per warp, cycles waiting for the partner warp,
# pruning its 16 rows, waiting for the partner to finish
pr = pb[1:, :11, :]
out["scheduled_prunes_first_unit"] = {
    "wait_partner_in": stats(pr[:, :, 1] - pr[:, :, 0]), "prune_16_rows": stats(pr[:, :, 2] - pr[:, :, 1]),
    "wait_partner_out": stats(pr[:, :, 3] - pr[:, :, 2]),
    "prune_16_rows_by_prune_index_mean": [float(x) for x in (pr[:, :, 2] - pr[:, :, 1]).mean(axis=0)]}
print(json.dumps(out, indent=1))
json.dump(out, open("gpurun_out/trace_timeline.json", "w"), indent=1)
np.save("gpurun_out/trace_raw.npy", buf[:, :3000])
