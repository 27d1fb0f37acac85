"""Timeline of the LAST tcgen05 launch of an IVF search (phase B of the last query batch) on
cluster 0's leader CTA; needs the trace build (see scripts/trace_timeline.py). Saves the raw
buffers to gpurun_out/trace_ivf_raw.npz and prints a summary JSON: cycles per tile inside units,
cycles between units (gaps in the MMA cadence), scheduled prunes."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import newsrecommend_b200.faiss as nf  # noqa: E402
from newsrecommend_b200 import _lib, synth  # noqa: E402

NQ = int(os.environ.get("NQ", 131072))
xb, topics = synth.g_skew(synth.N_ARTICLES, 250, 42, return_topics=True)
xq = synth.user_profiles(xb, topics, NQ, 44)
ivf = nf.IndexIVFFlat(nf.IndexFlatIP(250), 250, 250, nf.METRIC_INNER_PRODUCT)
xb_d, xq_d = torch.from_numpy(xb).cuda(), torch.from_numpy(xq).cuda()
ivf.train(xb_d)
ivf.add(xb_d)
ivf.nprobe = 16
ivf.search(xq_d, 50)
pb = np.zeros((9, 64, 4), dtype=np.int64)
_lib.lib.nrb_debug_trace_prune_read(C.c_void_p(pb.ctypes.data), C.c_size_t(pb.nbytes), 1)
ivf.search(xq_d, 50)
torch.cuda.synchronize()
_lib.lib.nrb_debug_trace_prune_read(C.c_void_p(pb.ctypes.data), C.c_size_t(pb.nbytes), 0)
T = 4096
buf = np.zeros((9, T, 4), dtype=np.int64)
assert _lib.lib.nrb_debug_trace_read(C.c_void_p(buf.ctypes.data), C.c_size_t(buf.nbytes)) == 0
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/trace_ivf_raw.npz", tiles=buf, prunes=pb)
mma = buf[0]
n = int((mma[:, 1] > 0).sum())
cad = np.diff(mma[:n, 0]).astype(np.float64)
gaps = cad[cad > 8000]
inside = cad[cad <= 8000]
epi = buf[1:]
sel = np.concatenate([(e[:n, 3] - e[:n, 2]) for e in epi]).astype(np.float64)
out = dict(tiles_traced=n, total_cycles=float(mma[n - 1, 1] - mma[0, 0]) if n else 0,
           unit_boundaries=int(gaps.size), tiles_per_unit=float(n / max(1, gaps.size + 1)),
           cadence_inside_units=dict(mean=float(inside.mean()), p50=float(np.median(inside)), p90=float(np.percentile(inside, 90))),
           gap_between_units=dict(mean=float(gaps.mean()) if gaps.size else 0, p50=float(np.median(gaps)) if gaps.size else 0,
                                  p90=float(np.percentile(gaps, 90)) if gaps.size else 0, max=float(gaps.max()) if gaps.size else 0,
                                  total=float(gaps.sum())),
           epi_select=dict(mean=float(sel.mean()), p90=float(np.percentile(sel, 90))),
           prunes_recorded=int((pb[1:, :, 0] > 0).sum()),
           prune_cycles_mean=float((pb[1:, :, 2] - pb[1:, :, 1])[pb[1:, :, 0] > 0].mean()) if (pb[1:, :, 0] > 0).any() else 0)
print(json.dumps(out))
