"""Stage breakdown of the catalog-sharded config-1 search under torchrun (one rank per GPU):
CUDA-event time of search_local (K0 + K2 + refine), wire pack, exchange (start -> wait) and K4
merge per step, the K2 kernel time inside search_local, and the host time to queue one step.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/profile_sharded.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import newsrecommend_b200.faiss as nf  # noqa: E402
from newsrecommend_b200 import _lib, synth  # noqa: E402
from newsrecommend_b200.sharded import ShardedIndexFlat  # noqa: E402

NB, D, NQ, K = 364_047, 250, int(os.environ.get("NQ", 50_000)), 50
xb, topics = synth.g_skew(NB, D, 42, return_topics=True)
xq = torch.from_numpy(synth.user_profiles(xb, topics, NQ, 43)).cuda()
idx = ShardedIndexFlat(D, nf.METRIC_INNER_PRODUCT)
idx.add_global(xb)
stages = {}


def timed(name, fn):
    def wrap(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = fn(*a, **kw)
        e1.record()
        stages.setdefault(name, []).append((e0, e1, time.perf_counter() - t0))
        return out
    return wrap


class Codec:
    pack = staticmethod(timed("pack_topk", idx.codec.pack))
    merge = staticmethod(timed("merge_k4", idx.codec.merge))


idx.codec = Codec
idx.search_local = timed("search_local", idx.search_local)
orig_start = idx._exchange_start


def start(P, per):
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    recv, work = orig_start(P, per)

    class W:
        def wait(self_inner):
            work.wait()
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            stages.setdefault("exchange", []).append((e0, e1, 0.0))
    return recv, (W() if work is not None else None)


idx._exchange_start = start
for _ in range(5):
    idx.search(xq, K, gather=False)
torch.cuda.synchronize()
dist.barrier()
stages.clear()
_lib.profile_enable(True)
_lib.profile_read()
STEPS = 20
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
ev0.record()
for _ in range(STEPS):
    idx.search(xq, K, gather=False)
ev1.record()
host_s = time.perf_counter() - t0
torch.cuda.synchronize()
kms, kn = _lib.profile_read()
out = {"rank": rank, "world": world, "nq": NQ, "ms_per_step": ev0.elapsed_time(ev1) / STEPS,
       "host_ms_to_queue_a_step": host_s / STEPS * 1e3, "k2_kernel_ms_per_step": kms / STEPS,
       "device_ms": {k: sum(a.elapsed_time(b) for a, b, _ in v) / STEPS for k, v in stages.items()},
       "host_ms": {k: sum(h for _, _, h in v) / STEPS * 1e3 for k, v in stages.items()}}
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    print(json.dumps(gathered))
dist.destroy_process_group()
