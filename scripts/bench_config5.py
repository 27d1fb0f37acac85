"""BASELINE configs[4]: 10,000,000 x 256 catalog row-sharded over the N GPUs of one box, 1,000,000
queries, top-100 inner product, NCCL exchange + K4 merge (north_star item 4).

  python scripts/bench_config5.py                         (N = 1)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P scripts/bench_config5.py [--exchange alltoall|allgather] [--steps 2]

Strong scaling: the job is fixed, value = 1,000,000 queries / max-over-ranks device time of one
search. Every rank generates ITS catalog rows on its GPU (seeded per 1M-row chunk, so the catalog
does not depend on N) and the full query batch (same seed on every rank; a checksum all-reduce
verifies that the ranks hold identical queries). Parity: a query sample goes through the oracle
chunk by chunk on every rank's own rows, the per-shard oracle candidates are gathered and merged
on the host, and the sharded search must reproduce that global top-100. One JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

NB, D, NQ, K, CH = 10_000_000, 256, 1_000_000, 100, 1_000_000
G, R = 4096, 16


def topic_model(dev):
    g = torch.Generator(device=dev).manual_seed(45)
    w = 1.0 / torch.arange(1, G + 1, device=dev, dtype=torch.float32) ** 0.8
    w /= w.sum()
    centers = torch.randn((G, R), generator=g, device=dev)
    q, _ = torch.linalg.qr(torch.randn((D, R), generator=g, device=dev))
    return w, centers, q.T.contiguous()


def gen_rows(n, seed, model, dev, noise=0.02):
    w, centers, W = model
    g = torch.Generator(device=dev).manual_seed(seed)
    comp = torch.multinomial(w, n, replacement=True, generator=g)
    z = centers[comp] + 0.7 * torch.randn((n, R), generator=g, device=dev)
    return (z @ W + noise * torch.randn((n, D), generator=g, device=dev)).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--exchange", default="alltoall", choices=["alltoall", "allgather"])
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--parity-queries", type=int, default=256)
    ap.add_argument("--nq", type=int, default=NQ)
    ap.add_argument("--breakdown", action="store_true", help="also time the stages of one search with CUDA events")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import newsrecommend_b200.faiss as nf
    from newsrecommend_b200 import _lib
    from newsrecommend_b200.parity import compare_topk
    from newsrecommend_b200.sharded import ShardedIndexFlat, owned_rows, shard_range
    from oracle import faiss_oracle as fo
    fo.build()

    model = topic_model(dev)
    nq = args.nq
    xq = gen_rows(nq, 46, model, dev, noise=0.05)
    if world > 1:
        cs = xq.double().sum().reshape(1)
        lo_, hi_ = cs.clone(), cs.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        assert float(lo_) == float(hi_), "ranks generated different query batches"
    ns = args.parity_queries
    pick = np.linspace(0, nq - 1, ns).astype(np.int64)
    xq_s = xq[torch.from_numpy(pick).to(dev)].cpu().numpy()

    # ---- build: this rank's rows, 1M-row chunks (chunk c is the same whatever N is)
    lo, hi = shard_range(NB, world, rank)
    idx = ShardedIndexFlat(D, nf.METRIC_INNER_PRODUCT, exchange=args.exchange)
    local_index = idx.local
    t_add = 0.0
    cand_D, cand_I = [], []
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=max(1, len(os.sched_getaffinity(0)) // world))
    except Exception:  # noqa: BLE001
        pass
    for c in range(NB // CH):
        c0, c1 = c * CH, (c + 1) * CH
        a, b = max(lo, c0), min(hi, c1)
        if a >= b:
            continue
        rows = gen_rows(CH, 1000 + c, model, dev)[a - c0:b - c0]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        local_index.add(rows)
        torch.cuda.synchronize()
        t_add += time.perf_counter() - t0
        Dc, Ic = fo.knn_fast(xq_s, rows.cpu().numpy(), K, 0)  # oracle on this chunk of this shard
        cand_D.append(Dc)
        cand_I.append(Ic + a)
        del rows
    idx.id_base, idx.ntotal = lo, NB
    idx._bases_host = [shard_range(NB, world, r)[0] for r in range(world)]
    torch.cuda.empty_cache()

    # oracle global answer for the sample: merge of every shard's chunk candidates
    Dl, Il = np.concatenate(cand_D, 1), np.concatenate(cand_I, 1)
    o = np.argsort(-Dl, axis=1, kind="stable")[:, :K]
    Dl, Il = np.take_along_axis(Dl, o, 1), np.take_along_axis(Il, o, 1)
    if world > 1:
        gd = [None] * world
        dist.all_gather_object(gd, (Dl, Il))
        Dl, Il = np.concatenate([g[0] for g in gd], 1), np.concatenate([g[1] for g in gd], 1)
        o = np.argsort(-Dl, axis=1, kind="stable")[:, :K]
        Dl, Il = np.take_along_axis(Dl, o, 1), np.take_along_axis(Il, o, 1)
    Do, Io = Dl, Il

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-timed search (queries resident; results stay partitioned by query range)
    idx.search(xq[: min(nq, 2 * 132_608)], K, gather=False)  # warm-up on two chunks
    sync_all()
    f0 = int(_lib.lib.nrb_fallback_query_count())
    _lib.profile_enable(True)
    _lib.profile_read()
    times = []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        Dm, Im, spans = idx.search(xq, K, gather=False)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t))
    kms, kn = _lib.profile_read()
    _lib.profile_enable(False)
    ms = min(times)
    # parity: the sample rows this rank owns
    rows_own = np.concatenate([np.arange(a, b) for a, b in spans]) if spans else np.empty(0, np.int64)
    pos = {int(r): i for i, r in enumerate(rows_own.tolist())} if rows_own.size < 5_000_000 else None
    mine = [(j, pos[int(p)]) for j, p in enumerate(pick) if pos is not None and int(p) in pos]
    rep = dict(ok=True, recall=1.0, exact_ordered=1.0, max_rel_score_err=0.0, tie_exempt_queries=0)
    if mine:
        js = np.array([m[0] for m in mine])
        ps = torch.tensor([m[1] for m in mine], device=dev)
        rep = compare_topk(Dm[ps].cpu().numpy(), Im[ps].cpu().numpy(), Do[js], Io[js], 0)
    flags = torch.tensor([1.0 if rep["ok"] else 0.0, float(len(mine))], device=dev, dtype=torch.float64)
    if world > 1:
        okt = flags[:1].clone()
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        dist.all_reduce(flags[1:], op=dist.ReduceOp.SUM)
        flags[0] = okt[0]
    del Dm, Im
    torch.cuda.empty_cache()

    breakdown = None
    if args.breakdown:
        # one more search with every stage bracketed by CUDA events (device time) and perf_counter (host time)
        ev, host = {}, {}

        def wrap(obj, name, label):
            fn = getattr(obj, name)

            def inner(*a, **kw):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record()
                out = fn(*a, **kw)
                e1.record()
                host[label] = host.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
                ev.setdefault(label, []).append((e0, e1))
                return out
            setattr(obj, name, inner)
            return fn
        saved = [(idx, "search_local", wrap(idx, "search_local", "search_local (K0 + K2 + refine [+ fallback])")),
                 (idx.codec, "pack", wrap(idx.codec, "pack", "pack_topk")),
                 (idx.codec, "merge", wrap(idx.codec, "merge", "merge (K4)"))]
        sync_all()
        _lib.profile_enable(True)
        _lib.profile_read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        idx.search(xq, K, gather=False)
        e1.record()
        t_host = (time.perf_counter() - t0) * 1e3
        sync_all()
        kms2, kn2 = _lib.profile_read()
        _lib.profile_enable(False)
        for obj, name, fn in saved:
            setattr(obj, name, fn)
        breakdown = dict(total_device_ms=e0.elapsed_time(e1), host_ms_until_last_launch=t_host, k2_kernel_ms=kms2, k2_launches=kn2,
                         device_ms={k: sum(a.elapsed_time(b) for a, b in v) for k, v in ev.items()}, host_ms=host)
        torch.cuda.empty_cache()

    # ---- end to end from page-locked host memory (each rank moves only its own rows over PCIe)
    xq_pin = torch.empty((nq, D), dtype=torch.float32, pin_memory=True)
    xq_pin.copy_(xq)
    D_pin = torch.empty((nq, K), dtype=torch.float32, pin_memory=True).numpy()
    I_pin = torch.empty((nq, K), dtype=torch.int64, pin_memory=True).numpy()
    sync_all()
    t0 = time.perf_counter()
    idx.search_host(xq_pin.numpy(), K, D_pin, I_pin)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    hbm = torch.tensor([torch.cuda.max_memory_allocated() / 1e9], device=dev)
    if world > 1:
        dist.all_reduce(hbm, op=dist.ReduceOp.MAX)
    if rank == 0:
        alg = 2.0 * nq * (hi - lo) * D
        out = dict(config="BASELINE configs[4]: 10M x 256 catalog row-sharded, %d queries, top-100 IP" % nq, n_gpus=world,
                   exchange=args.exchange, value=nq / (ms / 1e3), unit="queries/s", scaling="strong", ms_per_search=ms,
                   ms_all_steps=times, kernel_ms_per_search=kms / max(1, args.steps), kernel_launches=kn,
                   alg_tflops_per_gpu=alg / (kms / max(1, args.steps) / 1e3) / 1e12 if kms else None,
                   e2e_qps=nq / float(te), e2e_h2d_bytes=nq * D * 4, e2e_d2h_bytes=nq * K * 12,
                   add_seconds_per_gpu=t_add, rows_per_gpu=hi - lo, hbm_gb_max=float(hbm),
                   fallback_queries=int(_lib.lib.nrb_fallback_query_count()) - f0, breakdown=breakdown,
                   parity_sample=dict(queries=int(flags[1]), ok=bool(flags[0] >= 1.0), exact_ordered=rep["exact_ordered"],
                                      recall=rep["recall"], max_rel_score_err=rep["max_rel_score_err"],
                                      how="oracle port per 1M-row chunk of every shard, host merge, compared on the owning rank"))
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and not bool(flags[0] >= 1.0):
        raise SystemExit("config 5: parity FAILED")


if __name__ == "__main__":
    main()
