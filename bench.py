#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: queries/s of exact top-50 inner-product
search over the Tianchi-news-shaped catalog (364,047 items x 250-d fp32), 50,000 user queries
per step (the search Retrieval.py:32 performs per user, in its north-star IndexFlatIP form).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

* our arm: one process per GPU (torchrun for N > 1). N = 1: the whole catalog on one B200.
  N > 1: the catalog is row-sharded over the ranks, every rank searches all queries on its
  shard, per-shard (D, I) are all-gathered over NCCL and merged by the K4 kernel ("strong"
  scaling: the job is fixed, value = queries of one step x steps / max-over-ranks device time).
* --impl reference: the reference's CPU path for the same search. faiss (the library the
  reference calls) is not installable in this image, so this times the faiss-equivalent oracle
  port (oracle/faiss_oracle.py: OpenBLAS sgemm blocks + heap handler, all host threads) on a
  bounded query sample per step.

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NB, D, NQ, K = 364_047, 250, 50_000, 50
METRIC = "queries/s, top-50 IP over 364k x 250 items (flat exact search)"
WORKLOAD = "flat_ip_top50: 50,000 queries x 364,047 items x 250-d fp32 (BASELINE configs[0], the config the metric is quoted on)"
ALG_FLOP = 2.0 * NQ * NB * D  # 9.101e12 per step: padding and the 3x TF32 passes not counted


def make_data():
    from newsrecommend_b200 import synth
    xb, topics = synth.g_skew(NB, D, 42, return_topics=True)
    xq = synth.user_profiles(xb, topics, NQ, 43)
    return xb, xq


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}
    try:
        m = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        p = {"hbm_gbs": m["hbm_gbs"], "bf16": m["bf16_tflops"], "bf16_sustained": m["bf16_tflops_sustained"],
             "src": "measured"}
    except Exception:  # noqa: BLE001
        pass
    # TF32 cuBLAS peak measured on this pool by scripts/gpu_probe.py (same method as the bf16
    # figure; profiles/r01_probe.json): 740.7 burst / 604.0 sustained TFLOP/s.
    p["tf32"], p["tf32_sustained"] = 740.7, 604.0
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs must use every host core they can."""
    cores = len(os.sched_getaffinity(0))
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)  # OpenBLAS behind numpy (the sgemm blocks)
    except Exception:  # noqa: BLE001
        pass
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)  # the oracle's OpenMP loops
    except Exception:  # noqa: BLE001
        pass
    return cores


def cpu_baseline(xb, xq, budget_s=12.0):
    """faiss-equivalent oracle port on the host cores, bounded sample of the same workload."""
    from oracle import faiss_oracle as fo
    fo.build()
    cores = use_all_host_threads()
    fo.knn_fast(xq[:512], xb, K, 0)  # warm-up (thread pools, page-in)
    t0 = time.perf_counter()
    fo.knn_fast(xq[:2048], xb, K, 0)
    rate = 2048 / (time.perf_counter() - t0)
    n = int(min(NQ, max(4096, rate * budget_s)))
    n = n // 4096 * 4096 or 4096
    t0 = time.perf_counter()
    fo.knn_fast(xq[:n], xb, K, 0)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"first {n} of the {NQ} queries against the full 364,047-item catalog, {dt:.1f} s, "
                      f"oracle port (OpenBLAS sgemm blocks 4096 x 16384 + faiss heap handler, OpenMP)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import faiss_oracle as fo
    fo.build()
    xb, xq = make_data()
    cores = use_all_host_threads()
    fo.knn_fast(xq[:512], xb, K, 0)
    t0 = time.perf_counter()
    fo.knn_fast(xq[:2048], xb, K, 0)
    rate = 2048 / (time.perf_counter() - t0)
    steps, warm = args.steps, args.warmup
    budget = 150.0 / max(1, steps + warm)  # whole run within a few minutes
    n = int(min(NQ, max(2048, rate * budget)))
    n = max(2048, n // 2048 * 2048)
    for _ in range(warm):
        fo.knn_fast(xq[:n], xb, K, 0)
    t0 = time.perf_counter()
    for _ in range(steps):
        fo.knn_fast(xq[:n], xb, K, 0)
    dt = time.perf_counter() - t0
    v = n * steps / dt
    sample = (f"each step = first {n} of {NQ} queries against the full catalog; faiss is not installable here, "
              f"this is the faiss-equivalent oracle port on all {cores} host threads")
    args.out.emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "queries_per_step": n, "k": K},
        "cpu_baseline": {"value": v, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import newsrecommend_b200.faiss as nf
    from newsrecommend_b200 import _lib
    from newsrecommend_b200.parity import compare_topk
    from newsrecommend_b200.sharded import ShardedIndexFlat

    xb, xq = make_data()
    path_id = {"auto": _lib.PATH_AUTO, "tc": _lib.PATH_TC, "tc1": _lib.PATH_TC1, "tc16": _lib.PATH_TC16}[args.path]
    one_pass = args.path in ("auto", "tc1", "tc16")
    f16 = args.path in ("auto", "tc16")

    # Decompositions for N > 1 (the job stays config 1: 50,000 queries x 364,047 items):
    #  "queries": the packed catalog (1.1 GB) is replicated on every GPU and the query batch is
    #             split into N contiguous slices -- no data-path collective (the headline);
    #  "catalog": north_star item 4 -- catalog rows sharded, every rank searches all queries on
    #             its shard, NCCL all-gather of the per-shard (D, I) + K4 merge on every rank.
    def q_slice(n):
        per = (n + world - 1) // world
        return min(n, rank * per), min(n, (rank + 1) * per)

    def build(mode):
        if mode == "catalog" and world > 1:
            idx = ShardedIndexFlat(D, nf.METRIC_INNER_PRODUCT)
            idx.add_global(xb)
            idx.local.path = path_id
            lo, hi = 0, NQ
            planes = idx.local._query_planes(K)
            search = lambda q: idx.search(q, K)  # noqa: E731
            rows = idx.local.ntotal
        else:
            idx = nf.IndexFlatIP(D)
            idx.add(xb)
            idx.path = path_id
            lo, hi = q_slice(NQ) if world > 1 else (0, NQ)
            planes = idx._query_planes(K)
            search = lambda q: idx.search_packed(q, K)  # noqa: E731
            rows = idx.ntotal
        xq_dev = torch.from_numpy(xq[lo:hi]).cuda()
        # host side of the end-to-end leg: page-locked query batch and result arrays, handed to
        # the public API as numpy arrays (faiss's search(x, k, D, I) convention)
        xq_pin = torch.from_numpy(xq[lo:hi]).pin_memory().numpy()
        D_pin = torch.empty((hi - lo, K), dtype=torch.float32, pin_memory=True).numpy()
        I_pin = torch.empty((hi - lo, K), dtype=torch.int64, pin_memory=True).numpy()

        def step_device():
            q = nf.PackedMatrix.from_tensor(xq_dev, planes=planes)  # K0 on the fresh query batch
            return search(q)

        def step_e2e():
            if mode == "catalog" and world > 1:
                xd = torch.from_numpy(xq_pin).cuda(non_blocking=True)  # H2D from pinned host memory
                q = nf.PackedMatrix.from_tensor(xd, planes=planes)
                Dd, Id = search(q)
                return nf._to_host_pair(Dd, Id, D_pin, I_pin)  # D2H + stream sync
            return idx.search(xq_pin, K, D=D_pin, I=I_pin)  # the call a user makes: numpy in, numpy out

        return dict(step_device=step_device, step_e2e=step_e2e, lo=lo, hi=hi, rows=rows, mode=mode)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def measure(run, sampler=None):
        for _ in range(args.warmup):
            run["step_device"]()
        sync_all()
        if sampler is not None:
            sampler.start()
        _lib.profile_enable(True)
        _lib.profile_read()
        n0 = _lib.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            run["step_device"]()
        ev1.record()
        sync_all()
        ms = ev0.elapsed_time(ev1)
        launches = _lib.launch_count() - n0
        kern_ms, kern_n = _lib.profile_read()
        _lib.profile_enable(False)
        clocks = sampler.stop() if sampler is not None else None
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # end to end through the public API with host buffers
        for _ in range(max(1, min(args.warmup, 3))):
            run["step_e2e"]()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            Dh, Ih = run["step_e2e"]()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return dict(ms=ms, launches=launches, kern_ms=kern_ms, kern_n=kern_n, clocks=clocks,
                    e2e_s=float(t.item()), Dh=Dh, Ih=Ih)

    main_mode = args.shard if world > 1 else "queries"
    run = build(main_mode)
    res = measure(run, ClockSampler(local) if rank == 0 else None)
    ms, launches, kern_ms, kern_n, clocks, e2e_s = (res[k] for k in ("ms", "launches", "kern_ms", "kern_n", "clocks", "e2e_s"))
    Dh, Ih = res["Dh"], res["Ih"]
    run = dict(lo=run["lo"], hi=run["hi"], rows=run["rows"], mode=main_mode)  # drop the index (frees HBM)
    alt = None
    if world > 1 and args.alt:
        alt_mode = "catalog" if main_mode == "queries" else "queries"
        torch.cuda.empty_cache()
        r2 = measure(build(alt_mode))
        alt = {"decomposition": alt_mode, "value": NQ * args.steps / (r2["ms"] / 1e3), "ms_per_step": r2["ms"] / args.steps,
               "e2e": NQ * args.steps / r2["e2e_s"], "kernel_ms_avg": r2["kern_ms"] / max(1, r2["kern_n"])}

    if rank == 0:
        pk = peaks()
        value = NQ * args.steps / (ms / 1e3)
        # roofline of the dominant kernel (topk_tc_kernel): algorithmic flops of this rank's shard
        alg = 2.0 * (run["hi"] - run["lo"]) * run["rows"] * D  # this rank's share of the job
        kavg_s = kern_ms / max(1, kern_n) / 1e3
        achieved = alg / kavg_s / 1e12 if kavg_s > 0 else 0.0
        tf32_peak = pk["tf32_sustained"]
        passes = 1.0 if one_pass else 3.0
        if f16:
            # fp16 operands run at the bf16 rate: the denominator is MEASURED_PEAKS' dense bf16 figure
            # (sustained: the kernel is timed inside back-to-back steps)
            peak = pk["bf16_sustained"]
            kname = ("topk_tc3_kernel<IP, fp16> (tcgen05 kind::f16 filter on power-of-two scaled fp16 planes "
                     "with error margin; exact fp32 refine follows)")
            note = ("peak = MEASURED_PEAKS (%s) dense bf16/fp16 sustained %.1f TFLOP/s (burst %.1f); cuBLAS TF32 "
                    "sustained on this pool %.1f" % (pk["src"], pk["bf16_sustained"], pk["bf16"], tf32_peak))
            pipe_frac = achieved * 256.0 / 250.0 / pk["bf16_sustained"]
        else:
            peak = tf32_peak / passes
            kname = ("topk_tc3_kernel<IP> (tcgen05 1xTF32 filter with error margin; exact fp32 refine follows)"
                     if one_pass else "topk_tc2_kernel<IP> (tcgen05 3xTF32 + fused selection)")
            note = ("peak = cuBLAS TF32 sustained %.1f TFLOP/s (scripts/gpu_probe.py, same method as "
                    "MEASURED_PEAKS.json; profiles/r01_probe.json) / %d TF32 pass(es) per product; "
                    "MEASURED_PEAKS (%s) bf16 sustained %.1f" % (tf32_peak, int(passes), pk["src"], pk["bf16_sustained"]))
            pipe_frac = achieved * passes * 256.0 / 250.0 / tf32_peak
        roof = {
            "bound": "tensor",
            "kernel": kname,
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_note": note,
            "tensor_pipe_frac": pipe_frac,
            "frac_of_bf16_peak": achieved / pk["bf16_sustained"],
            "kernel_ms_avg": kavg_s * 1e3, "kernel_launches_timed": kern_n,
            "alg_flop_per_launch": alg,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at N = 1
            # (ncu --set full; fp16 filter: profiles/r01_ncu_full_tc16_fp16_filter.csv, 0.86 GB + 0.22 GB;
            # tf32 filter: profiles/r01_ncu_full_tc3_final.csv, 7.53 GB + 1.22 GB); the
            # 3xTF32 kernel was captured before the last planner change (profiles/r01_ncu_full_prof_tc2.csv)
            "traffic": (1.08e9 if f16 else 8.75e9 if one_pass else 50.2e9) if world == 1 else None,
        }
        # correctness spot check inside the bench: a query sample against the oracle
        from oracle import faiss_oracle as fo
        fo.build()
        ns = min(512, run["hi"] - run["lo"])
        Do, Io = fo.knn_fast(xq[run["lo"]:run["lo"] + ns], xb, K, 0)
        rep = compare_topk(Dh[:ns], Ih[:ns], Do, Io, 0)
        out = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": ("f32 (fp16 tcgen05 filter with a proven error margin + exact fp32 rescoring)" if f16 else
                      "f32 (1xTF32 tcgen05 filter + exact fp32 rescoring)" if one_pass
                      else "f32 (3xTF32 tcgen05, fp32 accumulate)"),
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "k": K,
                       "parallelism": ("single GPU" if world == 1 else
                                       (f"queries split over {world} GPUs, packed catalog (1.3 GB) replicated, no data-path collective"
                                        if main_mode == "queries" else
                                        f"catalog row-sharded over {world} GPUs, NCCL all-gather + K4 merge")),
                       "l2": ("inputs larger than L2 (the kernel streams the 186 MB fp16 catalog plane, the refine gathers "
                              "from the 373 MB fp32 plane; L2 is 126 MB)" if f16 else
                              "inputs larger than L2 (catalog hi+lo planes 746 MB per full catalog)"),
                       "path": args.path, "fallback_queries": int(_lib.lib.nrb_fallback_query_count()),
                       "timed": "K0 query pack + K2 tcgen05 distance/selection + select" + (" + exact refine" if one_pass else "") +
                                (" + NCCL all-gather + K4 merge" if (world > 1 and main_mode == "catalog") else "")},
            "roofline": roof,
            "e2e": {"value": NQ * args.steps / e2e_s, "unit": "queries/s",
                    "h2d_bytes_per_step": NQ * D * 4, "d2h_bytes_per_step": NQ * K * 12},
            "gpu_launches": int(launches), "clocks": clocks, "other_decomposition": alt,
            "parity_sample": {"queries": ns, "ok": rep["ok"], "recall_at_50": rep["recall"],
                              "exact_ordered": rep["exact_ordered"], "max_rel_score_err": rep["max_rel_score_err"]},
        }
        if world == 1 and not os.environ.get("NRB_BENCH_SKIP_CPU"):  # skipped only for ncu captures
            out["cpu_baseline"] = cpu_baseline(xb, xq)
        args.out.emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class _StdoutGuard:
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON
    line on stdout. Route fd 1 to stderr for the whole run and emit the JSON on the saved fd."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, obj):
        sys.stdout.flush()
        os.write(self.real, (json.dumps(obj) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shard", default="queries", choices=["queries", "catalog"],
                    help="N > 1: split the query batch (catalog replicated; default) or shard the catalog "
                         "(north_star item 4: per-shard top-k + NCCL all-gather + merge)")
    ap.add_argument("--no-alt", dest="alt", action="store_false",
                    help="N > 1: do not also time the other decomposition")
    ap.add_argument("--path", default="auto", choices=["auto", "tc", "tc1", "tc16"],
                    help="auto/tc16 = fp16 filter + exact fp32 refine (default), tc1 = the same with the 1xTF32 "
                         "filter, tc = 3xTF32")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    args.out = _StdoutGuard()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
