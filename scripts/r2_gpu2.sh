#!/bin/bash
# round-2 check on a 2-GPU box: GPU suite (incl. the torchrun NCCL test), bench N=1, bench N=2 (both layouts)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; tail -5 gpurun_out/r2_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 600 gpurun_out/r2_bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 1500 gpurun_out/r2_bench_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --exchange allgather --no-alt > gpurun_out/r2_bench_n2_ag.json 2> gpurun_out/r2_bench_n2_ag.err; tail -c 600 gpurun_out/r2_bench_n2_ag.err
python - <<'PY'
import json
for f in ["r2_bench_n1", "r2_bench_n2", "r2_bench_n2_ag"]:
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        r = d["roofline"]
        print(f, "q/s %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"],
              "kernel_ms %.3f" % r["kernel_ms_avg"], "frac %.3f" % r["frac"], d.get("clocks"), d.get("parity_sample"),
              d.get("other_decomposition"))
    except Exception as e:
        print(f, "FAILED", e)
PY
