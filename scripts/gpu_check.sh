#!/bin/bash
# One gpurun call: the GPU parity suite, then the flat benchmark on both precision paths.
# Usage (on the GPU box, from the repo root): bash scripts/gpu_check.sh [pytest -k expression]
timeout 300 python -m pytest tests -m gpu -x -q ${1:+-k "$1"} > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tc1.json 2> gpurun_out/bench_err.log
timeout 200 python bench.py --steps 10 --warmup 3 --path tc > gpurun_out/bench_tc.json 2>> gpurun_out/bench_err.log
python - <<'PY'
import json
for f in ["bench_tc1", "bench_tc"]:
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        r = d["roofline"]
        print(f, "q/s %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"],
              "kernel_ms %.3f" % r["kernel_ms_avg"], "TF %.1f" % r["achieved"], "frac %.3f" % r["frac"],
              d.get("clocks"), d.get("parity_sample"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -5 gpurun_out/bench_err.log
