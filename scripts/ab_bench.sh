#!/bin/bash
# A/B timing of library builds on the same box: bash scripts/ab_bench.sh libA.so libB.so ...
# (paths relative to newsrecommend_b200/). Alternates the builds twice to average out clock drift.
for round in 1 2; do
  for lib in "$@"; do
    NRB_LIB=$PWD/newsrecommend_b200/$lib timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/ab_tmp.json 2>> gpurun_out/ab_err.log
    python - "$lib" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/ab_tmp.json"))
    print(sys.argv[1], "kernel_ms %.3f" % d["roofline"]["kernel_ms_avg"], "step_ms %.3f" % d["ms_per_step"],
          "e2e %.0f" % d["e2e"]["value"], "mhz", d["clocks"]["sm_mhz"], "ok", d["parity_sample"]["ok"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
done
