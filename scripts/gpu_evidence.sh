#!/bin/bash
# Evidence run for profiles/: bench arms, ncu launch list, one ncu --set full capture of the
# dominant kernel, and the other BASELINE configs. Every step has its own timeout.
set -u
O=gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 > $O/ev_bench_default.json 2> $O/ev_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --path tc1 > $O/ev_bench_tc1.json 2>> $O/ev_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --path tc > $O/ev_bench_tc.json 2>> $O/ev_err.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/ev_bench_reference.json 2>> $O/ev_err.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ev_launches.csv python bench.py --steps 2 --warmup 1 > $O/ev_ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:topk_tc3 -s 3 -c 1 -o $O/ev_prof_tc16 -f python bench.py --steps 2 --warmup 3 > $O/ev_ncu2.log 2>&1
timeout 400 python scripts/bench_configs.py 4 5 > $O/ev_configs45.json 2>> $O/ev_err.log
timeout 400 python scripts/bench_ivf.py > $O/ev_ivf.json 2>> $O/ev_err.log
timeout 400 python scripts/bench_stage.py > $O/ev_stage.json 2>> $O/ev_err.log
NRB_LIB=$PWD/newsrecommend_b200/libnrb200_trace.so timeout 300 python scripts/trace_timeline.py > $O/ev_trace.log 2>&1
timeout 300 python -m pytest tests -m gpu -q > $O/ev_tests.log 2>&1; tail -2 $O/ev_tests.log
tail -3 $O/ev_err.log
python - <<'PY'
import json
for f in ["ev_bench_default", "ev_bench_tc1", "ev_bench_tc", "ev_bench_reference"]:
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("kernel_ms_avg"),
              (d.get("roofline") or {}).get("frac"), d.get("cpu_baseline"), d.get("clocks"))
    except Exception as e:
        print(f, "FAILED", e)
PY
