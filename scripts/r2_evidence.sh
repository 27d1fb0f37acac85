#!/bin/bash
# Round-2 evidence run (one GPU): full GPU suite, both bench arms, ncu launch lists, one ncu
# --set full capture of the dominant flat kernel and one of the IVF scan kernel, IVF / k-means /
# small-batch / stage benches. Every step has its own timeout; outputs go to gpurun_out/r02f_*.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02f_tests.log 2>&1; tail -3 $O/r02f_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02f_bench_default.json 2> $O/r02f_err.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02f_bench_reference.json 2>> $O/r02f_err.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02f_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 > $O/r02f_ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:topk_tc3 -s 3 -c 1 -o $O/r02f_prof_tc16 -f python bench.py --steps 2 --warmup 3 > $O/r02f_ncu2.log 2>&1
timeout 400 python scripts/bench_ivf.py > $O/r02f_ivf.json 2>> $O/r02f_err.log
timeout 400 python scripts/bench_kmeans.py > $O/r02f_kmeans.json 2>> $O/r02f_err.log
timeout 600 python scripts/bench_small.py > $O/r02f_small.json 2>> $O/r02f_err.log
timeout 400 python scripts/bench_stage.py > $O/r02f_stage.json 2>> $O/r02f_err.log
timeout 400 python scripts/bench_robustness.py > $O/r02f_robustness.json 2>> $O/r02f_err.log
timeout 600 python scripts/bench_sweep.py > $O/r02f_sweep.jsonl 2>> $O/r02f_err.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02f_ivf_launches.csv python scripts/ivf_ncu_target.py > $O/r02f_ivf_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02f_km_launches.csv python scripts/bench_kmeans_ncu.py > $O/r02f_km_ncu.log 2>&1
tail -5 $O/r02f_err.log
python - <<'PY'
import json
for f in ["r02f_bench_default", "r02f_bench_reference", "r02f_ivf", "r02f_kmeans", "r02f_small", "r02f_stage"]:
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        s = json.dumps(d)
        print(f, s[:1800])
    except Exception as e:
        print(f, "FAILED", e)
PY
