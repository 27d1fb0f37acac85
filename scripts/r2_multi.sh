#!/bin/bash
# usage: bash scripts/r2_multi.sh N [tests]   -- multi-GPU evidence on a box with N GPUs:
# bench.py --gpus N (catalog-sharded, all-to-all by query range), config 5 (10M x 256, 1M queries,
# top-100) sharded N ways; with `tests`, first the NCCL torchrun tests.
N=${1:-2}
O=gpurun_out
mkdir -p $O
if [ "${2:-}" = "tests" ]; then
  timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_kmeans_ivf.py -x -q > $O/r02f_multi_tests_n$N.log 2>&1; tail -5 $O/r02f_multi_tests_n$N.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02f_bench_n$N.json 2> $O/r02f_bench_n$N.err; tail -c 800 $O/r02f_bench_n$N.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/bench_config5.py --breakdown > $O/r02f_cfg5_n$N.json 2> $O/r02f_cfg5_n$N.err; tail -c 800 $O/r02f_cfg5_n$N.err
python - <<PY
import json
for f in ["r02f_bench_n$N", "r02f_cfg5_n$N"]:
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, json.dumps(d)[:3000])
    except Exception as e:
        print(f, "FAILED", e)
PY
