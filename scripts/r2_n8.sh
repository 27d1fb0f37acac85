#!/bin/bash
# N GPUs: bench.py with the default layout (+ S = 1 and S = N beside it) and the stage breakdown of the fully sharded search
N=${1:-8}
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02b_bench_n$N.json 2> $O/r02b_bench_n$N.err; tail -c 600 $O/r02b_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 scripts/profile_sharded.py > $O/r02b_stages_n$N.json 2> $O/r02b_stages_n$N.err; tail -c 400 $O/r02b_stages_n$N.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02b_bench_n$N.json"))
print("main", d["detail"]["parallelism"][:90], "value %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "kernel %.3f" % d["roofline"]["kernel_ms_avg"], d["parity_sample"]["ok"], d["clocks"])
for a in d["other_decompositions"]:
    print(" alt S=%d" % a["shards"], "value %.0f" % a["value"], "ms %.3f" % a["ms_per_step"], "e2e %.0f" % a["e2e"], "kernel %.3f" % a["kernel_ms_avg"], a["parity_sample"]["ok"])
s = json.loads(open("gpurun_out/r02b_stages_n$N.json").read().strip().splitlines()[-1])
print("stages rank0", json.dumps(s[0]))
print("stages rank%d" % (len(s) - 1), json.dumps(s[-1]))
PY
