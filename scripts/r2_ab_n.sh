#!/bin/bash
# usage: bash scripts/r2_ab_n.sh N ENVVAR [bench args] -- A/B of a kernel switch on the N-GPU bench
N=${1:-2}; V=${2:-NRB_NO_FIRST_SHOT}; shift; shift
O=gpurun_out; mkdir -p $O
for mode in new old new old; do
  if [ "$mode" = "old" ]; then export $V=1; else unset $V; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 "$@" > $O/abn_$mode.json 2> $O/abn_$mode.err || tail -c 800 $O/abn_$mode.err
  python - <<PY
import json
d = json.load(open("gpurun_out/abn_$mode.json"))
print("$mode main value %.0f" % d["value"], "ms %.3f" % d["ms_per_step"], "kernel %.3f" % d["roofline"]["kernel_ms_avg"], d["parity_sample"]["ok"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
for a in d["other_decompositions"]:
    print("   alt S=%d" % a["shards"], "value %.0f" % a["value"], "ms %.3f" % a["ms_per_step"], "kernel %.3f" % a["kernel_ms_avg"])
PY
done
