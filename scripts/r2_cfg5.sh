#!/bin/bash
# usage: bash scripts/r2_cfg5.sh N   (on a box with N GPUs)
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1200 python scripts/bench_config5.py > gpurun_out/r2_cfg5_n1.json 2> gpurun_out/r2_cfg5_n1.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/bench_config5.py > gpurun_out/r2_cfg5_n$N.json 2> gpurun_out/r2_cfg5_n$N.err
fi
tail -c 1200 gpurun_out/r2_cfg5_n$N.err; cat gpurun_out/r2_cfg5_n$N.json
