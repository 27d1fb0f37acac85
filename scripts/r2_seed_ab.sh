#!/bin/bash
# usage: bash scripts/r2_seed_ab.sh N [tests] -- sample-pass seeding of the sharded search, on / off, N GPUs
N=${1:-2}
O=gpurun_out; mkdir -p $O
if [ "${2:-}" = "tests" ]; then
  timeout 900 python -m pytest tests/test_gpu_flat.py tests/test_gpu_sharded.py -x -q > $O/seed_tests.log 2>&1; tail -4 $O/seed_tests.log
fi
for rows in 8192 0 2048 8192 0; do
  NRB_SEED_ROWS=$rows timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-alt > $O/seed_n${N}_$rows.json 2> $O/seed_n${N}_$rows.err || tail -c 1500 $O/seed_n${N}_$rows.err
  python -c "
import json
d=json.load(open('gpurun_out/seed_n${N}_$rows.json')); r=d['roofline']; print('seed_rows $rows: q/s %.0f'%d['value'], 'ms %.3f'%d['ms_per_step'], 'kernel_ms %.3f'%r['kernel_ms_avg'], 'launches', r['kernel_launches_timed'], 'e2e %.0f'%d['e2e']['value'], d['parity_sample']['ok'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
