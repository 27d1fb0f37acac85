#!/bin/bash
# A/B of a kernel change on one GPU: parity tests, then bench.py and the IVF bench with and without it.
# usage: bash scripts/r2_ab.sh ENVVAR   (ENVVAR=1 disables the change, e.g. NRB_NO_PAIR_PRUNE)
V=${1:-NRB_NO_PAIR_PRUNE}
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_flat.py tests/test_gpu_kmeans_ivf.py tests/test_gpu_fullsize.py tests/test_gpu_small.py tests/test_gpu_sharded.py -x -q > $O/ab_tests.log 2>&1; tail -4 $O/ab_tests.log
for mode in new old new old; do
  if [ "$mode" = "old" ]; then export $V=1; else unset $V; fi
  timeout 300 python bench.py --steps 20 --warmup 5 > $O/ab_bench_$mode.json 2>> $O/ab.err; python -c "
import json
d=json.load(open('gpurun_out/ab_bench_$mode.json')); r=d['roofline']; print('$mode flat q/s %.0f'%d['value'], 'kernel_ms %.3f'%r['kernel_ms_avg'], 'frac %.3f'%r['frac'], 'e2e %.0f'%d['e2e']['value'], d['parity_sample']['ok'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  timeout 300 python scripts/bench_ivf.py > $O/ab_ivf_$mode.json 2>> $O/ab.err; python -c "
import json
d=json.loads(open('gpurun_out/ab_ivf_$mode.json').read().strip().splitlines()[-1]); print('$mode ivf search_s %.4f'%d['search_s'], 'qps %.0f'%d['search_qps'], 'kernels_ms %.2f'%d['scan_kernel_ms'], d['parity_sample'])"
done
tail -5 $O/ab.err
