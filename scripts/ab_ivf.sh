#!/bin/bash
# A/B of library builds on the IVF benchmark (and the flat one): bash scripts/ab_ivf.sh libA.so libB.so
for lib in "$@"; do
  NRB_LIB=$PWD/newsrecommend_b200/$lib timeout 300 python scripts/bench_ivf.py 2>> gpurun_out/ab_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', 'ivf search_s %.4f'%d['search_s'], 'kernels_ms %.2f'%d['scan_kernel_ms'], 'train_s %.3f'%d['train_s'], d['parity_sample'])"
done
