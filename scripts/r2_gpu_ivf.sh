#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kmeans_ivf.py tests/test_gpu_fullsize.py tests/test_gpu_pipeline.py -x -q > gpurun_out/r2_ivf_tests.log 2>&1; tail -6 gpurun_out/r2_ivf_tests.log
for mode in 0 1 default; do
  if [ "$mode" = "default" ]; then unset NRB_IVF_MODE; else export NRB_IVF_MODE=$mode; fi
  timeout 300 python scripts/bench_ivf.py 2>> gpurun_out/r2_ivf_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mode $mode', 'search_s %.4f'%d['search_s'], 'qps %.0f'%d['search_qps'], 'kernels_ms %.2f'%d['scan_kernel_ms'], d['parity_sample'])"
done
unset NRB_IVF_MODE
NRB_IVF_BATCH=262144 timeout 300 python scripts/bench_ivf.py > gpurun_out/r2_ivf_b262k.json 2>> gpurun_out/r2_ivf_ab.err; python -c "
import json
d=json.loads(open('gpurun_out/r2_ivf_b262k.json').read().strip().splitlines()[-1]); print('batch262k', 'search_s %.4f'%d['search_s'], 'qps %.0f'%d['search_qps'], 'kernels_ms %.2f'%d['scan_kernel_ms'], d['parity_sample'])"
tail -5 gpurun_out/r2_ivf_ab.err
