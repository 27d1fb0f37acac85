#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_flat.py tests/test_gpu_kmeans_ivf.py tests/test_gpu_fullsize.py tests/test_gpu_small.py -x -q > gpurun_out/r2_os_tests.log 2>&1; tail -4 gpurun_out/r2_os_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_os_bench.json 2>> gpurun_out/r2_os.err; python -c "
import json
d=json.load(open('gpurun_out/r2_os_bench.json')); r=d['roofline']; print('flat q/s %.0f'%d['value'], 'kernel_ms %.3f'%r['kernel_ms_avg'], 'frac %.3f'%r['frac'], 'e2e %.0f'%d['e2e']['value'], d['parity_sample']['ok'], d['clocks'])"
timeout 300 python scripts/bench_ivf.py > gpurun_out/r2_ivf_new.json 2>> gpurun_out/r2_os.err; python -c "
import json
d=json.loads(open('gpurun_out/r2_ivf_new.json').read().strip().splitlines()[-1]); print('ivf search_s %.4f'%d['search_s'], 'qps %.0f'%d['search_qps'], 'kernels_ms %.2f'%d['scan_kernel_ms'], 'train_s %.4f'%d['train_s'], 'add_s %.4f'%d['add_s'], d['parity_sample'])"
timeout 300 python scripts/bench_kmeans.py > gpurun_out/r2_kmeans.json 2>> gpurun_out/r2_os.err; python -c "
import json
d=json.loads(open('gpurun_out/r2_kmeans.json').read().strip().splitlines()[-1])
for r in d['train']: print(r['metric'], r['niter'], 'gpu_s %.4f'%r['gpu_seconds'], 'ms/iter %.3f'%r['gpu_ms_per_iteration'], 'obj_last', r['obj_last'], r.get('cpu_obj_last'), r.get('speedup'))
print(d['k1b_update'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_ivf_launches.csv python scripts/ivf_ncu_target.py > gpurun_out/r2_ivf_ncu.log 2>&1; tail -2 gpurun_out/r2_ivf_ncu.log
tail -5 gpurun_out/r2_os.err
