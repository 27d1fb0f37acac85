#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_kmeans_ivf.py tests/test_gpu_pipeline.py -x -q > gpurun_out/r2_fullsize_tests.log 2>&1; tail -8 gpurun_out/r2_fullsize_tests.log
bash scripts/r2_cfg5.sh 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_cfg5_launches.csv python scripts/bench_config5.py --nq 132608 --steps 1 --parity-queries 8 > gpurun_out/r2_cfg5_ncu.log 2>&1; tail -2 gpurun_out/r2_cfg5_ncu.log | cut -c1-300
make -C newsrecommend_b200/csrc trace > gpurun_out/r2_trace_build.log 2>&1
NRB_LIB=$PWD/newsrecommend_b200/libnrb200_trace.so timeout 600 python scripts/trace_ivf.py > gpurun_out/r2_trace_ivf.json 2> gpurun_out/r2_trace_ivf.err; tail -c 600 gpurun_out/r2_trace_ivf.err; cat gpurun_out/r2_trace_ivf.json
