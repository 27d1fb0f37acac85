#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kmeans_ivf.py tests/test_gpu_pipeline.py tests/test_gpu_small.py tests/test_gpu_reference_script.py -x -q -s > gpurun_out/r2_km_tests.log 2>&1; tail -15 gpurun_out/r2_km_tests.log
timeout 900 python scripts/bench_kmeans.py > gpurun_out/r2_kmeans.json 2> gpurun_out/r2_kmeans.err; tail -c 1500 gpurun_out/r2_kmeans.err; cat gpurun_out/r2_kmeans.json
timeout 600 python scripts/bench_small.py > gpurun_out/r2_small.json 2> gpurun_out/r2_small.err; tail -c 1500 gpurun_out/r2_small.err; cat gpurun_out/r2_small.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_km_launches.csv python scripts/bench_kmeans_ncu.py > gpurun_out/r2_km_ncu.log 2>&1; tail -3 gpurun_out/r2_km_ncu.log
