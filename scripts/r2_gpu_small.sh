#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_small.py tests/test_gpu_reference_script.py -x -q -s > gpurun_out/r2_small_tests.log 2>&1; tail -25 gpurun_out/r2_small_tests.log
timeout 600 python scripts/bench_small.py > gpurun_out/r2_small.json 2> gpurun_out/r2_small.err; tail -c 1500 gpurun_out/r2_small.err; cat gpurun_out/r2_small.json
