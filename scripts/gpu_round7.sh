#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "crf" > gpurun_out/t7_crf.log 2>&1; echo "crf tests exit $?"; tail -n 8 gpurun_out/t7_crf.log
timeout -k 5 600 python -m pytest tests/test_api_gpu.py -q -m gpu -s > gpurun_out/t7_api.log 2>&1; echo "api tests exit $?"; tail -n 25 gpurun_out/t7_api.log
timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; echo "bench exit $?"; cat gpurun_out/bench_v7.json; tail -3 gpurun_out/bench_v7.err
timeout -k 5 300 python scripts/crf_sweep.py > gpurun_out/crf_sweep_v7.json 2> gpurun_out/crf_sweep_v7.err; cat gpurun_out/crf_sweep_v7.json
