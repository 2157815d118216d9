#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_train_kernels_gpu.py -q -x -m gpu > gpurun_out/t10_train.log 2>&1; echo "train kernel tests exit $?"; tail -n 12 gpurun_out/t10_train.log
timeout -k 5 600 python -m pytest tests/test_api_gpu.py -q -m gpu -s -k "finetune" > gpurun_out/t10_api.log 2>&1; echo "api finetune tests exit $?"; tail -n 8 gpurun_out/t10_api.log
timeout -k 5 400 python bench.py --workload train --steps 16 --warmup 4 > gpurun_out/bench_train10.json 2> gpurun_out/bench_train10.err; echo "train bench exit $?"; cat gpurun_out/bench_train10.json; tail -5 gpurun_out/bench_train10.err
