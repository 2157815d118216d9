#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "attention" > gpurun_out/t5_attn.log 2>&1; echo "attn fwd tests exit $?"; tail -n 6 gpurun_out/t5_attn.log
timeout -k 5 300 python -m pytest tests/test_train_kernels_gpu.py -q -x -m gpu -k attention_bwd > gpurun_out/t5_attnbwd.log 2>&1; echo "attn bwd tests exit $?"; tail -n 15 gpurun_out/t5_attnbwd.log
timeout -k 5 600 python -m pytest tests/test_api_gpu.py -q -x -m gpu -s > gpurun_out/t5_api.log 2>&1; echo "api tests exit $?"; tail -n 30 gpurun_out/t5_api.log
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; echo "bench exit $?"; cat gpurun_out/bench_v5.json; tail -3 gpurun_out/bench_v5.err
