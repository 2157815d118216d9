#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_train_kernels_gpu.py -q -x -m gpu -k attention_bwd > gpurun_out/t11_attnbwd.log 2>&1; echo "attn bwd tests exit $?"; tail -n 6 gpurun_out/t11_attnbwd.log
timeout -k 5 300 python -m pytest tests/test_api_gpu.py -q -m gpu -s -k "finetune" > gpurun_out/t11_api.log 2>&1; echo "api finetune tests exit $?"; tail -n 6 gpurun_out/t11_api.log
timeout -k 5 400 python bench.py --workload train --steps 16 --warmup 4 > gpurun_out/bench_train11.json 2> gpurun_out/bench_train11.err; echo "train bench exit $?"; cat gpurun_out/bench_train11.json | cut -c1-400; tail -5 gpurun_out/bench_train11.err
