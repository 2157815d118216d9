#!/bin/bash
# residual/bias moved into LayerNorm, relaxed arrive + shuffled bias in the GEMM epilogue, dropout everywhere
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py -q -x -m gpu > gpurun_out/t13_kernels.log 2>&1; echo "kernel tests exit $?"; tail -n 8 gpurun_out/t13_kernels.log
timeout -k 5 600 python -m pytest tests/test_api_gpu.py tests/test_trainer_gpu.py -q -x -m gpu -s > gpurun_out/t13_api.log 2>&1; echo "api tests exit $?"; grep -v "^$" gpurun_out/t13_api.log | tail -n 14
SHAPES=qkv,attn_out,ffn_up,ffn_down,plain_f32 timeout -k 5 200 python scripts/gemm_bench.py > gpurun_out/gemm_shapes13.json 2>&1; cat gpurun_out/gemm_shapes13.json
timeout -k 5 400 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench exit $?"; cut -c1-1800 gpurun_out/bench13.json; tail -3 gpurun_out/bench13.err
timeout -k 5 400 python bench.py --workload train --steps 16 --warmup 4 > gpurun_out/bench_train13.json 2> gpurun_out/bench_train13.err; echo "train bench exit $?"; cut -c1-1500 gpurun_out/bench_train13.json; tail -3 gpurun_out/bench_train13.err
