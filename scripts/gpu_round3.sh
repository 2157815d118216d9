#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "gemm or attention" > gpurun_out/t3_gemm.log 2>&1; echo "gemm+attn tests exit $?"; tail -n 8 gpurun_out/t3_gemm.log
timeout 600 python -m pytest tests/test_train_kernels_gpu.py -q -m gpu > gpurun_out/t3_train.log 2>&1; echo "train kernels tests exit $?"; tail -n 40 gpurun_out/t3_train.log
timeout 200 python scripts/gemm_bench.py > gpurun_out/gemm_bench_v3.json 2>gpurun_out/gemm_bench_v3.err; cat gpurun_out/gemm_bench_v3.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; echo "bench exit $?"; cat gpurun_out/bench_v3.json; tail -3 gpurun_out/bench_v3.err
