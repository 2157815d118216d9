#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_kernels_gpu.py -q -m gpu -k attention_bwd > gpurun_out/t4_attnbwd.log 2>&1; echo "attn bwd tests exit $?"; tail -n 30 gpurun_out/t4_attnbwd.log
SHAPES=attn_out timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 3 -c 2 -o gpurun_out/prof_attnout -f python scripts/gemm_bench.py > gpurun_out/ncu_attnout.log 2>&1; echo "ncu exit $?"
