#!/bin/bash
# trainer test + ncu capture of the attention-out GEMM (RESID_F32 epilogue, short K)
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_trainer_gpu.py -q -x -m gpu -s > gpurun_out/t12_trainer.log 2>&1; echo "trainer test exit $?"; tail -n 8 gpurun_out/t12_trainer.log
SHAPES=attn_out,plain_f32 timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 3 -c 1 \
    -o gpurun_out/prof_gemm_attnout -f python scripts/gemm_bench.py > gpurun_out/ncu_attnout.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_attnout.log
SHAPES=attn_out,plain_f32,ffn_down timeout -k 5 200 python scripts/gemm_bench.py
