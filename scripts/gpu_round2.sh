#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k attention > gpurun_out/test_attn2.log 2>&1; echo "attn2 tests exit $?"; tail -n 12 gpurun_out/test_attn2.log
KBNER_GEMM=1 timeout 200 python scripts/gemm_bench.py > gpurun_out/gemm_bench_v1.json 2>gpurun_out/gemm_bench_v1.err; cat gpurun_out/gemm_bench_v1.json
KBNER_GEMM=2 timeout 200 python scripts/gemm_bench.py > gpurun_out/gemm_bench_v2.json 2>gpurun_out/gemm_bench_v2.err; cat gpurun_out/gemm_bench_v2.json
KBNER_GEMM=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g2b.json 2> gpurun_out/bench_g2b.err; echo "bench exit $?"; cat gpurun_out/bench_g2b.json
KBNER_GEMM=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_tn -s 6 -c 5 -o gpurun_out/prof_gemm2 -f python scripts/gemm_bench.py > gpurun_out/ncu_gemm2.log 2>&1; echo "ncu exit $?"
KBNER_GEMM=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 30 -c 2 -o gpurun_out/prof_attn2 -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_attn2.log 2>&1; echo "ncu attn exit $?"
