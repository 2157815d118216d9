#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python scripts/gemm_ln_bench.py > gpurun_out/gemm_ln17.json 2> gpurun_out/gemm_ln17.err; echo "exit $?"; cat gpurun_out/gemm_ln17.json; tail -3 gpurun_out/gemm_ln17.err
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:gemm_ln_kernel -s 3 -c 1 -o gpurun_out/prof_gemm_ln -f python scripts/gemm_ln_bench.py > gpurun_out/ncu_gemm_ln.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_gemm_ln.log
