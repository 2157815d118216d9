#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python scripts/gemm_sweep.py > gpurun_out/gemm_sweep14.json 2> gpurun_out/gemm_sweep14.err; echo "sweep exit $?"; cat gpurun_out/gemm_sweep14.json; tail -3 gpurun_out/gemm_sweep14.err
