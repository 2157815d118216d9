#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" > gpurun_out/t45.log 2>&1; echo "== attention tests: exit $?"; tail -n 15 gpurun_out/t45.log
