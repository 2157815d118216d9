#!/bin/bash
# Runs the GPU parity tests group by group (separate processes: a trapped kernel poisons only its group).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for grp in crf gemm "layernorm or embed or tagproj" attention; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "$grp" > "gpurun_out/test_${name}.log" 2>&1
  echo "== $grp: exit $?" | tee -a gpurun_out/summary.txt
  tail -n 25 "gpurun_out/test_${name}.log"
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "== smoke: exit $?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/smoke.log
timeout 900 python -m pytest tests/test_api_gpu.py -q -x -m gpu -s > gpurun_out/test_api.log 2>&1; echo "== api: exit $?" | tee -a gpurun_out/summary.txt
tail -n 30 gpurun_out/test_api.log
