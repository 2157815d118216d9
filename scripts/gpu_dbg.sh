#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/dbg_cases.log
for c in "1 256 1 256" "1 256 1 200" "1 200 1 200" "1 130 1 130"; do
  echo "== case $c" >> gpurun_out/dbg_cases.log
  timeout -k 3 25 python -u scripts/dbg_attn_bwd.py $c >> gpurun_out/dbg_cases.log 2>&1
  echo "exit $?" >> gpurun_out/dbg_cases.log
done
cat gpurun_out/dbg_cases.log
