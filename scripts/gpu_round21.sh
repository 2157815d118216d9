#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py -q -x -m gpu -k "attention" > gpurun_out/t21_attn.log 2>&1; echo "attention tests exit $?"; tail -n 5 gpurun_out/t21_attn.log
timeout -k 5 300 python -m pytest tests/test_api_gpu.py -q -x -m gpu -s -k "large_parity or small_end or finetune" > gpurun_out/t21_api.log 2>&1; echo "api tests exit $?"; grep -v "^$" gpurun_out/t21_api.log | tail -n 8
timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench21.json 2> gpurun_out/bench21.err; echo "bench exit $?"; python - <<P
import json
d=json.load(open("gpurun_out/bench21.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"])
P
tail -2 gpurun_out/bench21.err
