#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "layernorm_fused" > gpurun_out/t18_fused.log 2>&1; echo "fused kernel tests exit $?"; tail -n 12 gpurun_out/t18_fused.log
timeout -k 5 300 python scripts/gemm_ln_bench.py > gpurun_out/gemm_ln18.json 2> gpurun_out/gemm_ln18.err; echo "exit $?"; cat gpurun_out/gemm_ln18.json; tail -3 gpurun_out/gemm_ln18.err
timeout -k 5 300 python -m pytest tests/test_api_gpu.py -q -x -m gpu -s -k "large_parity or small_end or remove_x or long_sentence" > gpurun_out/t18_api.log 2>&1; echo "api tests exit $?"; grep -v "^$" gpurun_out/t18_api.log | tail -n 6
KBNER_FUSE_LN=1 timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench18.json 2> gpurun_out/bench18.err; echo "bench exit $?"; python - <<P
import json
d=json.load(open("gpurun_out/bench18.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["gemm_ms_per_step"], d["clocks"])
P
tail -2 gpurun_out/bench18.err
