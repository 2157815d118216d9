#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "layernorm_fused" > gpurun_out/t16_fused.log 2>&1; echo "fused kernel tests exit $?"; tail -n 12 gpurun_out/t16_fused.log
timeout -k 5 300 python -m pytest tests/test_api_gpu.py -q -x -m gpu -s -k "large_parity or small_end or remove_x or long_sentence" > gpurun_out/t16_api.log 2>&1; echo "api tests exit $?"; grep -v "^$" gpurun_out/t16_api.log | tail -n 6
for f in 1 0; do
KBNER_FUSE_LN=$f timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench16_f$f.json 2> gpurun_out/bench16_f$f.err; echo "fuse=$f exit $?"; python - <<P
import json
d=json.load(open("gpurun_out/bench16_f$f.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"], d["clocks"])
P
tail -2 gpurun_out/bench16_f$f.err
done
