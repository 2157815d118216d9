#!/bin/bash
mkdir -p gpurun_out
for c in 1 2 4; do
KBNER_MCHUNKS=$c timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench15_c$c.json 2> gpurun_out/bench15_c$c.err; echo "chunks=$c exit $?"; python - <<P
import json
d=json.load(open("gpurun_out/bench15_c$c.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["gemm_ms_per_step"], d["clocks"])
P
done
KBNER_MCHUNKS=2 timeout -k 5 300 python -m pytest tests/test_api_gpu.py -q -x -m gpu -k "large_parity or small_end" 2>&1 | tail -3
