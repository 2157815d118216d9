#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -q -m gpu -x > gpurun_out/t42.log 2>&1; echo "== gpu tests: exit $?"; tail -n 8 gpurun_out/t42.log
timeout -k 5 120 python scripts/attn_bench.py 2>&1 | tail -1
timeout -k 5 300 python bench.py --no-cpu --steps 40 > gpurun_out/bench_infer_r42.json 2> gpurun_out/bench_infer.err; echo "== infer bench: exit $?"; python - <<P
import json
d=json.load(open("gpurun_out/bench_infer_r42.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
P
