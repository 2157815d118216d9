#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -q -m gpu -x > gpurun_out/t40.log 2>&1; echo "== gpu tests: exit $?"; tail -n 6 gpurun_out/t40.log
timeout -k 5 300 python scripts/attn_bench.py 2>&1 | tail -1
timeout -k 5 600 python bench.py --no-cpu --steps 40 > gpurun_out/bench_infer_r40.json 2> gpurun_out/bench_infer.err; echo "== infer bench: exit $?"; python - <<P
import json
d=json.load(open("gpurun_out/bench_infer_r40.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
P
