#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu > gpurun_out/t33.log 2>&1; echo "== gpu tests: exit $?"; tail -n 6 gpurun_out/t33.log
timeout -k 5 300 python scripts/tagproj_bench.py 2>&1 | tail -2
timeout -k 5 300 python scripts/profile_host.py > gpurun_out/host_profile_r33.txt 2>&1; head -30 gpurun_out/host_profile_r33.txt
timeout -k 5 600 python bench.py --no-cpu --steps 40 > gpurun_out/bench_infer_r33.json 2> gpurun_out/bench_infer.err; echo "== infer bench: exit $?"; python - <<P
import json
d=json.load(open("gpurun_out/bench_infer_r33.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
P
