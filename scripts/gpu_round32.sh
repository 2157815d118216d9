#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu > gpurun_out/t32.log 2>&1; echo "== gpu tests: exit $?"; tail -n 25 gpurun_out/t32.log
timeout -k 5 600 python bench.py --workload train --no-cpu > gpurun_out/bench_train_r32.json 2> gpurun_out/bench_train.err; echo "== train bench: exit $?"; python - <<P
import json
d=json.load(open("gpurun_out/bench_train_r32.json")); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
P
timeout -k 5 600 python bench.py --no-cpu > gpurun_out/bench_infer_r32.json 2> gpurun_out/bench_infer.err; echo "== infer bench: exit $?"; python - <<P
import json
d=json.load(open("gpurun_out/bench_infer_r32.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
P
