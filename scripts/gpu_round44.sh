#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests -q -m gpu -x > gpurun_out/t44.log 2>&1; echo "== gpu tests: exit $?"; tail -n 3 gpurun_out/t44.log
timeout -k 5 600 python bench.py > gpurun_out/bench_r44_infer.json 2> gpurun_out/bench_infer.err; echo "== bench default: exit $?"; python - <<P
import json
d=json.load(open("gpurun_out/bench_r44_infer.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], d["cpu_baseline"]["value"])
P
timeout -k 5 600 python bench.py --workload train --no-cpu > gpurun_out/bench_r44_train.json 2> gpurun_out/bench_train.err; echo "== train: exit $?"; python - <<P
import json
d=json.load(open("gpurun_out/bench_r44_train.json")); print(d["value"], d["ms_per_step"], d["clocks"])
P
