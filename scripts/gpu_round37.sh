#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -q -m gpu > gpurun_out/t37.log 2>&1; echo "== gpu tests: exit $?"; tail -n 12 gpurun_out/t37.log
timeout -k 5 300 python scripts/train_kernels_bench.py > gpurun_out/train_kernels_r37.json 2> gpurun_out/tk.err; cat gpurun_out/train_kernels_r37.json; tail -3 gpurun_out/tk.err
SWEEP_B=8,4096 timeout -k 5 300 python scripts/crf_sweep.py > gpurun_out/crf_sweep_r37.json 2>gpurun_out/crf_sweep.err; python - <<P
import json
d=json.load(open("gpurun_out/crf_sweep_r37.json"))
for r in d["rows"]: print(r["B"], {k:v["ms"] for k,v in r.items() if k!="B"})
P
timeout -k 5 600 python bench.py --workload train --no-cpu > gpurun_out/bench_train_r37.json 2> gpurun_out/bench_train.err; echo "== train bench: exit $?"; python - <<P
import json
d=json.load(open("gpurun_out/bench_train_r37.json")); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"], d["final_loss"])
P
