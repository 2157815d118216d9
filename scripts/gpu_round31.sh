#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu > gpurun_out/t31.log 2>&1; echo "== gpu tests: exit $?"; tail -n 25 gpurun_out/t31.log
timeout -k 5 300 python scripts/crf_sweep.py > gpurun_out/crf_sweep_r31.json 2> gpurun_out/crf_sweep.err; echo "sweep exit $?"
python - <<P
import json
d=json.load(open("gpurun_out/crf_sweep_r31.json"))
print("vit", [(r["B"], r["viterbi"]["ms"], r["viterbi"]["frac_hbm"]) for r in d["rows"]])
print("nll", [(r["B"], r["nll_fwd"]["ms"], r["nll_fwd"]["frac_hbm"]) for r in d["rows"]])
P
