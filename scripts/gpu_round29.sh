#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py tests/test_train_kernels_gpu.py -x -q -m gpu -k "crf or viterbi or Viterbi or golden or small_end or nll or loss or finetune" > gpurun_out/t30.log 2>&1; echo "== crf tests: exit $?"; tail -n 8 gpurun_out/t30.log
timeout -k 5 300 python scripts/crf_sweep.py > gpurun_out/crf_sweep_r30.json 2> gpurun_out/crf_sweep.err; echo "sweep exit $?"
python - <<P
import json
d=json.load(open("gpurun_out/crf_sweep_r30.json"))
print("vit", [(r["B"], r["viterbi"]["ms"], r["viterbi"]["frac_hbm"]) for r in d["rows"]])
print("nll", [(r["B"], r["nll_fwd"]["ms"], r["nll_fwd"]["frac_hbm"]) for r in d["rows"]])
P
L=29 timeout -k 5 300 python scripts/crf_sweep.py > gpurun_out/crf_sweep_r30_L29.json 2>> gpurun_out/crf_sweep.err
python - <<P
import json
d=json.load(open("gpurun_out/crf_sweep_r30_L29.json"))
print("L29 vit", [(r["B"], r["viterbi"]["ms"], r["viterbi"]["frac_hbm"]) for r in d["rows"]])
print("L29 nll", [(r["B"], r["nll_fwd"]["ms"], r["nll_fwd"]["frac_hbm"]) for r in d["rows"]])
P
