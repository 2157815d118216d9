#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py -x -q -m gpu -k "crf or viterbi or Viterbi or golden or small_end" > gpurun_out/t28.log 2>&1; echo "== crf tests: exit $?"; tail -n 5 gpurun_out/t28.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
for WM in 2048 1000000; do
KBNER_VIT_WIDE_MAX=$WM timeout -k 5 300 python scripts/crf_sweep.py > gpurun_out/crf_sweep_r28_wm$WM.json 2> gpurun_out/crf_sweep.err; echo "sweep wm=$WM exit $?"
python - <<P
import json
d=json.load(open("gpurun_out/crf_sweep_r28_wm$WM.json"))
print([(r["B"], r["viterbi"]["ms"], r["viterbi"]["frac_hbm"]) for r in d["rows"]])
P
done
L=29 timeout -k 5 300 python scripts/crf_sweep.py > gpurun_out/crf_sweep_r28_L29.json 2>> gpurun_out/crf_sweep.err
python - <<P
import json
d=json.load(open("gpurun_out/crf_sweep_r28_L29.json"))
print("L29", [(r["B"], r["viterbi"]["ms"], r["viterbi"]["frac_hbm"]) for r in d["rows"]])
P
SWEEP_B=4096 timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:"crf_viterbi" -s 6 -c 1 \
    -o gpurun_out/prof_vit5_b4096 -f python scripts/crf_sweep.py > gpurun_out/ncu_crf2.log 2>&1
echo "capture b4096 exit $?"
