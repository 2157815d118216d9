#!/bin/bash
mkdir -p gpurun_out
SWEEP_B=64,4096 timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:"crf_viterbi" -s 6 -c 1 \
    -o gpurun_out/prof_vit2_b64 -f python scripts/crf_sweep.py > gpurun_out/ncu_crf.log 2>&1
echo "capture b64 exit $?"
SWEEP_B=4096 timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:"crf_viterbi" -s 6 -c 1 \
    -o gpurun_out/prof_vit2_b4096 -f python scripts/crf_sweep.py > gpurun_out/ncu_crf2.log 2>&1
echo "capture b4096 exit $?"
