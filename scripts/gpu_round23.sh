#!/bin/bash
# Viterbi v2: parity tests, smoke, sweeps (L=13 and L=29)
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py -x -q -m gpu -k "crf or viterbi or Viterbi or golden or small_end" > gpurun_out/t23.log 2>&1; echo "== crf tests: exit $?"; tail -n 15 gpurun_out/t23.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout -k 5 300 python scripts/crf_sweep.py > gpurun_out/crf_sweep_r23.json 2> gpurun_out/crf_sweep.err; echo "sweep exit $?"; cat gpurun_out/crf_sweep_r23.json; tail -3 gpurun_out/crf_sweep.err
L=29 timeout -k 5 300 python scripts/crf_sweep.py > gpurun_out/crf_sweep_r23_L29.json 2>> gpurun_out/crf_sweep.err; echo "sweep exit $?"; cat gpurun_out/crf_sweep_r23_L29.json
