#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -q -m gpu > gpurun_out/t34.log 2>&1; echo "== gpu tests: exit $?"; tail -n 12 gpurun_out/t34.log
SWEEP_B=8,32,256,4096 timeout -k 5 300 python scripts/crf_sweep.py > gpurun_out/crf_sweep_r34.json 2>gpurun_out/crf_sweep.err; tail -c 1500 gpurun_out/crf_sweep_r34.json; tail -3 gpurun_out/crf_sweep.err
timeout -k 5 300 python scripts/tagproj_bench.py 2>&1 | tail -2
timeout -k 5 600 python bench.py --workload train --no-cpu > gpurun_out/bench_train_r34.json 2> gpurun_out/bench_train.err; echo "== train bench: exit $?"; cat gpurun_out/bench_train_r34.json | cut -c1-400
