#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_api_gpu.py -q -m gpu -s > gpurun_out/t8_api.log 2>&1; echo "api tests exit $?"; tail -n 25 gpurun_out/t8_api.log
timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; echo "bench exit $?"; cat gpurun_out/bench_v8.json; tail -3 gpurun_out/bench_v8.err
timeout -k 5 400 python bench.py --workload train --steps 16 --warmup 4 > gpurun_out/bench_train8.json 2> gpurun_out/bench_train8.err; echo "train bench exit $?"; cat gpurun_out/bench_train8.json; tail -5 gpurun_out/bench_train8.err
SWEEP_B=64 timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:crf_viterbi -s 3 -c 1 -o gpurun_out/prof_viterbi64 -f python scripts/crf_sweep.py > gpurun_out/ncu_vit.log 2>&1; echo "ncu viterbi exit $?"
