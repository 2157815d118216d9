#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_api_gpu.py -q -x -m gpu -s > gpurun_out/t6_api.log 2>&1; echo "api tests exit $?"; tail -n 25 gpurun_out/t6_api.log
timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; echo "bench exit $?"; cat gpurun_out/bench_v6.json; tail -3 gpurun_out/bench_v6.err
timeout -k 5 300 python scripts/profile_host.py > gpurun_out/host_profile.txt 2>&1; echo "profile exit $?"; head -60 gpurun_out/host_profile.txt
timeout -k 5 400 python bench.py --workload train --steps 8 --warmup 4 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train bench exit $?"; cat gpurun_out/bench_train.json; tail -5 gpurun_out/bench_train.err
