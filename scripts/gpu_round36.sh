#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches_train_r36.csv python bench.py --workload train --steps 4 --warmup 4 --no-cpu > gpurun_out/ncu_train.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches_train_r36.csv; tail -2 gpurun_out/ncu_train.log
