#!/bin/bash
mkdir -p gpurun_out
KBNER_GRAPHS=0 timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/launches_train.csv python bench.py --workload train --steps 4 --warmup 4 > gpurun_out/ncu_train.log 2>&1; echo "train launch list exit $?"
tail -3 gpurun_out/ncu_train.log
