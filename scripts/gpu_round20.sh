#!/bin/bash
# refresh the ncu evidence for the final kernels: launch list of the bench command + full captures
mkdir -p gpurun_out
KBNER_GRAPHS=0 timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 260 --csv \
    --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench_final.log 2>&1
echo "launch list exit $?"
KBNER_GRAPHS=0 timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:"attention_fwd_kernel|gemm_ln_kernel|gemm_bf16_kernel" -s 240 -c 5 \
    -o gpurun_out/prof_final -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_final.log 2>&1
echo "capture exit $?"
tail -3 gpurun_out/ncu_final.log
