#!/bin/bash
# HEAD verification after the container re-creation: full gpu tests, smoke, bench lines, launch list, attention microbench,
# CRF sweep + full ncu capture of the CRF kernels at B=4096.
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -x -q -m gpu > gpurun_out/full_tests.log 2>&1; echo "== pytest -m gpu: exit $?"; tail -n 4 gpurun_out/full_tests.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/full_smoke.log 2>&1; echo "== smoke: exit $?"; tail -n 3 gpurun_out/full_smoke.log
timeout -k 5 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "== bench: exit $?"; cat gpurun_out/bench_default.json; tail -n 3 gpurun_out/bench_default.err
timeout -k 5 600 python bench.py --workload train --no-cpu > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "== train bench: exit $?"; cat gpurun_out/bench_train.json
timeout -k 5 300 python scripts/attn_bench.py > gpurun_out/attn_bench.json 2>&1; cat gpurun_out/attn_bench.json
timeout -k 5 300 python scripts/crf_sweep.py > gpurun_out/crf_sweep_r22.json 2> gpurun_out/crf_sweep.err; echo "sweep exit $?"; cat gpurun_out/crf_sweep_r22.json
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv \
    --log-file gpurun_out/launches_r22.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit $?"
SWEEP_B=4096 timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:"crf_" -s 6 -c 2 \
    -o gpurun_out/prof_crf_r22 -f python scripts/crf_sweep.py > gpurun_out/ncu_crf.log 2>&1
echo "crf capture exit $?"
ls -la gpurun_out
