#!/bin/bash
# What the driver runs at round end, in one call: every -m gpu test (one pytest process, as the driver does), smoke(),
# the default bench line (with cpu_baseline) and the reference arm.
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -x -q -m gpu > gpurun_out/full_tests.log 2>&1; echo "== pytest -m gpu: exit $?"; tail -n 4 gpurun_out/full_tests.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/full_smoke.log 2>&1; echo "== smoke: exit $?"; tail -n 3 gpurun_out/full_smoke.log
timeout -k 5 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "== bench: exit $?"; cat gpurun_out/bench_default.json; tail -n 3 gpurun_out/bench_default.err
timeout -k 5 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_default.json 2> gpurun_out/bench_ref_default.err; echo "== ref: exit $?"; cat gpurun_out/bench_ref_default.json; tail -n 3 gpurun_out/bench_ref_default.err
