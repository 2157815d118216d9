#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python scripts/e2e_gaps.py > gpurun_out/e2e_gaps_r35.json 2> gpurun_out/e2e_gaps.err; echo "exit $?"; cat gpurun_out/e2e_gaps_r35.json | tr '\n' ' ' | cut -c1-3500; tail -3 gpurun_out/e2e_gaps.err
