#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 200 python scripts/attn_timeline.py > gpurun_out/attn_timeline_r41.json 2> gpurun_out/attn_timeline.err; echo "exit $?"; cat gpurun_out/attn_timeline_r41.json; tail -3 gpurun_out/attn_timeline.err
