#!/bin/bash
# Attention-backward timeline (debug build of the library on the box only).
set -u
TAG=${1:-abtl}; OUT=gpurun_out/$TAG; mkdir -p $OUT
KBNER_EXTRA_NVCC_FLAGS=-DKBNER_ATTN_BWD_DEBUG python kb-ner_b200/csrc/build.py --force > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
for d in 1 0; do
  DROP=$d timeout 200 python scripts/attn_bwd_timeline.py > $OUT/timeline_drop$d.json 2> $OUT/timeline.err; echo "rc=$?"; cat $OUT/timeline_drop$d.json; tail -3 $OUT/timeline.err
done
timeout 200 python scripts/attn_bench.py; R=8 timeout 100 python scripts/attn_bench.py
