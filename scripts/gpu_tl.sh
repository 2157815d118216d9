#!/bin/bash
set -u
TAG=${1:-tl}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 200 python scripts/gemm_ln_timeline.py > $OUT/timeline.json 2> $OUT/timeline.err; echo "timeline rc=$?"; cat $OUT/timeline.json; tail -3 $OUT/timeline.err
