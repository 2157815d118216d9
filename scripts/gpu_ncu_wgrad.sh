#!/bin/bash
# ncu --set full of the weight-gradient GEMMs (grouped + one separate stream-K launch) at the fine-tuning shape.
set -u
TAG=${1:-nw}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_wgrad_group' -s 4 -c 2 -o $OUT/wgrad_full -f \
    python scripts/wgrad_bench.py > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/ncu.log
ls -la $OUT
