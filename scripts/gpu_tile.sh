#!/bin/bash
set -u
TAG=${1:-tile}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
for tn in 128 256; do
  KBNER_GEMM_TN=$tn timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py -m gpu -q -x -k "gemm" > $OUT/pytest_tn$tn.log 2>&1; echo "pytest tn=$tn rc=$?"; tail -3 $OUT/pytest_tn$tn.log
  KBNER_GEMM_TN=$tn timeout 200 python scripts/gemm_tile_bench.py > $OUT/tile_tn$tn.json 2> $OUT/tile_tn$tn.err; cat $OUT/tile_tn$tn.json; tail -2 $OUT/tile_tn$tn.err
done
