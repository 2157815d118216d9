#!/bin/bash
# compute-sanitizer over the kernel-level parity tests (full-size cases excluded: the tools slow kernels 10-100x).
#   gpurun --timeout 1800 -- bash scripts/gpu_sanitize.sh <tag>
set -u
TAG=${1:-san}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
SEL='not full_size and not linearity'
for tool in memcheck racecheck synccheck; do
  timeout 800 compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 7 \
      python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py tests/test_precision_gpu.py -m gpu -q -x -k "$SEL and not configs2 and not precision_modes and not rounding_floor" \
      > $OUT/$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|hazard" $OUT/$tool.log | tail -5
done
