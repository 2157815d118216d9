#!/bin/bash
# compute-sanitizer over the kernel-level parity tests (full-size cases excluded: the tools slow kernels 10-100x).
#   gpurun --timeout 1800 -- bash scripts/gpu_sanitize.sh <tag> [tools...]        default tools: memcheck synccheck racecheck
set -u
TAG=${1:-san}; shift || true
TOOLS=${*:-memcheck synccheck racecheck}
OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
SEL=${KBNER_SAN_SEL:-'not full_size and not linearity and not configs2 and not precision_modes and not rounding_floor'}
# late round-2 kernels only (profiles/r02/sanitizer_san5.txt):
#   KBNER_SAN_SEL='((layernorm_fused and not 16384) or wgrad or attention_bwd or with_dropout or (gemm and gelu)) and not dgelu'
for tool in $TOOLS; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 7 \
      python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py tests/test_precision_gpu.py -m gpu -q -k "$SEL" \
      > $OUT/$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $OUT/$tool.log | tail -4
done
