#!/bin/bash
# compute-sanitizer over the tests of the kernels changed late in round 2 (fused GEMM+LayerNorm grid kernel, grouped weight
# gradients, attention backward, GELU epilogues); the 16384-row cases are left out (the tools slow kernels 10-100x).
set -u
TAG=${1:-san}; shift || true
TOOLS=${*:-memcheck synccheck}
OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
SEL='(layernorm_fused and not 16384) or wgrad or attention_bwd or with_dropout or (gemm and gelu)'
SEL="($SEL) and not dgelu"
for tool in $TOOLS; do
  timeout 700 compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 7 \
      python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py -m gpu -q -k "$SEL" > $OUT/$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $OUT/$tool.log | tail -4
done
