#!/bin/bash
# One GPU session: the GPU test suite, the bench line, and the ncu launch list of one bench step.
#   gpurun --timeout 1500 -- bash scripts/gpu_ci.sh <tag> [pytest-args...]
# Everything lands under gpurun_out/<tag>/ (copied into profiles/ by hand when it is evidence).
set -u
TAG=${1:-ci}; shift || true
OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 1200 python -m pytest tests -m gpu -q -s "$@" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -5 $OUT/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
