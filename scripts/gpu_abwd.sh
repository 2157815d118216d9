#!/bin/bash
# Attention backward: tests (release build), bench, then the timeline with the debug build.
set -u
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 600 python -m pytest tests/test_train_kernels_gpu.py tests/test_kernels_gpu.py -m gpu -q -x -k "attention" > $OUT/pytest_attn.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_attn.log
timeout 200 python scripts/attn_bench.py; R=8 timeout 100 python scripts/attn_bench.py
timeout 600 python bench.py --steps 20 --warmup 5 --workload train > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "train rc=$?"
python - $OUT/bench_train.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
t=d.get("train", d)
print("train", t.get("value"), t.get("ms_per_step"), t.get("final_loss"))
PY
KBNER_EXTRA_NVCC_FLAGS=-DKBNER_ATTN_BWD_DEBUG python kb-ner_b200/csrc/build.py --force > $OUT/build_dbg.log 2>&1 || { tail -30 $OUT/build_dbg.log; exit 1; }
DROP=1 timeout 200 python scripts/attn_bwd_timeline.py > $OUT/timeline_drop1.json 2> $OUT/timeline.err; echo "rc=$?"; cat $OUT/timeline_drop1.json; tail -3 $OUT/timeline.err
