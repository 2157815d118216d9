#!/bin/bash
set -u
TAG=${1:-wg}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 600 python -m pytest tests/test_train_kernels_gpu.py -m gpu -q -x -k "wgrad" > $OUT/pytest_wgrad.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_wgrad.log
timeout 200 python scripts/wgrad_bench.py | tail -2 | tee $OUT/wgrad_tiles.json
KBNER_WGRAD_GROUP_MODE=stream timeout 200 python scripts/wgrad_bench.py | tail -1 | tee $OUT/wgrad_stream.json
timeout 600 python bench.py --steps 20 --warmup 5 --workload train > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "train rc=$?"
KBNER_WGRAD_GROUP=0 timeout 600 python bench.py --steps 20 --warmup 5 --workload train > $OUT/bench_train_nogroup.json 2> $OUT/bench_train_nogroup.err; echo "train(nogroup) rc=$?"
python - $OUT <<'PY'
import json,sys
for f in ("bench_train.json","bench_train_nogroup.json"):
    d=json.loads(open(sys.argv[1]+"/"+f).read().strip().splitlines()[-1]); t=d.get("train", d)
    print(f, t.get("value"), t.get("ms_per_step"), t.get("final_loss"))
PY
