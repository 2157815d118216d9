#!/bin/bash
set -u
TAG=${1:-wg}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py -m gpu -q -x -k "gemm or wgrad" > $OUT/pytest_gemm.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gemm.log
KBNER_GEMM_STREAMK=0 timeout 200 python scripts/wgrad_bench.py | tee $OUT/wgrad_splitk.json
timeout 200 python scripts/wgrad_bench.py | tee $OUT/wgrad_streamk.json
KBNER_WGRAD_GROUP=0 timeout 600 python bench.py --steps 20 --warmup 5 --workload train > $OUT/bench_train_nogroup.json 2> $OUT/bench_train_nogroup.err; echo "train(nogroup) rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --workload train > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "train rc=$?"
KBNER_GEMM_STREAMK=0 timeout 600 python bench.py --steps 20 --warmup 5 --workload train > $OUT/bench_train_splitk.json 2> $OUT/bench_train_splitk.err; echo "train rc=$?"
python - $OUT <<'PY'
import json,sys
for f in ("bench_train.json","bench_train_nogroup.json","bench_train_splitk.json"):
    d=json.loads(open(sys.argv[1]+"/"+f).read().strip().splitlines()[-1]); t=d.get("train", d)
    print(f, t.get("value"), t.get("ms_per_step"), t.get("final_loss"))
PY
