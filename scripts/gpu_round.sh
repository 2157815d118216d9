#!/bin/bash
# One GPU session for a kernel change: the kernel tests, the GEMM micro-benches, then the bench legs asked for.
#   gpurun --timeout 1200 -- bash scripts/gpu_round.sh <tag> [infer|train|all|none]
set -u
TAG=${1:-rd}; LEG=${2:-infer}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py -m gpu -q -x > $OUT/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -4 $OUT/pytest_kernels.log
timeout 200 python scripts/gemm_tile_bench.py > $OUT/tile.json 2> $OUT/tile.err; cat $OUT/tile.json; tail -2 $OUT/tile.err
timeout 200 python scripts/gemm_ln_bench.py 2> $OUT/gemm_ln_bench.err | head -3 > $OUT/gemm_ln_bench.json; cat $OUT/gemm_ln_bench.json
timeout 300 python scripts/train_kernels_bench.py > $OUT/train_kernels.json 2> $OUT/train_kernels.err; tail -c 2500 $OUT/train_kernels.json; echo
if [ "$LEG" != "none" ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 --workload $LEG --no-cpu > $OUT/bench_$LEG.json 2> $OUT/bench_$LEG.err; echo "bench $LEG rc=$?"
  python - $OUT/bench_$LEG.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d.get("roofline",{})
print("value", d.get("value"), "ms", d.get("ms_per_step"), "frac", r.get("frac"), "gemm_ms", r.get("gemm_ms_per_step"), "e2e", d.get("e2e",{}).get("value"), "clocks", d.get("clocks"))
t=d.get("train")
if t: print("train", {k:t[k] for k in t if k in ("value","ms_per_step","unit")})
p=d.get("parity")
if p: print("parity", json.dumps(p)[:600])
PY
fi
