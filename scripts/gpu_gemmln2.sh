#!/bin/bash
set -u
TAG=${1:-gl}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "layernorm_fused" > $OUT/pytest_gemmln.log 2>&1; echo "pytest rc=$?"
tail -8 $OUT/pytest_gemmln.log
timeout 200 python scripts/gemm_ln_timeline.py > $OUT/timeline.json 2> $OUT/timeline.err; echo "timeline rc=$?"; cat $OUT/timeline.json; tail -3 $OUT/timeline.err
timeout 200 python scripts/gemm_ln_bench.py 2> $OUT/bench.err | head -3 > $OUT/gemm_ln_bench.json; cat $OUT/gemm_ln_bench.json
timeout 400 python bench.py --steps 20 --warmup 5 --workload infer --no-cpu > $OUT/bench_infer.json 2> $OUT/bench_infer.err; echo "infer rc=$?"
python - $TAG <<'PY'
import json,sys
d=json.loads(open("gpurun_out/%s/bench_infer.json" % sys.argv[1]).read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["gemm_ms_per_step"], d["e2e"]["value"], d["clocks"])
PY
