#!/bin/bash
# Fused GEMM + LayerNorm, grid vs cluster version: the kernel tests, the micro-bench, then the headline bench line.
#   gpurun --timeout 900 -- bash scripts/gpu_gemmln.sh <tag>
set -u
TAG=${1:-gl}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "layernorm_fused" > $OUT/pytest_gemmln.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/pytest_gemmln.log
timeout 200 python scripts/gemm_ln_bench.py > $OUT/gemm_ln_bench.json 2> $OUT/gemm_ln_bench.err; echo "bench rc=$?"; cat $OUT/gemm_ln_bench.json; tail -3 $OUT/gemm_ln_bench.err
timeout 400 python bench.py --steps 20 --warmup 5 --workload infer --no-cpu > $OUT/bench_infer.json 2> $OUT/bench_infer.err; echo "infer rc=$?"
tail -c 1800 $OUT/bench_infer.json; tail -3 $OUT/bench_infer.err
KBNER_GEMM_LN=cluster timeout 400 python bench.py --steps 20 --warmup 5 --workload infer --no-cpu > $OUT/bench_infer_cluster.json 2> $OUT/bench_infer_cluster.err; echo "infer(cluster) rc=$?"
python - $TAG <<'PY'
import json,sys
for f in ("bench_infer.json","bench_infer_cluster.json"):
    try:
        d=json.loads(open("gpurun_out/%s/%s" % (sys.argv[1] if len(sys.argv)>1 else "gl", f)).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("gemm_ms_per_step"), d.get("e2e",{}).get("value"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
