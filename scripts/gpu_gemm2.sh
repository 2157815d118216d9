#!/bin/bash
mkdir -p gpurun_out
KBNER_GEMM=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k gemm > gpurun_out/test_gemm2.log 2>&1; echo "gemm2 tests exit $?"
tail -n 15 gpurun_out/test_gemm2.log
KBNER_GEMM=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err; echo "bench g1 exit $?"
KBNER_GEMM=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "bench g2 exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_g1.json","gpurun_out/bench_g2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["achieved"], d["roofline"]["gemm_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
