#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py -q -x -m gpu -k "gemm or attention or large_parity or small_end or finetune_grad" > gpurun_out/t19.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/t19.log
for f in 1 0 1; do
KBNER_PDL=$f timeout -k 5 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench19_p$f.json 2> gpurun_out/bench19_p$f.err; echo "pdl=$f exit $?"; python - <<P
import json
d=json.load(open("gpurun_out/bench19_p$f.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"])
P
tail -2 gpurun_out/bench19_p$f.err
done
KBNER_PDL=1 timeout -k 5 400 python bench.py --workload train --steps 16 --warmup 4 > gpurun_out/bench_train19.json 2> gpurun_out/bench_train19.err; echo "train bench exit $?"; cut -c1-420 gpurun_out/bench_train19.json; tail -3 gpurun_out/bench_train19.err
