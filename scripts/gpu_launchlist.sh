#!/bin/bash
# The launch list the bench's roofline is cross-checked against: ncu --metrics gpu__time_duration.sum over the kernels of
# `bench.py --workload infer` (the library's kernels only; the graph's kernel nodes are profiled one by one).
#   gpurun --timeout 900 -- bash scripts/gpu_launchlist.sh <tag>
set -u
TAG=${1:-ll}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'attention_|crf_|embed_ln|gather_tagproj|gemm_|layernorm' -c 1400 --csv --log-file $OUT/launches_infer.csv \
    python bench.py --steps 2 --warmup 3 --workload infer --no-cpu > $OUT/ncu_infer.log 2>&1; echo "ncu infer rc=$?"
python scripts/summarize_launches.py $OUT/launches_infer.csv | head -20
