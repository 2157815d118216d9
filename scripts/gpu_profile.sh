#!/bin/bash
# Profiling session on one GPU: attention alone, ncu launch lists of one inference step and one fine-tuning cycle, and
# ncu --set full of the attention forward + one layer's GEMMs out of the inference step.
#   gpurun --timeout 1500 -- bash scripts/gpu_profile.sh <tag>
set -u
TAG=${1:-prof}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 300 python scripts/attn_bench.py > $OUT/attn_bench.json 2>$OUT/attn.err; cat $OUT/attn_bench.json
timeout 300 python scripts/train_kernels_bench.py > $OUT/train_kernels.json 2>>$OUT/attn.err; tail -c 1500 $OUT/train_kernels.json; echo
# launch lists (graphs off so that ncu sees every kernel by name)
KBNER_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'attention_|crf_|embed_ln|gather_tagproj|gemm_|layernorm' -c 400 --csv --log-file $OUT/launches_infer.csv \
    python bench.py --steps 2 --warmup 3 --workload infer --no-cpu > $OUT/ncu_infer.log 2>&1; echo "ncu infer rc=$?"
KBNER_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_train.csv \
    python bench.py --steps 4 --warmup 3 --workload train > $OUT/ncu_train.log 2>&1; echo "ncu train rc=$?"
# full capture: the attention forward and the four GEMMs of a layer deep inside the second timed inference step
KBNER_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_fwd|gemm_' -s 260 -c 5 -o $OUT/layer_full -f \
    python bench.py --steps 2 --warmup 3 --workload infer --no-cpu > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
