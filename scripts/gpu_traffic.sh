#!/bin/bash
# Per-kernel time, DRAM bytes and tensor-pipe activity of the fine-tuning step and of the inference step (ncu, graphs off).
set -u
TAG=${1:-tr}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
KBNER_GRAPHS=0 timeout 900 ncu --metrics $M --clock-control none -k regex:'kbner|gemm|attention|layernorm|colsum|crf|adamw|embed|gather' -s 1200 -c 700 --csv --log-file $OUT/traffic_train.csv \
    python bench.py --steps 4 --warmup 3 --workload train > $OUT/ncu_train.log 2>&1; echo "ncu train rc=$?"
KBNER_GRAPHS=0 timeout 600 ncu --metrics $M --clock-control none -k regex:'attention_|crf_|embed_ln|gather_tagproj|gemm_|layernorm' -s 300 -c 130 --csv --log-file $OUT/traffic_infer.csv \
    python bench.py --steps 2 --warmup 3 --workload infer --no-cpu > $OUT/ncu_infer.log 2>&1; echo "ncu infer rc=$?"
ls -la $OUT
