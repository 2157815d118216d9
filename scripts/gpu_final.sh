#!/bin/bash
# Round-end evidence pass: tests, smoke, the three bench arms, sweeps / micro-benches, ncu launch list + full captures.
R=${ROUND:-r43}
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests_$R.txt 2>&1; echo "== gpu tests: exit $?"; tail -n 4 gpurun_out/gpu_tests_$R.txt
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -k 5 900 python bench.py > gpurun_out/bench_${R}_infer.json 2> gpurun_out/bench_infer.err; echo "== bench default: exit $?"; cut -c1-1500 gpurun_out/bench_${R}_infer.json
timeout -k 5 900 python bench.py --impl reference > gpurun_out/bench_${R}_reference_arm.json 2> gpurun_out/bench_ref.err; echo "== bench reference arm: exit $?"; cut -c1-700 gpurun_out/bench_${R}_reference_arm.json
timeout -k 5 600 python bench.py --workload train --no-cpu > gpurun_out/bench_${R}_train.json 2> gpurun_out/bench_train.err; echo "== bench train: exit $?"; cut -c1-300 gpurun_out/bench_${R}_train.json
timeout -k 5 300 python scripts/crf_sweep.py > gpurun_out/crf_sweep_$R.json 2> gpurun_out/crf_sweep.err; echo "== crf sweep: exit $?"
L=29 SWEEP_B=64,4096 timeout -k 5 300 python scripts/crf_sweep.py > gpurun_out/crf_sweep_${R}_L29.json 2>> gpurun_out/crf_sweep.err
timeout -k 5 300 python scripts/train_kernels_bench.py > gpurun_out/train_kernels_$R.json 2> gpurun_out/tk.err; cat gpurun_out/train_kernels_$R.json
timeout -k 5 300 python scripts/attn_bench.py > gpurun_out/attn_bench_$R.json 2> gpurun_out/ab.err; cat gpurun_out/attn_bench_$R.json
NCU="ncu --set full --clock-control none --import-source on -f"
timeout -k 5 300 $NCU -k regex:attention_fwd_kernel -c 1 -o gpurun_out/prof_attn_fwd_$R python scripts/attn_bench.py > gpurun_out/ncu_attn_fwd.log 2>&1; echo "attn fwd capture exit $?"
timeout -k 5 300 $NCU -k regex:crf_ -c 6 -o gpurun_out/prof_crf_$R python scripts/crf_once.py > gpurun_out/ncu_crf.log 2>&1; echo "crf capture exit $?"
REPS=1 WARM=0 timeout -k 5 300 $NCU -k regex:"layernorm_bwd|gemm_bf16_kernel|gather_tagproj_bwd|colsum" -c 7 -o gpurun_out/prof_trainmisc_$R python scripts/train_kernels_bench.py > gpurun_out/ncu_trainmisc.log 2>&1; echo "train misc capture exit $?"
timeout -k 5 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 500 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "launch list exit $?"
ls -la gpurun_out | tail -30
