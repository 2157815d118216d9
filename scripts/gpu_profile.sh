#!/bin/bash
# ncu evidence: (1) launch list of the bench command, (2) full capture of the dominant kernels.
mkdir -p gpurun_out
ROUND=${ROUND:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv \
    --log-file gpurun_out/launches_${ROUND}.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 200 -c 4 \
    -o gpurun_out/prof_gemm_${ROUND} -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:"attention_fwd|layernorm_fwd|crf_viterbi" -s 100 -c 3 \
    -o gpurun_out/prof_misc_${ROUND} -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_misc.log 2>&1
echo "misc capture exit $?"
python scripts/crf_sweep.py > gpurun_out/crf_sweep_${ROUND}.json 2> gpurun_out/crf_sweep.err; echo "sweep exit $?"
ls -la gpurun_out
