#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout -k 5 300 $NCU -k regex:attention_fwd_kernel -c 1 -o gpurun_out/prof_attn_fwd_r38 python scripts/attn_bench.py > gpurun_out/ncu_attn_fwd.log 2>&1; echo "attn fwd capture exit $?"
timeout -k 5 300 $NCU -k regex:attention_bwd_kernel -c 1 -o gpurun_out/prof_attn_bwd_r38 python scripts/attn_bench.py > gpurun_out/ncu_attn_bwd.log 2>&1; echo "attn bwd capture exit $?"
SWEEP_B=4096 timeout -k 5 300 $NCU -k regex:crf_ -s 6 -c 4 -o gpurun_out/prof_crf_r38 python scripts/crf_sweep.py > gpurun_out/ncu_crf.log 2>&1; echo "crf capture exit $?"
timeout -k 5 300 $NCU -k regex:"layernorm_bwd|colsum|tagproj_bwd" -c 4 -o gpurun_out/prof_trainhbm_r38 python scripts/train_kernels_bench.py > gpurun_out/ncu_trainhbm.log 2>&1; echo "train hbm capture exit $?"
ls -la gpurun_out/*.ncu-rep
