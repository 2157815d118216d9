#!/bin/bash
# ncu --set full of the three CRF kernels at 4096 x 512 x L (report -> gpurun_out/<tag>/crf_L<L>.ncu-rep) + the sweep.
#   gpurun --timeout 900 -- bash scripts/gpu_prof_crf.sh <tag>
set -u
TAG=${1:-crf}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_precision_gpu.py -m gpu -q -s > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for L in 13 29; do
  L=$L timeout 300 python scripts/crf_sweep.py > $OUT/crf_sweep_L$L.json 2>$OUT/sweep.err; tail -c 1500 $OUT/crf_sweep_L$L.json; echo
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:crf_ -c 6 -o $OUT/crf_L13 -f python scripts/crf_once.py > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu.log
