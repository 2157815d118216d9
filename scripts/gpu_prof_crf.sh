#!/bin/bash
# The CRF kernels alone: parity tests, the configs[4] sweep (L = 13 and 29; optionally with the two-tags-per-lane Viterbi
# above B sentences: VIT_WIDE_MAX), and -- with NCU=1 -- ncu --set full of the three kernels at 4096 x 512 x 13.
#   gpurun --timeout 900 -- env NCU=1 bash scripts/gpu_prof_crf.sh <tag>
set -u
TAG=${1:-crf}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py -m gpu -q -k "crf or viterbi" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for L in 13 29; do
  L=$L SWEEP_B=64,1024,4096,16384 timeout 300 python scripts/crf_sweep.py > $OUT/crf_sweep_L$L.json 2>$OUT/sweep.err; python - <<PY
import json
d=json.load(open("$OUT/crf_sweep_L$L.json"))
for r in d["rows"]: print("L=$L B=%5d" % r["B"], {k:(v["ms"], v["frac_hbm"]) for k,v in r.items() if isinstance(v,dict)})
PY
done
if [ -n "${VIT_WIDE_MAX:-}" ]; then
  for L in 13 29; do
    KBNER_VIT_WIDE_MAX=$VIT_WIDE_MAX L=$L SWEEP_B=4096,16384 timeout 300 python scripts/crf_sweep.py > $OUT/crf_sweep_narrow_L$L.json 2>>$OUT/sweep.err; python - <<PY
import json
d=json.load(open("$OUT/crf_sweep_narrow_L$L.json"))
for r in d["rows"]: print("two tags per lane: L=$L B=%5d" % r["B"], r["viterbi"])
PY
  done
fi
if [ -n "${NCU:-}" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:crf_ -c 6 -o $OUT/crf_L13 -f python scripts/crf_once.py > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu.log
fi
