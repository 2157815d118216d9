#!/bin/bash
# N-GPU session (gpurun --gpus N): the 2-rank DDP parity test, then the bench under torchrun -- the driver's own launch
# line -- with the gradient-exchange variants of the fine-tuning leg.
#   gpurun --gpus 2 --timeout 1500 -- bash scripts/gpu_multi.sh <tag> <N> [variants...]      variants: default fp32 overlap
set -u
TAG=${1:-multi}; N=${2:-2}; shift 2 || true
VARIANTS=${*:-default}
OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
nvidia-smi -L | tee $OUT/gpus.txt
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_ddp_gpu.py -m gpu -q -s > $OUT/pytest_ddp.log 2>&1; echo "ddp pytest rc=$?"; tail -4 $OUT/pytest_ddp.log
fi
run() {  # name, workload, env...
  local name=$1 wl=$2; shift 2
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 5 --workload $wl > $OUT/bench_n${N}_$name.json 2> $OUT/bench_n${N}_$name.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_n${N}_$name.json").read().strip().split("\n")[-1])
    t=d.get("train", d)
    print("infer" if "train" in d else "train-only", d.get("value"), d.get("e2e",{}).get("value"), "| train", t.get("value"), t.get("ms_per_step"), t.get("grad_exchange"))
except Exception as e:
    print("no json:", e); print(open("$OUT/bench_n${N}_$name.err").read()[-1500:])
PY
}
for v in $VARIANTS; do
  case $v in
    default) run all all KBNER_X=1 ;;
    fp32) run train_fp32 train KBNER_GRAD_COMM=fp32 ;;
    bf16) run train_bf16 train KBNER_GRAD_COMM=bf16 ;;
    overlap) run train_overlap train KBNER_OVERLAP_ALLREDUCE=1 ;;
    overlap0) run train_overlap_nocarve train KBNER_OVERLAP_ALLREDUCE=1 KBNER_OVERLAP_SM_CARVEOUT=0 ;;
    overlap32) run train_overlap_carve32 train KBNER_OVERLAP_ALLREDUCE=1 KBNER_OVERLAP_SM_CARVEOUT=32 ;;
  esac
done
