#!/bin/bash
# N-GPU validation of the bench contract (one rank per GPU under torchrun, NCCL only for barrier / gradient all-reduce)
N=${1:-2}
mkdir -p gpurun_out
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "infer N=$N exit $?"; cat gpurun_out/bench_n${N}.json; tail -3 gpurun_out/bench_n${N}.err
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload train --steps 8 --warmup 4 > gpurun_out/bench_train_n${N}.json 2> gpurun_out/bench_train_n${N}.err; echo "train N=$N exit $?"; cat gpurun_out/bench_train_n${N}.json; tail -3 gpurun_out/bench_train_n${N}.err
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_n${N}.json 2> gpurun_out/bench_ref_n${N}.err; echo "ref N=$N exit $?"; cat gpurun_out/bench_ref_n${N}.json
