#!/bin/bash
# 8-GPU validation: default bench (unit-sharded, NCCL all-gather) exactly as the driver launches it, plus the reference arm
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_unit_$N.json 2> gpurun_out/bench_unit_$N.err; echo "bench unit rc=$?"; head -c 400 gpurun_out/bench_unit_$N.json; echo; tail -n 3 gpurun_out/bench_unit_$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_$N.json 2> gpurun_out/bench_ref_$N.err; echo "ref rc=$?"; head -c 300 gpurun_out/bench_ref_$N.json; echo
