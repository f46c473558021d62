#!/bin/bash
# Multi-GPU validation (gpurun --gpus N): unit-sharded plan correctness + bench at N GPUs (both shard modes)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/multigpu_check_$N.log 2>&1; echo "check rc=$?"; grep -E "rank|MULTIGPU|Error|error" gpurun_out/multigpu_check_$N.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_unit_$N.json 2> gpurun_out/bench_unit_$N.err; echo "bench unit rc=$?"; tail -c 1500 gpurun_out/bench_unit_$N.json; tail -n 3 gpurun_out/bench_unit_$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline --shard scene > gpurun_out/bench_scene_$N.json 2> gpurun_out/bench_scene_$N.err; echo "bench scene rc=$?"; tail -c 600 gpurun_out/bench_scene_$N.json
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err; echo "bench 1 rc=$?"; head -c 300 gpurun_out/bench_1.json
