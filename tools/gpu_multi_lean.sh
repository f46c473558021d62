#!/bin/bash
# Lean multi-GPU validation (gpurun --gpus N): unit-sharded plan correctness + bench at N GPUs (both shard modes)
N=${1:-4}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "compress" > gpurun_out/pytest_compress.log 2>&1; echo "pytest compress rc=$?"; tail -n 3 gpurun_out/pytest_compress.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/multigpu_check_$N.log 2>&1; echo "check rc=$?"; grep -E "rank|MULTIGPU|Error|error" gpurun_out/multigpu_check_$N.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_unit_$N.json 2> gpurun_out/bench_unit_$N.err; echo "bench unit rc=$?"; tail -c 1800 gpurun_out/bench_unit_$N.json; tail -n 3 gpurun_out/bench_unit_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 3 --shard scene > gpurun_out/bench_scene_$N.json 2> gpurun_out/bench_scene_$N.err; echo "bench scene rc=$?"; tail -c 400 gpurun_out/bench_scene_$N.json
