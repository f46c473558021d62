#!/bin/bash
# final multi-GPU sanity after kernel changes: correctness check + default bench at N GPUs
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/multigpu_check_$N.log 2>&1; echo "check rc=$?"; grep -E "^rank|MULTIGPU" gpurun_out/multigpu_check_$N.log | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_final_$N.json 2> gpurun_out/bench_final_$N.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_final_$N.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_detections']['value'], d['config']['parallelism'][:60])"
