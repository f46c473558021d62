#!/bin/bash
# neighbour-exchange validation at N GPUs: correctness check (default exchange) + bench with both exchanges
N=${1:-2}
mkdir -p gpurun_out
if [ "$2" != "nocheck" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/multigpu_check_$N.log 2>&1; echo "check rc=$?"; grep -E "^rank|MULTIGPU" gpurun_out/multigpu_check_$N.log | tail -10
fi
for X in neighbours allgather; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 --exchange $X > gpurun_out/bench_${X}_$N.json 2> gpurun_out/bench_${X}_$N.err; echo "bench $X rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_${X}_$N.json')); print('$X', d['value'], d['ms_per_step'], d['e2e_detections']['value'])"; tail -n 2 gpurun_out/bench_${X}_$N.err
done
