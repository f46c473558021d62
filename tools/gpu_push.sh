#!/bin/bash
# N-GPU validation of the device-side x_3 exchange (exchange=push): correctness of every rank's slice against the
# single-GPU plan and the oracle, then value-only bench lines push vs all-gather.   usage: bash tools/gpu_push.sh <N> <tag>
N=${1:-2}; TAG=${2:-r02e}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
V2X_EXCHANGE=push V2X_CHECK_ONLY=v2v timeout -k 10 300 $TR --master-port 29511 tests/multigpu_check.py > $O/${TAG}_multigpu_check_push_${N}gpu.log 2>&1
echo "check rc=$?"; grep -v "^W\|warn" $O/${TAG}_multigpu_check_push_${N}gpu.log | tail -n 12 | cut -c1-400
for ex in push allgather push allgather; do
  timeout -k 10 240 $TR --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 --value-only --exchange $ex >> $O/${TAG}_bench_value_${ex}_${N}gpu.json 2> $O/${TAG}_bench_${ex}_${N}gpu.err
  echo "bench $ex rc=$?"; tail -n 1 $O/${TAG}_bench_value_${ex}_${N}gpu.json | cut -c1-300; tail -n 2 $O/${TAG}_bench_${ex}_${N}gpu.err | cut -c1-300
done
