#!/bin/bash
# Round-2 multi-GPU pass (gpurun --gpus N): correctness of the unit-sharded plan, then the default bench exactly as the
# driver launches it (unit-sharded, NCCL all-gather, mixed precision) and the neighbour-exchange variant.
N=${1:-8}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/multigpu_check.py > $O/r02_multigpu_check_$N.log 2>&1; echo "check rc=$?"; grep -E "rank 0|MULTIGPU" $O/r02_multigpu_check_$N.log | tail -4
timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 > $O/r02_bench_unit_allgather_${N}gpu.json 2> $O/r02_bench_unit_$N.err; echo "bench allgather rc=$?"; python -c "
import json; d=json.load(open('$O/r02_bench_unit_allgather_${N}gpu.json')); print('allgather', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'det', d['e2e_detections']['value'])"
timeout 400 $TR --master-port 29513 bench.py --gpus $N --steps 30 --warmup 3 --exchange neighbours > $O/r02_bench_unit_neighbours_${N}gpu.json 2>> $O/r02_bench_unit_$N.err; echo "bench neighbours rc=$?"; python -c "
import json; d=json.load(open('$O/r02_bench_unit_neighbours_${N}gpu.json')); print('neighbours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
nvidia-smi topo -m > $O/r02_topo_$N.txt 2>&1; numactl -H >> $O/r02_topo_$N.txt 2>&1
