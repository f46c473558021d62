#!/bin/bash
# Staged GPU bring-up: each stage in its own process (a trapped kernel kills the CUDA context)
# and under its own timeout (a hung kernel must not eat the box).  Logs land in gpurun_out/.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; tail -n 25 gpurun_out/$name.log; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run pack   python -m pytest tests/test_gpu_ops.py -q -x -s -k "pack_input" -p no:cacheprovider
run xcheck python -m pytest tests/test_gpu_ops.py -q -s -k "crosscheck" -p no:cacheprovider
run tc1    python -m pytest tests/test_gpu_ops.py -q -s -k "test_conv_bn_relu and tc and c32_s1" -p no:cacheprovider
run tc     python -m pytest tests/test_gpu_ops.py -q -s -k "tc and not c32_s1" -p no:cacheprovider
run bn     python -m pytest tests/test_gpu_ops.py -q -s -k "block_n" -p no:cacheprovider
run warp   python -m pytest tests/test_gpu_ops.py -q -s -k "warp_mean" -p no:cacheprovider
run nets   python -m pytest tests/test_gpu_nets.py -q -s -p no:cacheprovider
