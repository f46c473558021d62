#!/bin/bash
# Quick GPU iteration: op + net tests, bench, per-launch ablation table.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench.err
timeout 600 python tools/ablate.py 8 1 > gpurun_out/ablate.log 2>&1; cat gpurun_out/ablate.log | tail -34
