#!/bin/bash
# Quick GPU iteration: op + net tests, bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "when2com|passed|failed|Error|partial|warp_gated" gpurun_out/pytest_gpu.log | tail -40
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/bench.err
