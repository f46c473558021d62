#!/bin/bash
# A/B of the conv kernel variants through the per-launch ablation table, then the GPU parity tests.
mkdir -p gpurun_out
V2X_ONE_CTA=1 timeout 300 python tools/ablate.py 8 1 > gpurun_out/ablate_onecta.log 2>&1; tail -32 gpurun_out/ablate_onecta.log | cut -c1-40
timeout 300 python tools/ablate.py 8 1 > gpurun_out/ablate.log 2>&1; tail -32 gpurun_out/ablate.log
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
