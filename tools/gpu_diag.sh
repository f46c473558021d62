#!/bin/bash
# Diagnostics: raw MMA rates, per-layer ablation table, source-level ncu captures of selected conv launches.
mkdir -p gpurun_out
timeout 120 tools/mma_probe > gpurun_out/mma_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/mma_probe.txt
timeout 600 python tools/ablate.py 8 1 > gpurun_out/ablate.log 2>&1; tail -34 gpurun_out/ablate.log
bash tools/gpu_prof.sh "$@"
