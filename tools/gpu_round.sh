#!/bin/bash
# Full GPU validation pass: tests, smoke, bench (ours + reference arm), ncu launch list + full capture of one step.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench.json; echo; tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; head -c 700 gpurun_out/bench_ref.json; echo
timeout 600 python bench.py --precision bf16x3 --steps 20 --no-cpu-baseline > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; echo "bench x3 rc=$?"; head -c 300 gpurun_out/bench_bf16x3.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_|warp_mean|pack_input" -s 27 -c 54 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"conv_|warp_mean|pack_input" -s 27 -c 27 -o gpurun_out/prof_step -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
