#!/bin/bash
# Round-2 GPU validation pass (one gpurun call): tests, smoke, the four BASELINE configs of bench.py (+ reference arm),
# ncu launch list of the default bench command, ncu --set full of one eager step per config (tools/prof_step.py).
# usage: bash tools/gpu_r02.sh <tag> [quick]
TAG=${1:-r02}; MODE=${2:-full}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 $O/${TAG}_pytest_gpu.log
cp $O/parity.json $O/${TAG}_parity.json 2>/dev/null
timeout 300 python __graft_entry__.py --smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 $O/${TAG}_smoke.log
timeout 900 python bench.py > $O/${TAG}_bench_v2v_det_mixed.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; head -c 400 $O/${TAG}_bench_v2v_det_mixed.json; echo; tail -n 3 $O/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"; head -c 900 $O/${TAG}_bench_reference_arm.json; echo
for p in bf16 fp16x3; do
  timeout 600 python bench.py --precision $p --steps 20 --no-cpu-baseline > $O/${TAG}_bench_v2v_det_$p.json 2>> $O/${TAG}_bench.err; echo "bench $p rc=$?"; head -c 200 $O/${TAG}_bench_v2v_det_$p.json; echo
done
[ "$MODE" = quick ] && exit 0
for c in faf_lower w2c_seg faf_upper_dp; do
  timeout 900 python bench.py --config $c --steps 20 > $O/${TAG}_bench_${c}_mixed.json 2>> $O/${TAG}_bench.err; echo "bench $c rc=$?"; head -c 300 $O/${TAG}_bench_${c}_mixed.json; echo
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
for spec in v2v_det:mixed v2v_det:bf16 faf_lower:mixed w2c_seg:mixed; do
  c=${spec%%:*}; p=${spec##*:}
  timeout 900 ncu --set full --clock-control none --profile-from-start off -o $O/prof_${c}_${p} -f python tools/prof_step.py --config $c --precision $p > $O/${TAG}_ncu_${c}_${p}.log 2>&1; echo "ncu full $c $p rc=$?"
  # the report of a whole step is 60-100 MB (gpurun returns at most 64 MiB in total): keep the raw page as CSV only
  ncu -i $O/prof_${c}_${p}.ncu-rep --page raw --csv > $O/prof_${c}_${p}_raw.csv 2>/dev/null; rm -f $O/prof_${c}_${p}.ncu-rep
done
du -sh $O
