#!/bin/bash
# ncu --set full captures of selected conv launches (conv_tc launch index within a step: pre_2=1 gru=12 c8_1=21 head1=23)
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; skip=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s $skip -c 1 -o gpurun_out/prof_$name -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
done
