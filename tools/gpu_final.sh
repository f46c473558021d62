#!/bin/bash
# Final single-GPU pass of the round: the tests touched last, smoke, the default bench line (+ reference arm), and the ncu
# table of the kernels outside the bench configs' launch lists.   usage: bash tools/gpu_final.sh <tag>
TAG=${1:-r02f}
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_train.py -m gpu -q -s -p no:cacheprovider -k "warp_reduce_bwd or seg_train" 2>&1 | grep -v "^\.$" > $O/${TAG}_pytest_train_subset.log; echo "pytest rc=$?"; tail -n 3 $O/${TAG}_pytest_train_subset.log; grep "warp_reduce_bwd\|BN running" $O/${TAG}_pytest_train_subset.log | cut -c1-200
timeout 300 python __graft_entry__.py --smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O/${TAG}_smoke.log
timeout 900 python bench.py > $O/${TAG}_bench_v2v_det_mixed.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; head -c 300 $O/${TAG}_bench_v2v_det_mixed.json; echo; tail -n 2 $O/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"; head -c 400 $O/${TAG}_bench_reference_arm.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --profile-from-start off --csv --log-file $O/${TAG}_aux.csv python tools/prof_aux.py > $O/${TAG}_ncu_aux.log 2>&1; echo "ncu aux rc=$?"; tail -n 3 $O/${TAG}_ncu_aux.log
python tools/aux_report.py $O/${TAG}_aux.csv $O/${TAG}_aux_kernels.md | head -30
du -sh $O/${TAG}_aux.csv
